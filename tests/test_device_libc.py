"""The libc a model may use inside INIT has device bodies in csrc/rs_prelude.cuh (a model that parses an input file:
models/traffic/init.c).  They are plain C++; here they are compiled for the host and checked against glibc."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#define RS_EMU 1
#define RS_DEVICE_BUILD 1
#include <rs_prelude.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static int bad;
static void check_strtod(const char *s)
{
	char *e1, *e2;
	double a = rs_strtod(s, &e1), b = strtod(s, &e2);
	if (memcmp(&a, &b, 8) != 0 || e1 != e2) { printf("strtod(\"%s\"): %a/%ld vs %a/%ld\n", s, a, (long)(e1 - s), b, (long)(e2 - s)); bad++; }
}
int main(void)
{
	const char *fixed[] = { "60", "0.5", "-1", "1", "1.5", "0.05", "3.25e2", "  12.75xyz", "abc", "", ".", "-", "+.5", "5.", "1e", "1e+", "1e5",
				"0.000123", "123456789012345", "9007199254740991", "-0", "1E-5", "007", "1.0e22", "4.2.1", "1e-22", "1e999", "-1e999",
				"1e-999", "0e999" };
	for (unsigned i = 0; i < sizeof(fixed) / sizeof(fixed[0]); i++)
		check_strtod(fixed[i]);
	/* decimals with at most 15 significant digits and a small exponent: correctly rounded, i.e. glibc's bits */
	unsigned long long s = 88172645463325252ULL;
	char buf[64];
	for (int i = 0; i < 200000; i++) {
		s ^= s << 13; s ^= s >> 7; s ^= s << 17;
		const unsigned long long mant = s % 1000000000000000ULL;
		const int frac = (int)((s >> 50) % 16);
		snprintf(buf, sizeof(buf), "%s%llu", (s >> 60) & 1 ? "-" : "", mant);
		const size_t n = strlen(buf);
		if (frac && (size_t)frac < n - ((s >> 60) & 1)) {
			memmove(buf + n - frac + 1, buf + n - frac, frac + 1);
			buf[n - frac] = '.';
		}
		check_strtod(buf);
	}
	/* strtok_r / strcmp / strcpy / strlen on a line of the input format */
	char l1[] = "RO-FI-roma,\t roma,  ,fiera\t,1.5", l2[sizeof(l1)];
	memcpy(l2, l1, sizeof(l1));
	char *s1, *s2, *t1 = rs_strtok_r(l1, ", \t", &s1), *t2 = strtok_r(l2, ", \t", &s2);
	int ntok = 0;
	while (t1 || t2) {
		if (!t1 || !t2 || strcmp(t1, t2) != 0 || rs_strcmp(t1, t2) != 0 || rs_strlen(t1) != strlen(t2)) { printf("strtok_r differs at token %d\n", ntok); bad++; break; }
		ntok++;
		t1 = rs_strtok_r(NULL, ", \t", &s1);
		t2 = strtok_r(NULL, ", \t", &s2);
	}
	if (ntok != 4) { printf("tokens: %d\n", ntok); bad++; }
	if ((rs_strcmp("abc", "abd") < 0) != (strcmp("abc", "abd") < 0) || (rs_strcmp("b", "abc") > 0) != (strcmp("b", "abc") > 0) || rs_strcmp("", "") != 0) bad++;
	char cp[8];
	if (rs_strcpy(cp, "seven_7") != cp || strcmp(cp, "seven_7") != 0) bad++;
	if (rs_c_abs((int)-3.7) != 3 || rs_c_abs((int)0.9) != 0 || rs_atoi(" -42x") != -42) bad++;
	/* fgets on an image: lines keep their newline, a last line without one is returned as is, then NULL */
	const char img[] = "ab\n\ncdefgh\nxy";
	rs_FILE f = { img, (uint32_t)(sizeof(img) - 1), 0 };
	char line[5];
	const char *want[] = { "ab\n", "\n", "cdef", "gh\n", "xy" };
	for (int i = 0; i < 5; i++)
		if (!rs_fgets(line, sizeof(line), &f) || strcmp(line, want[i]) != 0) { printf("fgets %d: '%s'\n", i, line); bad++; }
	if (rs_fgets(line, sizeof(line), &f) || !rs_feof(&f)) bad++;
	rs_rewind(&f);
	if (!rs_fgets(line, sizeof(line), &f) || strcmp(line, "ab\n") != 0 || rs_fgetc(&f) != '\n') bad++;
	printf("bad=%d\n", bad);
	return bad != 0;
}
"""


def test_device_libc_matches_glibc(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    inc = [os.path.join(REPO, "root-sim_b200", "include"), os.path.join(REPO, "root-sim_b200", "csrc"), os.path.join(REPO, "include")]
    subprocess.run(["g++", "-std=gnu++17", "-O1", "-ffp-contract=off", "-fpermissive", "-w"] + ["-I" + i for i in inc] + [str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
