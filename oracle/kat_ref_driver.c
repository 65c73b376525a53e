/* oracle/kat_ref_driver.c -- TEST INFRASTRUCTURE ONLY.  Calls the UNMODIFIED reference's compiled numerical.o
 * (oracle/_ref/obj/rootsim_lib_numerical_c.o, built by oracle/build_ref.sh from /root/reference/src/lib/numerical.c)
 * with the RNG state set to the seed of the reference's own tests/numerical.c:149, and prints the same lines as
 * tests/kat_numerical.c does for the oracle's restatement: tests/test_numerical_kat.py compares the two outputs and
 * the known answers of SURVEY.md section 8(c).  Compiled against the reference headers where they lie; nothing is
 * copied.  kat_ref_stubs.c satisfies the few runtime symbols numerical.o refers to outside these calls. */
#include <stdio.h>
#include <string.h>
#include <core/core.h>
#include <scheduler/process.h>
#include <lib/numerical.h>
__thread struct lp_struct *current;
int main(void)
{
	static struct lp_struct lp;
	memset(&lp, 0, sizeof lp);
	lp.numerical.seed = 7319936632422683443ULL;
	current = &lp;
	for (int i = 0; i < 4; i++) { double r = Random(); printf("random %.17g %llu\n", r, (unsigned long long)lp.numerical.seed); }
	double e = Expent(5.0);
	printf("expent %.17g %llu\n", e, (unsigned long long)lp.numerical.seed);
	int rr = RandomRange(0, 99);
	printf("randomrange %d\n", rr);
	double n1 = Normal(); double n2 = Normal();
	printf("normal %.17g %.17g\n", n1, n2);
	int rrnu = RandomRangeNonUniform(7, 3, 40);
	printf("rrnu %d\n", rrnu);
	double g3 = Gamma(3); printf("gamma3 %.17g\n", g3);
	double g8 = Gamma(8); printf("gamma8 %.17g\n", g8);
	double p = Poisson(); printf("poisson %.17g\n", p);
	int z = Zipf(1.5, 1000); printf("zipf %d\n", z);
	printf("seed %llu\n", (unsigned long long)lp.numerical.seed);
	return 0;
}
