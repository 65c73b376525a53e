"""Known-answer tests of the numerical library of the model API (src/lib/numerical.c:54-249): the oracle's
restatement against (a) the known answers SURVEY.md section 8(c) lists for the seed of the reference's own
tests/numerical.c:149, (b) the UNMODIFIED reference's compiled numerical.o run in this container (every
distribution, incl. Normal, Gamma, Poisson, Zipf and RandomRangeNonUniform), and the per-LP seed derivation."""
import os
import subprocess

import pytest

import rs_host
import rs_trace

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(REPO, "root-sim_b200", "include")


def build_and_run(tmp_path, portable):
    exe = str(tmp_path / ("kat_p" if portable else "kat"))
    cmd = ["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-w", "-I", INC] + (["-DRS_ORACLE_PORTABLE_MATH"] if portable else []) + \
        [os.path.join(REPO, "tests", "kat_numerical.c"), "-lm", "-o", exe]
    subprocess.run(cmd, check=True)
    return subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().split("\n")


def test_oracle_known_answers(tmp_path):
    out = build_and_run(tmp_path, False)
    kv = [l.split() for l in out]
    # Random() #1..4 and the seed after each
    assert [x[1] for x in kv[:4]] == ["0.62020022137081243", "0.88392965221594566", "0.75086992594838609", "0.49541479745161959"]
    assert [x[2] for x in kv[:4]] == ["3237597909594632663", "1309129663808262174", "4878398447187033221", "2801303647315449843"]
    assert kv[4] == ["expent", "0.51527206313459573", "1164245169535256809"]
    assert kv[5] == ["randomrange", "97"]
    # the two deviates of one Box-Muller pair (SURVEY lists them in the order a right-to-left printf printed them)
    assert sorted(kv[6][1:]) == sorted(["-0.46015168802038003", "-1.2907917928700912"])
    assert kv[6][1] == "-1.2907917928700912"          # first call returns v2*fac (numerical.c:159-164), confirmed below
    seeds = {x[1]: x[2] for x in kv if x[0] == "lpseed"}
    assert seeds == {"0": "12345678901234567", "1": "24691355379668751", "2": "49382715604938268"}


def test_oracle_equals_reference_numerical_object(tmp_path):
    """every distribution of numerical.c, oracle restatement (libm flavour) == the reference's own object code"""
    obj = os.path.join(REPO, "oracle", "_ref", "obj", "rootsim_lib_numerical_c.o")
    ref_src = os.environ.get("RS_REFERENCE", "/root/reference")
    if not os.path.exists(obj) or not os.path.isdir(os.path.join(ref_src, "src")):
        pytest.skip("the reference is not built here")
    gen = os.path.join(REPO, "oracle", "_ref", "gen")
    exe = str(tmp_path / "refkat")
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-DOS_LINUX", "-DNDEBUG", "-DHAVE_LP_REBINDING", "-D_GNU_SOURCE", "-include",
                    os.path.join(gen, "cfg.h"), "-w", "-I", gen, "-I", os.path.join(ref_src, "src"),
                    os.path.join(REPO, "oracle", "kat_ref_driver.c"), os.path.join(REPO, "oracle", "kat_ref_stubs.c"), obj, "-lm", "-o", exe],
                   check=True)
    ref = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().split("\n")
    mine = [l for l in build_and_run(tmp_path, False) if not l.startswith("lpseed")]
    assert mine == ref


def test_portable_log_flavour_keeps_the_integers(tmp_path):
    """rs_math.h log instead of glibc log: integer draws and the RNG stream position are unchanged, doubles within 2 ULP"""
    import struct
    a = [l.split() for l in build_and_run(tmp_path, False)]
    b = [l.split() for l in build_and_run(tmp_path, True)]
    for x, y in zip(a, b):
        assert x[0] == y[0]
        for u, v in zip(x[1:], y[1:]):
            if "." in u:
                ub, vb = (struct.unpack("<q", struct.pack("<d", float(t)))[0] for t in (u, v))
                assert abs(ub - vb) <= 2, (x, y)
            else:
                assert u == v, (x, y)


# bytes of rngtail's lp_state: 0-23 integers, 24-47 doubles from log/sqrt (bit-exact), 48-55 Gamma(8) (exp)
def rngtail_compare(rows, orows, exp_ulp):
    import struct
    assert rs_trace.compare(rows, orows, state_mask=list(range(48))) == []
    worst = 0
    for a, b in zip(rows, orows):
        ga, gb = (struct.unpack("<q", s[4][48:56])[0] for s in (a, b))
        worst = max(worst, abs(ga - gb))
    assert worst <= exp_ulp, worst


def test_rngtail_emu_vs_oracle(oracle_built):
    """every numerical.c distribution inside events: engine (emulation build) == oracle, all fields bit-exact"""
    lib = rs_host.build_model("rngtail", emu=True)
    n, T = 500, 150.0
    with rs_host.Engine(lib, private=True) as eng:
        eng.init(n_lps=n, master_seed=77, simulation_time=T, window_events=4096, arena_bytes=256)
        assert eng.run()
        rows = rs_trace.engine_rows(eng, 56)
    _, orows, _ = rs_trace.run_oracle("rngtail", n, sim_time=T, horizon=T, state_bytes=56, seed=77)
    rngtail_compare(rows, orows, 0)
    assert sum(r[0] for r in rows) > 20 * n


@pytest.mark.gpu
def test_rngtail_gpu_vs_oracle(oracle_built):
    """Device port of numerical.c (rs_prelude.cuh) against the oracle: counts, event hash, timestamps, seeds, every
    integer draw, Normal/Gamma(3)/Poisson bit-exact (rs_math.h log + IEEE sqrt); Gamma(8) goes through CUDA's exp():
    stated tolerance 4 ULP on the value, and the RNG stream position after it (part of the seed) must still be equal."""
    lib = rs_host.build_model("rngtail")
    n, T = 20000, 150.0
    with rs_host.Engine(lib) as eng:
        eng.init(n_lps=n, master_seed=77, simulation_time=T, window_events=1 << 18, arena_bytes=256)
        assert eng.run()
        rows = rs_trace.engine_rows(eng, 56)
    _, orows, _ = rs_trace.run_oracle("rngtail", n, sim_time=T, horizon=T, state_bytes=56, seed=77)
    rngtail_compare(rows, orows, 4)
