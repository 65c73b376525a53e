"""The CPU oracle (oracle/seqsim.c) is pinned against the real reference: the committed golden digests
(tests/golden/*.trace, produced by the unmodified reference, see make_golden.py) and, where the built
reference travels with the repo (oracle/_ref), a live run of it.  libm flavour => bit-exact."""
import os

import pytest

import rs_trace
from conftest import have_reference_binary

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PHOLD_MASK = [0, 4, 5, 6, 7]            # `traditional` + `events`; the rest of the struct is never written
PCS_MASK = list(range(0, 56))            # counters, lvt, ta (pointers start at byte 56)

CASES = [
    ("phold_lp64_full", "phold", 64, 0, None, 8, PHOLD_MASK),
    ("phold_lp1024_T50", "phold", 1024, 50, 50, 8, PHOLD_MASK),
    ("phold_lp100_T120", "phold", 100, 120, 120, 8, PHOLD_MASK),
    ("pcs_lp16_full", "pcs", 16, 0, None, 56, PCS_MASK),
    ("pcs_lp64_T3000", "pcs", 64, 3000, 3000, 56, PCS_MASK),
    ("packet_lp256_T20000", "packet", 256, 20000, 20000, 4, None),           # state: packet_count
    ("collector_lp256_T400", "collector", 256, 400, 400, 4, None),
    # models/traffic on generated road grids (state: the first 8 bytes are lp_state_type.lvt)
    ("traffic_grid_lp16_T20000", "traffic", 16, 20000, 20000, 8, None),
    ("traffic_grid_lp960_T3000", "traffic", 960, 3000, 3000, 8, None),
]


def model_files(name, model):
    """Input files a model reads inside INIT (models/traffic/init.c:289 opens ./topology.txt)."""
    if model != "traffic":
        return None
    import sys
    sys.path.insert(0, GOLD)
    import make_traffic_topology
    return {"topology.txt": make_traffic_topology.topology(name[:name.rindex("_T")])[0]}


@pytest.mark.parametrize("name,model,lps,simt,hor,sb,mask", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(oracle_built, name, model, lps, simt, hor, sb, mask):
    if not os.path.exists(os.path.join(oracle_built, "seq_" + model)):
        pytest.skip("oracle for %s not built (model sources live in the reference tree)" % model)
    ghdr, grows = rs_trace.parse_digest(os.path.join(GOLD, name + ".trace"))
    hdr, rows, _ = rs_trace.run_oracle(model, lps, sim_time=simt, horizon=hor, portable=False, state_bytes=sb,
                                       files=model_files(name, model))
    assert hdr["total"] == ghdr["total"]
    assert rs_trace.compare(rows, grows, state_mask=mask) == []


@pytest.mark.parametrize("lps,simt", [(37, 90), (256, 70), (2048, 30)])
def test_oracle_matches_live_reference_phold(oracle_built, lps, simt):
    if not have_reference_binary("phold"):
        pytest.skip("oracle/_ref not built")
    _, ref_rows, _ = rs_trace.run_reference("phold", lps, sim_time=simt, horizon=simt)
    _, rows, _ = rs_trace.run_oracle("phold", lps, sim_time=simt, horizon=simt, portable=False)
    assert rs_trace.compare(rows, ref_rows, state_mask=PHOLD_MASK) == []


def test_oracle_matches_live_reference_traffic(oracle_built):
    """models/traffic: INIT parses a file, cars are malloc'ed and freed inside events, RandomRange, a model-local
    Gaussian (log, sqrt), C's integer abs() on a double, exp() in the accident probability."""
    if not have_reference_binary("traffic"):
        pytest.skip("oracle/_ref not built")
    files = model_files("traffic_grid_lp960_T1500", "traffic")
    _, ref_rows, _ = rs_trace.run_reference("traffic", 960, sim_time=1500, horizon=1500, files=files)
    _, rows, _ = rs_trace.run_oracle("traffic", 960, sim_time=1500, horizon=1500, portable=False, files=files)
    assert rs_trace.compare(rows, ref_rows) == []
    assert sum(r[0] for r in rows) > 100000


def test_portable_math_changes_only_ulps_traffic(oracle_built):
    """models/traffic with rs_math.h's log AND exp in the model's own calls (the flavour the device is compared with
    bit-exactly): every integer of the trace equals the libm flavour's, timestamps stay within 64 ULP."""
    if not os.path.exists(os.path.join(oracle_built, "seq_traffic")):
        pytest.skip("traffic oracle not built")
    files = model_files("traffic_grid_lp960_T3000", "traffic")
    _, a, _ = rs_trace.run_oracle("traffic", 960, sim_time=3000, horizon=3000, portable=False, files=files)
    _, b, _ = rs_trace.run_oracle("traffic", 960, sim_time=3000, horizon=3000, portable=True, files=files)
    assert rs_trace.compare(a, b, state_mask=[], ts_ulp=64, check_hash=False) == []


def test_portable_log_changes_only_ulps(oracle_built):
    """The oracle built with rs_math.h's log (the flavour the device is compared with bit-exactly)
    agrees with the libm flavour on every integer of the trace; timestamps stay within 64 ULP."""
    if not os.path.exists(os.path.join(oracle_built, "seq_phold")):
        pytest.skip("phold oracle not built")
    _, a, _ = rs_trace.run_oracle("phold", 512, sim_time=80, horizon=80, portable=False)
    _, b, _ = rs_trace.run_oracle("phold", 512, sim_time=80, horizon=80, portable=True)
    assert rs_trace.compare(a, b, state_mask=PHOLD_MASK, ts_ulp=64, check_hash=False) == []


def test_known_answers_numerical(oracle_built):
    """Known-answer values recorded from the reference's own numerical.o."""
    import ctypes as C
    import rs_host
    lib = rs_host.load(rs_host.build_model("mixnet", emu=True))
    seeds = rs_host.lp_seeds(lib, 12345678901234567, 0, 1024)
    assert seeds[0] == 12345678901234567
    assert seeds[1] == 24691355379668751
    assert seeds[2] == 49382715604938268
    assert seeds[63] == 4162995299808552387
    assert seeds[64] == seeds[0] and seeds[65] == seeds[1] and seeds[1023] == seeds[63]
    hi = rs_host.lp_seeds(lib, 0xF234567890ABCDEF, 0, 64)
    assert hi[0] == 0x1f44567b0042cdf0
    assert hi[1] == hi[2] == 0x2d1000026f970000
    assert hi[63] == 0x262a2b3f4855e6f7
