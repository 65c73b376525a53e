"""Debug: one PHOLD run on the emulation build vs the oracle; prints the first differing LPs."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host, rs_trace
from test_emu_parity import PHOLD_MASK as M
n = int(sys.argv[1]); T = float(sys.argv[2])
over = dict(kv.split("=") for kv in sys.argv[3:])
cfg = dict(window_events=1 << 20, arena_bytes=128, max_pending=24 << 20)
for k, v in over.items():
    cfg[k] = float(v) if k == "bucket_width" else int(v)
lib = rs_host.build_model("phold", emu=True)
with rs_host.Engine(lib, private=True) as eng:
    if os.environ.get("RS_DEBUG"):
        eng.lib.rs_set_debug(int(os.environ["RS_DEBUG"]))
    eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, **cfg)
    assert eng.run(); st = eng.stats()
    rows = rs_trace.engine_rows(eng, 8)
_, o, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T, state_bytes=8)
bad = rs_trace.compare(rows, o, state_mask=M)
print("T", T, "windows", st.windows, "rb", st.rollbacks, "bad", len(bad)); [print("  ", b) for b in bad[:40]]
