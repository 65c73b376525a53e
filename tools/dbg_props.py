"""Debug: run PHOLD under two window configurations (emulation build or GPU) and compare with the oracle."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host, rs_trace
emu = os.environ.get("RS_EMU", "1") == "1"
n = int(sys.argv[1]); T = float(sys.argv[2])
lib = rs_host.build_model("phold", emu=emu)
def run(**cfg):
    with rs_host.Engine(lib, private=True) as eng:
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, **cfg)
        t = time.time(); assert eng.run(); st = eng.stats()
        rows = rs_trace.engine_rows(eng, 8)
        print("run", cfg, "windows", st.windows, "rb", st.rollbacks, "silent", st.silent_events, "ckpt", st.checkpoints, "%.1fs" % (time.time() - t), flush=True)
    return rows
a = run(window_events=1 << 20, arena_bytes=128, max_pending=24 << 20)
b = run(window_events=1 << 18, bucket_width=0.125, arena_bytes=128, max_pending=24 << 20)
_, o, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T, state_bytes=8)
from test_emu_parity import PHOLD_MASK as M
print("a vs b", rs_trace.compare(a, b, state_mask=M)[:5])
print("a vs oracle", rs_trace.compare(a, o, state_mask=M)[:5])
print("b vs oracle", rs_trace.compare(b, o, state_mask=M)[:5])
