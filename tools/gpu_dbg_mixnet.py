import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host, rs_trace
import subprocess
subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "own"], check=True)
for margs in ([], ["--mean", "4.0"], ["--mean", "2.0"]):
    for private in (True, False):
        eng = rs_host.Engine(rs_host.build_model("mixnet"), private=private)
        eng.model_args(margs)
        eng.init(n_lps=100, master_seed=42, simulation_time=50.0, window_events=1000, arena_bytes=1024, max_payload=16, bucket_width=0.25)
        eng.run()
        rows = rs_trace.engine_rows(eng, 32)
        st = eng.stats().as_dict()
        eng.fini()
        _, orows, _ = rs_trace.run_oracle("mixnet", 100, sim_time=50.0, horizon=50.0, state_bytes=32, seed=42, extra=margs)
        bad = rs_trace.compare(rows, orows)
        print(margs, "private" if private else "shared", "mismatches", len(bad), "committed", st["committed_events"], sum(r[0] for r in orows), "status", st["status_names"])
        if bad: print("   ", bad[:3], rows[0][4].hex(), orows[0][4].hex())
