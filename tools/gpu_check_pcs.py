#!/usr/bin/env python3
"""One-off long-horizon parity check of models/pcs on the device against the CPU oracle (too slow for pytest)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402

n, T = int(sys.argv[1]), float(sys.argv[2])
eng = rs_host.Engine(rs_host.build_model("pcs"), private=True)
eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=1 << 20, arena_bytes=32768, max_payload=48,
         bucket_width=0.25, n_buckets=4096, max_pending=256 * n + (1 << 20), gvt_snapshot_cycles=1 << 30)
t0 = time.time()
eng.run()
t1 = time.time()
st = eng.stats()
rows = rs_trace.engine_rows(eng, 56)
eng.fini()
_, orows, out = rs_trace.run_oracle("pcs", n, sim_time=T, horizon=T, state_bytes=56)
bad = rs_trace.compare(rows, orows)
print("PCS_CHECK lps=%d T=%g committed=%d oracle=%d device_s=%.3f ev/s=%.3g rollbacks=%d mismatches=%d %s | oracle: %s" % (
    n, T, st.committed_events, sum(r[0] for r in orows), st.device_seconds, (st.committed_events - n) / st.device_seconds,
    st.rollbacks, len(bad), "PASS" if not bad else bad[:3], out.strip().split("\n")[-1]))
