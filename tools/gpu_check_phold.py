#!/usr/bin/env python3
"""Parity check of the BENCH CONFIGURATION itself: models/phold, same sizing as bench.py (phold_cfg), committed
trace of the device run against the CPU oracle on the same box (too slow for pytest at 1M LPs, T=100)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import rs_host  # noqa: E402
import rs_trace  # noqa: E402

n, T = int(sys.argv[1]), float(sys.argv[2])
cfg = bench.phold_cfg(n, T)
cfg["flags"] = rs_host.RS_F_TRACE
eng = rs_host.Engine(rs_host.build_model("phold"), private=True)
eng.init(**cfg)
eng.run()
st = eng.stats()
rows = rs_trace.engine_rows(eng, 8)
eng.fini()
t0 = time.time()
_, orows, out = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T, state_bytes=8)
bad = rs_trace.compare(rows, orows, state_mask=[0, 4, 5, 6, 7])
print("PHOLD_CHECK lps=%d T=%g committed=%d oracle=%d device_s=%.3f ev/s=%.3g windows=%d rollbacks=%d mismatches=%d %s | oracle %.0f s: %s" % (
    n, T, st.committed_events, sum(r[0] for r in orows), st.device_seconds, (st.committed_events - n) / st.device_seconds,
    st.windows, st.rollbacks, len(bad), "PASS" if not bad else bad[:3], time.time() - t0, out.strip().split("\n")[-1]))
