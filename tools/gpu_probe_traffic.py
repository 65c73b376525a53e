#!/usr/bin/env python3
"""models/traffic on the device: a few window configurations of the bench workload (4096-LP road grid, one simulated
hour), parity checksum against the golden one, per-kernel device times of the last run (RS_F_KTIME)."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402
import make_traffic_topology as mt  # noqa: E402

name, T = "traffic_grid_lp4096", 3600.0
text, n = mt.topology(name)
gold = json.load(open(os.path.join(REPO, "tests", "golden", "phold_checksums.json")))["traffic_lp%d_T%g" % (n, T)]
lib = rs_host.build_model("traffic", emu=os.environ.get("RS_EMU") == "1")
devnull = os.open(os.devnull, os.O_WRONLY)
saved = os.dup(1)


def run(flags=0, **cfg):
    eng = rs_host.Engine(lib)
    eng.model_file("topology.txt", text.encode())
    base = dict(window_events=1 << 18, arena_bytes=1 << 17, max_payload=16, bucket_width=4.0, n_buckets=4096)
    base.update(cfg)
    os.dup2(devnull, 1)         # the model prints one line per LP in INIT
    t0 = time.time()
    eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, flags=rs_host.RS_F_TRACE | flags, **base)
    t1 = time.time()
    fin = False
    while not fin and time.time() - t1 < 60:
        fin = eng.run(max_windows=2)
    st = eng.stats()
    os.dup2(saved, 1)
    cs, traced = rs_trace.checksum_engine(eng, 0)
    kt = eng.kernel_times() if flags & rs_host.RS_F_KTIME else None
    eng.fini()
    ok = fin and ("%016x" % cs) == gold["checksum"] and traced == gold["traced_events_incl_init"]
    dev_s = st.device_seconds or st.loop_seconds
    print("TRAFFIC %-60s fin=%d parity=%s init_s=%.3f device_s=%.3f ev/s=%.4g windows=%d rounds=%d rollbacks=%d silent=%d launches=%d"
          % (cfg, fin, "PASS" if ok else "FAIL", t1 - t0, dev_s, (st.committed_events - n) / dev_s, st.windows, st.relax_rounds,
             st.rollbacks, st.silent_events, st.kernel_launches), flush=True)
    return kt


if os.environ.get("RS_PROBE_ONE") == "1":     # one run of the bench configuration (for a profiler)
    run()
    sys.exit(0)
run()
run(window_events=1 << 14)
run(window_events=1 << 16, bucket_width=1.0)
run(window_events=1 << 16, bucket_width=16.0)
run(window_events=1 << 20, bucket_width=16.0, n_buckets=1024)
run(ckpt_period=1)
run(ckpt_period=100)
kt = run(flags=rs_host.RS_F_KTIME)
if kt:
    tot = sum(v[0] for v in kt.values()) or 1.0
    for nm, (ms, launches) in sorted(kt.items(), key=lambda kv: -kv[1][0])[:14]:
        print("  %-22s %9.3f ms %6d launches %5.1f %%" % (nm, ms, launches, 100.0 * ms / tot))
