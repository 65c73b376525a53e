#!/usr/bin/env python3
"""Diagnostic run of models/pcs on the device engine (BASELINE config 3: 65536 cells)."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402


def main():
    n = int(sys.argv[1])
    T = float(sys.argv[2])
    over = dict(kv.split("=") for kv in sys.argv[3:])
    eng = rs_host.Engine(os.environ.get("RS_LIB") or rs_host.build_model("pcs"))
    if os.environ.get("RS_DEBUG"):
        eng.lib.rs_set_debug(int(os.environ["RS_DEBUG"]))
    t0 = time.time()
    cfg = dict(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=1 << 20, arena_bytes=32768,
               max_payload=48, bucket_width=0.25, n_buckets=4096, max_pending=256 * n + (1 << 20), flags=rs_host.RS_F_KTIME,
               gvt_snapshot_cycles=1 << 30)
    for k, v in over.items():
        cfg[k] = float(v) if k == "bucket_width" else int(v)
    eng.init(**cfg)
    t1 = time.time()
    eng.run()
    t2 = time.time()
    st = eng.stats().as_dict()
    kt = eng.kernel_times()
    eng.fini()
    ev = st["committed_events"] - n
    print(json.dumps({"model": "pcs", "lps": n, "T": T, "init_s": t1 - t0, "run_s": t2 - t1, "events_per_s_device": ev / st["device_seconds"],
                      "avg_ckpt_bytes": st["checkpoint_bytes"] / max(1, st["checkpoints"]),
                      "stats": st, "kernel_ms": {k: [round(v[0], 3), v[1]] for k, v in sorted(kt.items(), key=lambda x: -x[1][0])}}))


if __name__ == "__main__":
    main()
