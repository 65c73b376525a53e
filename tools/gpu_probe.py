#!/usr/bin/env python3
"""Diagnostic run of the device engine: stats + per-kernel CUDA-event times (RS_F_KTIME)."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import rs_host  # noqa: E402


def main():
    n = int(sys.argv[1])
    T = float(sys.argv[2])
    ktime = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    over = dict(kv.split("=") for kv in sys.argv[4:])
    cfg = bench.phold_cfg(n, T)
    cfg["flags"] = rs_host.RS_F_KTIME if ktime else 0
    for k, v in over.items():
        cfg[k] = float(v) if k in ("bucket_width", "window_span") else int(v)
    eng = rs_host.Engine(os.environ.get("RS_LIB") or rs_host.build_model("phold"))
    if os.environ.get("RS_DEBUG"):
        eng.lib.rs_set_debug(int(os.environ["RS_DEBUG"]))
    seeds = None
    if os.environ.get("RS_DISTINCT_SEEDS"):
        import ctypes as C
        import numpy as np
        g = np.arange(n, dtype=np.uint64)
        z = (g * np.uint64(0x9E3779B97F4A7C15) + np.uint64(bench.MASTER_SEED))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        lo = (z & np.uint64(0xFFFFFFFF)) % np.uint64(0x9068FFFE) + np.uint64(1)
        hi = (z >> np.uint64(32)) % np.uint64(0x464FFFFE) + np.uint64(1)
        dseeds = ((hi << np.uint64(32)) | lo).astype(np.uint64)
        seeds = (C.c_uint64 * n).from_buffer(dseeds)
    if os.environ.get("RS_SAME_SEEDS"):
        # every LP is a twin of every other: pure same-timestamp bursts on single LPs (profiling aid)
        import ctypes as C
        import numpy as np
        one = rs_host.lp_seeds(eng.lib, bench.MASTER_SEED, 0, 1)[0]
        dseeds = np.full(n, one, dtype=np.uint64)
        seeds = (C.c_uint64 * n).from_buffer(dseeds)
    t0 = time.time()
    eng.init(host_seeds=seeds, **cfg)
    t1 = time.time()
    fin = eng.run()
    t2 = time.time()
    st = eng.stats().as_dict()
    kt = eng.kernel_times()
    hot = None
    if os.environ.get("RS_HOTSTATS"):
        import numpy as np
        tr = eng.trace()
        cnt = np.sort(np.frombuffer(tr, dtype=[("count", "u8"), ("hash", "u8"), ("last_ts", "f8"), ("seed", "u8")])["count"].astype(np.int64))[::-1]
        hot = {"top": [int(x) for x in cnt[:8]], "over_1e4": int((cnt > 10000).sum()), "over_1e5": int((cnt > 100000).sum())}
    eng.fini()
    ev = st["committed_events"] - n
    print(json.dumps({"lps": n, "T": T, "init_s": t1 - t0, "run_s": t2 - t1, "events_per_s_device": ev / st["device_seconds"],
                      "stats": st, "hot": hot, "kernel_ms": {k: [round(v[0], 3), v[1]] for k, v in sorted(kt.items(), key=lambda x: -x[1][0])}}))


if __name__ == "__main__":
    main()
