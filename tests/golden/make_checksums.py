#!/usr/bin/env python3
"""Golden checksums of the configurations bench.py times (BASELINE.json configs[1], [2], [3] and [4]).

Runs the CPU oracle oracle/_build/seqp_phold (rs_math.h log: bit-identical with the device build) on the
exact bench configuration and stores the order-independent checksum of the per-LP committed-trace digests
(tests/rs_trace.py::checksum_arrays) in phold_checksums.json.  bench.py recomputes the same checksum from
the device engine's digests in its end-to-end steps, at every GPU count, and fails on a mismatch.

  python tests/golden/make_checksums.py                 # 1,048,576 LPs, T=100  (~8 min, 5 GB)
  python tests/golden/make_checksums.py 16777216 60     # 16,777,216 LPs, T=60  (~25 min, 30 GB)
  python tests/golden/make_checksums.py 65536 400 - pcs # models/pcs, 65,536 cells, T=400 (BASELINE configs[2], ~7 min)
  python tests/golden/make_checksums.py 4096 3600 - traffic  # models/traffic on the 4096-LP road grid, one simulated hour
  python tests/golden/make_checksums.py 68608 3600 - traffic # ... on the 68,608-LP grid (BASELINE configs[4]; ~8 min, INIT is O(LPs x file))
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rs_trace  # noqa: E402

OUT = os.path.join(HERE, "phold_checksums.json")


def key(n, T, model="phold"):
    return "%s_lp%d_T%g" % (model, n, T)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    T = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    digest = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None      # an existing digest file of exactly this run
    model = sys.argv[4] if len(sys.argv) > 4 else "phold"
    exe = os.path.join(rs_trace.ORACLE_BUILD, "seqp_" + model)
    with tempfile.TemporaryDirectory() as td:
        if digest is None:
            digest = os.path.join(td, "trace.txt")
            # PHOLD events carry no payload: pool the oracle's event records without payload room (16M LPs fit in memory)
            env = dict(os.environ, RS_TRACE_OUT=digest, RS_TRACE_HORIZON=repr(T))
            if model == "phold":
                env["RS_ORACLE_MAX_PAYLOAD"] = "0"
            if model == "traffic":      # the road grid with this many LPs (make_traffic_topology.py); INIT reads ./topology.txt
                import make_traffic_topology as mt
                name = [k for k in mt.GRIDS if mt.n_lps(*mt.GRIDS[k]) == n][0]
                with open(os.path.join(td, "topology.txt"), "w") as f:
                    f.write(mt.topology(name)[0])
            subprocess.run([exe, "--lp", str(n), "--seed", str(rs_trace.MASTER_SEED), "--simulation-time", repr(T)], env=env, check=True,
                           cwd=td, stdout=subprocess.DEVNULL if model == "traffic" else None)
        cs, total = rs_trace.checksum_digest_file(digest)
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    db[key(n, T, model)] = {"model": model, "n_lps": n, "simulation_time": T, "master_seed": rs_trace.MASTER_SEED,
                            "oracle": "oracle/seqsim.c, rs_math.h log (seqp_%s)" % model, "committed_events": total - n,
                            "traced_events_incl_init": total, "checksum": "%016x" % cs}
    json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)
    print(key(n, T, model), db[key(n, T, model)])


if __name__ == "__main__":
    main()
