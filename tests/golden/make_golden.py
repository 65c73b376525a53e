#!/usr/bin/env python3
"""Regenerate the golden trace digests from the UNMODIFIED reference.

Runs the reference binaries built by oracle/build_ref.sh (oracle/_ref/<model>_trace, i.e. the real
ROOT-Sim 2.0.0 runtime + the real model + oracle/trace_shim.c) in --sequential mode and stores the
per-LP digests here.  Needs /root/reference (this container only); the fixtures are committed so that
the tests can run where the reference is absent.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rs_trace  # noqa: E402

CASES = [
    # name, model, lps, simulation_time (0 = model's own termination), horizon, state bytes
    ("phold_lp64_full", "phold", 64, 0, None, 8),
    ("phold_lp1024_T50", "phold", 1024, 50, 50, 8),
    ("phold_lp100_T120", "phold", 100, 120, 120, 8),
    ("pcs_lp16_full", "pcs", 16, 0, None, 56),
    ("pcs_lp64_T3000", "pcs", 64, 3000, 3000, 56),
    # models/packet (graph topology, 24-byte payload that carries a pointer) and models/collector (star topology: LP 0
    # receives everything); horizons below the models' own termination (OnGVT)
    ("packet_lp256_T20000", "packet", 256, 20000, 20000, 4),
    ("collector_lp256_T400", "collector", 256, 400, 400, 4),
    # models/traffic on the generated road grids (make_traffic_topology.py; the model reads ./topology.txt in INIT)
    ("traffic_grid_lp16_T20000", "traffic", 16, 20000, 20000, 8),
    ("traffic_grid_lp960_T3000", "traffic", 960, 3000, 3000, 8),
]


def main():
    only = sys.argv[1:]         # optional: regenerate these models only
    for name, model, lps, simt, hor, sb in CASES:
        import subprocess, tempfile
        exe = os.path.join(rs_trace.ORACLE_REF, model + "_trace")
        if only and model not in only:
            continue
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, ".rootsim"))
            with open(os.path.join(td, ".rootsim", "numerical.conf"), "w") as f:
                f.write("%d\n" % rs_trace.MASTER_SEED)
            out = os.path.join(td, "trace.txt")
            if model == "traffic":
                import make_traffic_topology
                text, n = make_traffic_topology.topology(name[:name.rindex("_T")])
                assert n == lps
                with open(os.path.join(td, "topology.txt"), "w") as f:
                    f.write(text)
            env = dict(os.environ, HOME=td, RS_TRACE_OUT=out, RS_TRACE_STATE_BYTES=str(sb))
            if hor is not None:
                env["RS_TRACE_HORIZON"] = repr(float(hor))
            cmd = [exe, "--sequential", "--lp", str(lps), "--deterministic-seed", "--output-dir", os.path.join(td, "o")]
            if simt:
                cmd += ["--simulation-time", str(simt)]
            subprocess.run(cmd, env=env, cwd=td, check=True, capture_output=True)
            shutil.copy(out, os.path.join(HERE, name + ".trace"))
            print("wrote", name)


if __name__ == "__main__":
    main()
