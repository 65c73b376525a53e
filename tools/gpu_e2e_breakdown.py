"""Where the end-to-end time of one bench step goes: init / run / stats / trace / fini (second engine of the process)."""
import os, sys, time, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200")); sys.path.insert(0, REPO)
import bench, rs_host
sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_trace
n, T = int(sys.argv[1]), float(sys.argv[2])
lib_path = rs_host.build_model("phold")
lib = rs_host.load(lib_path)
if os.environ.get("RS_DEBUG"):
    lib.rs_set_debug(int(os.environ["RS_DEBUG"]))
seeds = rs_host.lp_seeds(lib, bench.MASTER_SEED, 0, n)
cfg = bench.phold_cfg(n, T)
cfg["flags"] = rs_host.RS_F_TRACE
for it in range(2):
    t = [time.perf_counter()]
    eng = rs_host.Engine(lib_path); eng.init(host_seeds=seeds, **cfg); t.append(time.perf_counter())
    eng.run(); t.append(time.perf_counter())
    st = eng.stats(); t.append(time.perf_counter())
    tr = eng.trace(); t.append(time.perf_counter())
    cs = rs_trace.checksum_engine(eng, 0); t.append(time.perf_counter())
    eng.fini(); t.append(time.perf_counter())
    d = dict(zip(["init", "run", "stats", "trace", "trace+checksum", "fini"], [round(b - a, 4) for a, b in zip(t, t[1:])]))
    d["device_s"] = round(st.device_seconds, 4)
    print(json.dumps(d))
