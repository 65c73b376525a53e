"""Serial handler latency on one device thread: python tools/gpu_probe_latency.py [model] [type]"""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
import rs_host
model = sys.argv[1] if len(sys.argv) > 1 else "phold"
etype = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = os.environ.get("RS_LIB") or rs_host.lib_path(model)
out = {}
with rs_host.Engine(lib, private=True) as eng:
    eng.init(n_lps=1024, master_seed=1, simulation_time=10.0, max_out=1 << 20)
    for silent in (True, False):
        for it in (1000, 20000):
            out["silent=%d iters=%d" % (silent, it)] = round(eng.probe_event(5, etype, it, silent), 1)
print(json.dumps(out))
