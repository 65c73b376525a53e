#!/usr/bin/env python3
"""models/traffic on several GPUs (run under torchrun, one rank per GPU; BASELINE config 5): a generated road grid
block-partitioned over the ranks, every rank loads the topology image and runs INIT for its own LPs on its GPU, exchange
by direct stores into the peers' memory.  The checksum of the union of the ranks' committed traces is compared with the
CPU oracle's golden checksum (tests/golden/phold_checksums.json, tests/golden/make_checksums.py <n> 3600 - traffic).
  torchrun --nproc-per-node 2 tools/multigpu_check_traffic.py [grid name] [T]"""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402
import make_traffic_topology as mt  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "traffic_grid_lp68608"
    T = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    text, n = mt.topology(name)
    gold = json.load(open(os.path.join(REPO, "tests", "golden", "phold_checksums.json"))).get("traffic_lp%d_T%g" % (n, T))
    lib_path = rs_host.build_model("traffic")
    lib = rs_host.load(lib_path)
    first, count = rs_host.lp_block(lib, n, world, rank)
    eng = rs_host.Engine(lib_path)
    eng.model_file("topology.txt", text.encode())
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)             # the model prints a line per LP in INIT
    t0 = time.time()
    eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=1 << 20, arena_bytes=1 << 17, max_payload=16,
             bucket_width=4.0, n_buckets=4096, device=local, rank=rank, nranks=world, lp_first=first, lp_count=count,
             flags=rs_host.RS_F_TRACE)
    t_init = time.time() - t0
    if world > 1:
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            assert lib.rs_nccl_unique_id(idbuf, None) == 0
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
        eng.comm_init_nccl(idbuf)
        if not os.environ.get("RS_NO_IPC"):
            mine = torch.tensor(list(eng.ipc_export()), dtype=torch.uint8, device="cuda")
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            eng.ipc_import([bytes(h.cpu().tolist()) for h in allh])
    t1 = time.time()
    assert eng.run()
    t_run = time.time() - t1
    os.dup2(saved, 1)
    st = eng.stats()
    cs, traced = rs_trace.checksum_engine(eng, first)
    eng.fini()
    mine = (int(cs), int(traced), st.committed_events, st.device_seconds, t_init, t_run, st.rollbacks, st.windows, st.relax_rounds)
    gathered = [mine]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    if rank == 0:
        total_cs = sum(g[0] for g in gathered) & ((1 << 64) - 1)
        total_traced = sum(g[1] for g in gathered)
        committed = sum(g[2] for g in gathered) - n
        dev = max(g[3] for g in gathered)
        ok = None if gold is None else (("%016x" % total_cs) == gold["checksum"] and total_traced == gold["traced_events_incl_init"])
        print("TRAFFIC_MULTIGPU grid=%s world=%d lps=%d T=%g committed=%d checksum=%016x parity=%s device_s=%.3f ev/s=%.4g init_s=%s run_wall_s=%.3f "
              "rollbacks=%d windows=%d rounds=%d" % (name, world, n, T, committed, total_cs, {None: "NO-GOLDEN", True: "PASS", False: "FAIL"}[ok], dev,
                                                     committed / dev, [round(g[4], 2) for g in gathered], max(g[5] for g in gathered),
                                                     sum(g[6] for g in gathered), gathered[0][7], gathered[0][8]), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
