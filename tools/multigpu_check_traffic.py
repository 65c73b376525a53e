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


def run(name, T, rank, world, local, dist, torch, window_events=1 << 20):
    """One run of models/traffic on the named road grid over the ranks of an initialised process group (world 1: no
    group needed).  Returns a dict on every rank (the aggregate values are those of the whole job)."""
    text, n = mt.topology(name)
    gp = os.path.join(REPO, "tests", "golden", "phold_checksums.json")
    gold = json.load(open(gp)).get("traffic_lp%d_T%g" % (n, T)) if os.path.exists(gp) else None
    lib_path = rs_host.build_model("traffic")
    lib = rs_host.load(lib_path)
    first, count = rs_host.lp_block(lib, n, world, rank)
    eng = rs_host.Engine(lib_path)
    eng.model_file("topology.txt", text.encode())
    sys.stdout.flush()
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)             # the model prints a line per LP in INIT
    try:
        t0 = time.time()
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=window_events, arena_bytes=1 << 17,
                 max_payload=16, bucket_width=4.0, n_buckets=4096, device=local, rank=rank, nranks=world, lp_first=first,
                 lp_count=count, flags=rs_host.RS_F_TRACE)
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
        fin = eng.run()
        t_run = time.time() - t1
        st = eng.stats()
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    cs, traced = rs_trace.checksum_engine(eng, first)
    eng.fini()
    eng.model_file("topology.txt", None)
    mine = (int(cs), int(traced), st.committed_events, st.device_seconds, t_init, t_run, st.rollbacks, st.windows, st.relax_rounds,
            int(bool(fin) and (st.status & ~0x1) == 0))
    gathered = [mine]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    total_cs = sum(g[0] for g in gathered) & ((1 << 64) - 1)
    total_traced = sum(g[1] for g in gathered)
    committed = sum(g[2] for g in gathered) - n
    dev = max(g[3] for g in gathered)
    finished = all(g[9] for g in gathered)
    ok = None if gold is None else (finished and ("%016x" % total_cs) == gold["checksum"] and total_traced == gold["traced_events_incl_init"])
    return {"workload": "models/traffic (unmodified reference sources), generated road grid %s (%d LPs), --simulation-time %g, master seed %d"
                        % (name, n, T, rs_trace.MASTER_SEED), "n_gpus": world, "lps": n, "unit": "events/s",
            "value": committed / dev if dev > 0 else None, "committed_events": committed, "device_s": dev, "finished": finished,
            "checksum": "%016x" % total_cs, "parity_checksum_ok": ok, "init_s_per_rank_incl_file_parse_on_device": [round(g[4], 2) for g in gathered],
            "run_wall_s": max(g[5] for g in gathered), "rollbacks": sum(g[6] for g in gathered), "windows": gathered[0][7], "rounds": gathered[0][8]}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "traffic_grid_lp68608"
    T = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = run(name, T, rank, world, local, dist, torch)
    if rank == 0:
        print("TRAFFIC_MULTIGPU grid=%s world=%d lps=%d T=%g committed=%d checksum=%s parity=%s device_s=%.3f ev/s=%.4g init_s=%s run_wall_s=%.3f "
              "rollbacks=%d windows=%d rounds=%d" % (name, world, r["lps"], T, r["committed_events"], r["checksum"],
                                                     {None: "NO-GOLDEN", True: "PASS", False: "FAIL"}[r["parity_checksum_ok"]], r["device_s"],
                                                     r["value"], r["init_s_per_rank_incl_file_parse_on_device"], r["run_wall_s"], r["rollbacks"],
                                                     r["windows"], r["rounds"]), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
