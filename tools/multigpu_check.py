#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU): PHOLD block-partitioned over the ranks with
the built-in NCCL transport; the union of the per-rank committed traces is compared bit-exactly with the
single-engine CPU oracle (oracle/_build/seqp_phold)."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    T = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib_path = rs_host.build_model("phold")
    lib = rs_host.load(lib_path)
    first, count = rs_host.lp_block(lib, n, world, rank)
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        assert lib.rs_nccl_unique_id(idbuf, None) == 0
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
    eng = rs_host.Engine(lib_path)
    eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=1 << 20, arena_bytes=128,
             device=local, rank=rank, nranks=world, lp_first=first, lp_count=count, max_pending=(n * 8 + (1 << 20)))
    eng.comm_init_nccl(idbuf)
    if not os.environ.get("RS_NO_IPC"):
        mine = torch.tensor(list(eng.ipc_export()), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        eng.ipc_import([bytes(h.cpu().tolist()) for h in allh])
    assert eng.run()
    st = eng.stats()
    rows = rs_trace.engine_rows(eng, 8)
    eng.fini()
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, [(r[0], r[1], r[2], r[3], r[4].hex()) for r in rows], st.as_dict()))
    if rank == 0:
        allrows = []
        for f, rr, _ in sorted(gathered):
            allrows += [(a, b, c, d, bytes.fromhex(e)) for a, b, c, d, e in rr]
        _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
        bad = rs_trace.compare(allrows, orows, state_mask=[0, 4, 5, 6, 7])
        tot = sum(g[2]["committed_events"] for g in gathered)
        late = sum(g[2]["late_events"] for g in gathered)
        print("MULTIGPU_CHECK exchange=%s world=%d lps=%d T=%g committed=%d (oracle %d) windows=%s device_s=%s mismatches=%d %s" % (
            "nccl-sendrecv" if os.environ.get("RS_NO_IPC") else "nvlink-peer-push", world, n, T, tot, sum(r[0] for r in orows), [g[2]["windows"] for g in gathered],
            [round(g[2]["device_seconds"], 3) for g in gathered], len(bad), "PASS" if not bad else bad[:3]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
