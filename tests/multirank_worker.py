"""Worker of tests/test_multirank.py: one rank of a multi-rank run of the engine's EMULATION build, with the
engine's two collectives (rs_comm_ops_t) implemented on torch.distributed/gloo over host memory."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "root-sim_b200"))
import rs_host  # noqa: E402
import rs_trace  # noqa: E402


def make_comm(world):
    def allreduce(ctx, buf, n, op, stream):
        arr = np.ctypeslib.as_array((C.c_longlong * n).from_address(buf))
        t = torch.from_numpy(arr)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op else dist.ReduceOp.SUM)
        return 0

    def alltoall(ctx, send, recv, nbytes, stream):
        s = torch.from_numpy(np.ctypeslib.as_array((C.c_ubyte * (nbytes * world)).from_address(send)))
        r = torch.from_numpy(np.ctypeslib.as_array((C.c_ubyte * (nbytes * world)).from_address(recv)))
        dist.all_to_all_single(r, s)
        return 0

    return rs_host.CommOps(None, rs_host.ALLREDUCE_FN(allreduce), rs_host.ALLTOALL_FN(alltoall))


def main():
    spec = json.loads(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = rs_host.lib_path(spec["model"], emu=True)
    eng = rs_host.Engine(lib, private=True)
    eng.model_args(spec.get("margs", []))
    for fn, text in spec.get("files", {}).items():      # input files the model reads inside INIT: every rank loads them
        eng.model_file(fn, text.encode())
    first, count = rs_host.lp_block(eng.lib, spec["n"], world, rank)
    cfg = dict(spec["cfg"], n_lps=spec["n"], rank=rank, nranks=world, lp_first=first, lp_count=count)
    eng.init(**cfg)
    eng.set_comm(make_comm(world))
    err = None
    try:
        assert eng.run()
    except rs_host.RsError as ex:
        if not spec.get("expect_error"):
            raise
        err = ex.code
    st = eng.stats()
    rows = rs_trace.engine_rows(eng, spec["state_bytes"]) if err is None else []
    eng.fini()
    out = {"first": first, "count": count, "rows": [[r[0], r[1], r[2], r[3], r[4].hex()] for r in rows], "error": err,
           "status_names": st.as_dict()["status_names"],
           "stats": {k: v for k, v in st.as_dict().items() if isinstance(v, (int, float))}}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        with open(spec["out"], "w") as f:
            json.dump(gathered, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
