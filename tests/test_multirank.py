"""Multi-rank path on the CPU: world_size 2 and 3 over torch.distributed/gloo, engine in emulation build.
The LPs are block-partitioned like the reference (src/core/core.c:275-300); events that cross ranks are
exchanged once per straggler round / once per window; every rank must take the same window decisions.
The union of the per-rank committed traces must equal the single-engine oracle BIT-EXACTLY."""
import json
import os
import socket
import subprocess
import sys

import pytest

import rs_host
import rs_trace

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_ranks(world, spec, tmp_path):
    spec = dict(spec, out=str(tmp_path / "out.json"))
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multirank_worker.py"), json.dumps(spec)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    got = json.load(open(spec["out"]))
    rows = []
    for g in sorted(got, key=lambda g: g["first"]):
        rows += [(r[0], r[1], r[2], r[3], bytes.fromhex(r[4])) for r in g["rows"]]
    return rows, got


@pytest.mark.parametrize("world,n,T", [(2, 512, 60), (3, 1000, 40), (2, 4096, 30)])
def test_phold_multirank_vs_oracle(oracle_built, tmp_path, world, n, T):
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    rs_host.build_model("phold", emu=True)
    spec = {"model": "phold", "n": n, "state_bytes": 8,
            "cfg": dict(master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=4096, arena_bytes=128)}
    rows, got = run_ranks(world, spec, tmp_path)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=[0, 4, 5, 6, 7]) == []
    assert sum(g["stats"]["committed_events"] for g in got) == sum(r[0] for r in orows)
    assert len({g["stats"]["windows"] for g in got}) == 1          # lock step: same number of windows everywhere


def test_phold_multirank_far_list(oracle_built, tmp_path):
    """a calendar ring of 4 time units on every rank: most events (mean increment 5) wait on the ranks' far lists and
    the agreed first bucket must take them into account"""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    rs_host.build_model("phold", emu=True)
    n, T = 700, 50
    spec = {"model": "phold", "n": n, "state_bytes": 8,
            "cfg": dict(master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=4096, arena_bytes=128,
                        n_buckets=64, bucket_width=0.0625)}
    rows, got = run_ranks(3, spec, tmp_path)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=[0, 4, 5, 6, 7]) == []
    assert sum(g["stats"]["committed_events"] for g in got) == sum(r[0] for r in orows)


def test_mixnet_multirank_vs_oracle(oracle_built, tmp_path):
    """payloads, zero-delay events across ranks (antimessages travel between ranks), model heap"""
    rs_host.build_model("mixnet", emu=True)
    n, T = 301, 120
    spec = {"model": "mixnet", "n": n, "state_bytes": 32, "margs": ["--order-sensitive", "--mean", "2.0"],
            "cfg": dict(master_seed=42, simulation_time=float(T), window_events=2048, arena_bytes=1024, max_payload=16, bucket_width=0.25)}
    rows, got = run_ranks(3, spec, tmp_path)
    _, orows, _ = rs_trace.run_oracle("mixnet", n, sim_time=T, horizon=T, state_bytes=32, seed=42, extra=spec["margs"])
    assert rs_trace.compare(rows, orows) == []
    assert sum(g["stats"]["antimessages"] for g in got) > 0


def test_fatal_status_on_one_rank_stops_all_ranks(oracle_built, tmp_path):
    """A capacity overflow on ONE rank (its exchange slot: xchg_cap far too small) must end the run on EVERY rank with
    an error, not leave the others waiting in a collective: the fatal flag travels with the per-round all-reduce and
    the ranks that are fine report RS_ST_PEER_FATAL."""
    rs_host.build_model("mixnet", emu=True)
    spec = {"model": "mixnet", "n": 301, "state_bytes": 32, "margs": ["--token-prob", "0.9"], "expect_error": True,
            "cfg": dict(master_seed=42, simulation_time=120.0, window_events=2048, arena_bytes=1024, max_payload=16, bucket_width=0.25, xchg_cap=2)}
    rows, got = run_ranks(3, spec, tmp_path)
    assert all(g["error"] == 6 for g in got), [g["error"] for g in got]          # RS_ERR_CAPACITY everywhere
    names = [set(g["status_names"]) for g in got]
    assert any("XCHG_FULL" in s for s in names)
    assert all(("XCHG_FULL" in s) or ("PEER_FATAL" in s) for s in names)


@pytest.mark.parametrize("world,name,T", [(2, "traffic_grid_lp960", 2000), (3, "traffic_grid_lp16", 20000), (4, "traffic_grid_lp960", 1000)])
def test_traffic_multirank_vs_oracle(oracle_built, tmp_path, world, name, T):
    """models/traffic over several ranks: every rank loads the topology image and runs INIT for its own block of LPs
    (neighbour ids are global), cars cross ranks inside 16-byte payloads, antimessages travel between ranks."""
    if not rs_host.model_sources("traffic") or not os.path.exists(os.path.join(oracle_built, "seqp_traffic")):
        pytest.skip("traffic sources live in the reference tree")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_traffic_topology
    text, n = make_traffic_topology.topology(name)
    rs_host.build_model("traffic", emu=True)
    spec = {"model": "traffic", "n": n, "state_bytes": 8, "files": {"topology.txt": text},
            "cfg": dict(master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=1 << 14, arena_bytes=1 << 17,
                        max_payload=16, bucket_width=4.0, n_buckets=4096)}
    rows, got = run_ranks(world, spec, tmp_path)
    _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text})
    assert rs_trace.compare(rows, orows) == []
    assert sum(g["stats"]["committed_events"] for g in got) == sum(r[0] for r in orows)


@pytest.mark.parametrize("world,hot,chunk", [(2, 8, 16), (3, 32, 0)])
def test_phold_multirank_with_chain_lps(oracle_built, tmp_path, world, hot, chunk):
    """hot_threshold on several ranks: LPs with big groups become chain LPs (a warp of their own, a visit
    budget per round) on every rank, and nothing they were due to execute may be dropped at the window commit."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    rs_host.build_model("phold", emu=True)
    n, T = 300, 90
    cfg = dict(master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=1 << 20, arena_bytes=128, hot_threshold=hot,
               bucket_width=0.25)
    if chunk:
        cfg["chunk_events"] = chunk
    rows, got = run_ranks(world, {"model": "phold", "n": n, "state_bytes": 8, "cfg": cfg}, tmp_path)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=[0, 4, 5, 6, 7]) == []
    assert sum(g["stats"]["committed_events"] for g in got) == sum(r[0] for r in orows)
    assert sum(g["stats"]["chain_events"] for g in got) > 0
