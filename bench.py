#!/usr/bin/env python3
"""bench.py -- committed events/sec of the B200-native Time Warp engine on PHOLD.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lps L] [--sim-time T]

One STEP = one complete bounded-horizon PHOLD run (BASELINE.json configs[1]: models/phold,
1,048,576 LPs, exponential timestamp increment, --simulation-time T, master seed 12345678901234567):
the engine is created and INIT is run on every LP OUTSIDE the timed region (inputs resident in HBM),
the timed region is the event loop (rs_dev_run to the horizon), timed on the device with CUDA events
recorded by the engine on its own stream.  `value` = committed events of K steps / their device time.
`e2e` = the same metric through the C-ABI with HOST buffers: per-LP seeds copied host->device, engine
creation, INIT, the event loop, and the per-LP committed-event digests copied device->host, all inside
a wall-clock timed region (the checksum of those digests, which is compared with the oracle's, is computed after it).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))

MASTER_SEED = 12345678901234567
B_ALG_PHOLD = 274.0          # algorithmic bytes per committed event, SURVEY.md section 8(d) / DESIGN.md
METRIC = "committed events/sec (PHOLD 1M LPs)"


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.index = index

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def phold_cfg(n_lps, sim_time, world=1):
    """Pool sizing for shipped PHOLD (population grows ~ N*0.9*e^{0.04 t}).
    Windows are WIDE (up to 4 simulated time units, 64 calendar buckets of 1/16): inside a window the engine
    paces the LPs in rounds, so an LP that receives a same-timestamp burst from its n_lps/64 twin senders
    (src/lib/numerical.c:399-409: only 64 distinct seeds) falls behind without stopping the others, and has the
    rest of the window to catch up."""
    import math
    pending = int(n_lps * (1.0 * math.exp(0.04 * sim_time) + 1.0) * 1.25) + (1 << 20)
    rate = 0.04 * 5.0 * n_lps * math.exp(0.04 * sim_time)          # events per simulated time unit at the end of the run
    if world > 1:       # max_pending sizes ONE rank's pool: its share of the population, plus room for the uneven hot LPs
        pending = int(pending / world * 1.5) + (1 << 20)
        rate = rate / world * 1.5
    span = 4.0
    win = int(rate * span * 1.1) + (1 << 20)                     # events of the widest window (those already pending when it starts)
    return dict(n_lps=n_lps, master_seed=MASTER_SEED, simulation_time=float(sim_time), window_events=win,
                max_window=win + (win >> 2), max_out=2 * win + (1 << 22), window_span=span,
                arena_bytes=128, max_pending=pending, bucket_width=0.0625, n_buckets=4096,
                gvt_snapshot_cycles=1 << 30, flags=0)


def run_ours(args):
    import rs_host
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib_path = os.environ.get("RS_LIB") or rs_host.build_model("phold")     # (RS_LIB: a diagnostic build of the same model)
    n_lps, T = args.lps, args.sim_time
    cfg = phold_cfg(n_lps, T, world)
    cfg["device"] = local
    lib = rs_host.load(lib_path)
    if os.environ.get("RS_DEBUG") and rank == int(os.environ.get("RS_DEBUG_RANK", "0")):
        lib.rs_set_debug(int(os.environ["RS_DEBUG"]))       # window log on stderr (diagnostic runs only)
    first, count = rs_host.lp_block(lib, n_lps, world, rank)
    seeds = rs_host.lp_seeds(lib, MASTER_SEED, first, count)
    if world > 1:
        cfg.update(rank=rank, nranks=world, lp_first=first, lp_count=count)

    comm = C.c_void_p()
    if world > 1:
        # ONE NCCL communicator for all the engines of this process: id made by rank 0, broadcast through the
        # launcher's process group; creation (collective, ~1 s with channel set-up) stays outside the timed region
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            assert lib.rs_nccl_unique_id(buf, None) == 0
        t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        buf = (C.c_ubyte * 128)(*t.cpu().tolist())
        assert lib.rs_nccl_comm_create(buf, rank, world, None, C.byref(comm)) == 0

    def connect_peers(eng):
        """direct exchange over NVLink peer memory: gather the ranks' inbox handles, map them (RS_NO_IPC=1: keep the
        fixed-size NCCL send/recv exchange)"""
        if os.environ.get("RS_NO_IPC"):
            return
        mine = torch.tensor(list(eng.ipc_export()), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ok = 1
        try:
            eng.ipc_import([bytes(h.cpu().tolist()) for h in allh])
        except rs_host.RsError as ex:
            print("bench: direct exchange not available on rank %d (%s)" % (rank, ex), file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag[0]) == 0:       # every rank takes the same path
            eng.ipc_import(None)

    import rs_trace

    def one_step(e2e, cfg=cfg, seeds=seeds, first=first, count=count):
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng = rs_host.Engine(lib_path)
        # the end-to-end steps carry the committed-trace digests (RS_F_TRACE: per-LP event hash) and are checked
        # against the oracle's golden checksum of this very configuration
        eng.init(host_seeds=seeds if e2e else None, **(dict(cfg, flags=rs_host.RS_F_TRACE) if e2e else cfg))
        if world > 1:
            eng.set_nccl_comm(comm)
            connect_peers(eng)
        fin = eng.run()
        st = eng.stats()
        digests = eng.trace() if e2e else None        # D2H of the per-LP digests (32 bytes per LP): the step's result
        t1 = time.perf_counter()
        cs = rs_trace.checksum_digests(digests, first) if e2e else None   # (the parity check itself is not part of the step)
        assert fin and (st.status & ~0x1) == 0, st.as_dict()
        committed = st.committed_events - count       # INIT events are processed at engine creation
        res = dict(committed=committed, device_s=st.device_seconds, wall_s=t1 - t0, launches=st.kernel_launches,
                   windows=st.windows, rollbacks=st.rollbacks, executed=st.executed_events, silent=st.silent_events,
                   rounds=st.relax_rounds, late=st.late_events, antimessages=st.antimessages)
        eng.fini()
        if world > 1:
            v = torch.tensor([res["device_s"], res["wall_s"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            c = torch.tensor([float(committed), float(res["launches"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            res["device_s"], res["wall_s"] = float(v[0]), float(v[1])
            res["committed"], res["launches"] = float(c[0]), float(c[1])
        if e2e:
            # partial checksums of the ranks' LP blocks add up (mod 2^64) to the checksum of the whole trace
            parts = [cs]
            if world > 1:
                parts = [None] * world
                dist.all_gather_object(parts, cs)
            res["checksum"] = sum(p[0] for p in parts) & 0xFFFFFFFFFFFFFFFF
            res["traced"] = sum(p[1] for p in parts)
        return res

    def check_parity(res, n, T):
        """compare with tests/golden/phold_checksums.json (oracle run of exactly this configuration); None = no golden"""
        gp = os.path.join(REPO, "tests", "golden", "phold_checksums.json")
        gold = json.load(open(gp)).get("phold_lp%d_T%g" % (n, T)) if os.path.exists(gp) else None
        if gold is None:
            return None, None
        ok = ("%016x" % res["checksum"]) == gold["checksum"] and res["traced"] == gold["traced_events_incl_init"]
        return ok, gold

    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local)
    sampler.start()
    steps = [one_step(False) for _ in range(args.steps)]
    e2e_steps = [one_step(True) for _ in range(max(1, min(args.steps, 2)))]
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        for tag, ss in (("step", steps), ("e2e", e2e_steps)):
            for x in ss:
                print("bench: %s device_s=%.4f wall_s=%.4f windows=%d" % (tag, x["device_s"], x["wall_s"], x["windows"]), file=sys.stderr)

    # one extra, untimed-for-the-metric step with CUDA events around EVERY kernel launch (RS_F_KTIME):
    # per-kernel device time for the roofline of the dominant kernel
    kcfg = dict(cfg, flags=rs_host.RS_F_KTIME)
    eng = rs_host.Engine(lib_path)
    eng.init(**kcfg)
    if world > 1:
        eng.set_nccl_comm(comm)
        connect_peers(eng)
    eng.run()
    kst = eng.stats()
    ktimes = eng.kernel_times()
    eng.fini()
    k_committed = kst.committed_events - count
    dom = max((k for k in ktimes if k not in ("k_init", "k_init_pool")), key=lambda k: ktimes[k][0])
    dom_ms, dom_launches = ktimes[dom]
    # events the dominant kernel itself executed (k_run_chain: the chain LPs' events, silent re-executions included;
    # any other kernel: all committed events)
    dom_events = kst.chain_events if dom == "k_run_chain" else k_committed

    # supplementary: the same model and size with a DISTINCT seed per LP (host-provided seed array through
    # the same C-ABI call).  The reference derives only 64 distinct seeds (gid % 64), which makes n_lps/64
    # LPs fire in lock step and produces LPs with ~10^5 x the average load; this run shows what the engine
    # does when the workload has no such serial chain.  Not the headline, not comparable with the reference.
    sup = None
    if not args.no_supplementary and world == 1:
        def splitmix(x):
            x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            return z ^ (z >> 31)
        import numpy as np
        g = np.arange(n_lps, dtype=np.uint64)
        x = (g * np.uint64(0x9E3779B97F4A7C15) + np.uint64(MASTER_SEED)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        lo = (z & np.uint64(0xFFFFFFFF)) % np.uint64(0x9068FFFE) + np.uint64(1)
        hi = (z >> np.uint64(32)) % np.uint64(0x464FFFFE) + np.uint64(1)
        dseeds = ((hi << np.uint64(32)) | lo).astype(np.uint64)
        arr = (C.c_uint64 * n_lps).from_buffer(dseeds)
        scfg = dict(cfg)
        best = None
        for _ in range(2):
            eng = rs_host.Engine(lib_path)
            eng.init(host_seeds=arr, **scfg)
            eng.run()
            st = eng.stats()
            eng.fini()
            v = (st.committed_events - n_lps) / st.device_seconds
            if best is None or v > best["value"]:
                best = {"value": v, "unit": "events/s", "committed_events": st.committed_events - n_lps, "windows": st.windows,
                        "rollbacks": st.rollbacks, "device_seconds": st.device_seconds,
                        "hbm_frac_whole_loop": B_ALG_PHOLD * v / 1e9 / measured_peak()[0]}
        sup = dict(best, workload="models/phold, %d LPs, T=%g, one DISTINCT seed per LP (host seed array): no twin LPs" % (n_lps, T))

    # BASELINE.json configs[2] (C3): models/pcs, 65,536 cells (256 x 256 hexagons), checkpoint-bound: --p 10 and --p 1
    pcs = None
    if world == 1 and not args.no_pcs:
        pcs = pcs_block(rs_host, rs_trace)

    # SURVEY.md section 8(f) rank 3 (BASELINE config 5): models/traffic on a generated road grid
    traffic_row = None
    if world == 1 and not args.no_traffic:
        try:
            traffic_row = traffic_block(rs_host, rs_trace)
        except Exception as ex:         # a widened row never costs the headline line
            traffic_row = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:300]), "parity_checksum_ok": None}

    # BASELINE.json configs[4] (C5): the same model on several GPUs, 68,608-LP road grid, one checked run (the flow of
    # tools/multigpu_check_traffic.py: every rank loads the topology image and runs INIT for its own LPs)
    if world > 1 and not args.no_traffic:
        try:
            sys.path.insert(0, os.path.join(REPO, "tools"))
            import multigpu_check_traffic
            traffic_row = multigpu_check_traffic.run(os.environ.get("RS_BENCH_TRAFFIC_GRID", "traffic_grid_lp68608"), 3600.0,
                                                     rank, world, local, dist, torch)
        except Exception as ex:         # a widened row never costs the headline line
            traffic_row = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:300]), "parity_checksum_ok": None}

    parity_ok, gold = None, None
    for x in e2e_steps:
        ok, gold = check_parity(x, n_lps, T)
        parity_ok = ok if parity_ok is None else (parity_ok and ok)

    # BASELINE.json configs[3] (C4): PHOLD with 16,777,216 LPs on 2/4/8 GPUs, --simulation-time 60, one checked step
    secondary = None
    if world > 1 and not args.no_c4:
        n4, T4 = 1 << 24, float(args.c4_sim_time)
        cfg4 = phold_cfg(n4, T4, world)
        f4, c4 = rs_host.lp_block(lib, n4, world, rank)
        cfg4.update(device=local, rank=rank, nranks=world, lp_first=f4, lp_count=c4)
        r4 = one_step(True, cfg=cfg4, seeds=rs_host.lp_seeds(lib, MASTER_SEED, f4, c4), first=f4, count=c4)
        ok4, gold4 = check_parity(r4, n4, T4)
        secondary = {"metric": "committed events/sec (PHOLD 16M LPs, BASELINE configs[3])", "value": r4["committed"] / r4["device_s"],
                     "unit": "events/s", "lps": n4, "simulation_time": T4, "committed_events": r4["committed"], "device_s": r4["device_s"],
                     "e2e_value": r4["committed"] / r4["wall_s"], "windows": r4["windows"], "rollbacks": r4["rollbacks"],
                     "parity_checksum_ok": ok4, "checksum": "%016x" % r4["checksum"],
                     "hbm_frac_whole_loop": B_ALG_PHOLD * r4["committed"] / r4["device_s"] / 1e9 / measured_peak()[0] / world}
        if rank == 0:
            print("bench: C4 16M LPs device_s=%.3f wall_s=%.3f parity=%s" % (r4["device_s"], r4["wall_s"], ok4), file=sys.stderr)

    dev_s = sum(s["device_s"] for s in steps)          # already the max over ranks per step
    committed = sum(s["committed"] for s in steps)      # already summed over ranks
    value = committed / dev_s
    e2e_value = sum(s["committed"] for s in e2e_steps) / sum(s["wall_s"] for s in e2e_steps)
    peak, peak_src = measured_peak()
    # DRAM traffic of the dominant kernel: from the committed ncu --set full capture (never measured under this run)
    traffic, traffic_src = None, None
    tp = os.path.join(REPO, "profiles", "r02_k_run_chain_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("kernel") == dom:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    # dominant kernel: algorithmic bytes of the events it executed / its own device time
    achieved = B_ALG_PHOLD * dom_events / (dom_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "models/phold (unmodified reference sources), %d LPs, --simulation-time %g, master seed %d, "
                               "one step = one full run; inputs larger than L2 (LP table + calendar >> 126 MB)" % (n_lps, T, MASTER_SEED),
                   "lps": n_lps, "simulation_time": T, "committed_events_per_step": committed / args.steps,
                   "windows_per_step": steps[0]["windows"], "rollbacks_per_step": steps[0]["rollbacks"],
                   "executed_per_step": steps[0]["executed"], "silent_per_step": steps[0]["silent"],
                   "bucket_width": cfg["bucket_width"],
                   "partition": "LPs block-partitioned over %d GPU(s), one engine per GPU; per pacing round the ranks push the records "
                                "that are due into the peers' inboxes over NVLink peer memory (one kernel) and agree by NCCL all-reduce" % world},
        "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": 8 * n_lps + C.sizeof(rs_host.Config),
                "d2h_bytes_per_step": 32 * n_lps + C.sizeof(rs_host.Stats)},
        "gpu_launches": int(sum(s["launches"] for s in steps)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src, "kernel": dom, "kernel_ms_per_launch": dom_ms / max(1, dom_launches),
                     "kernel_launches_per_step": dom_launches, "kernel_share_of_step": dom_ms * 1e-3 / kst.device_seconds,
                     "kernel_events_per_step": dom_events,
                     "algorithmic_bytes_per_event": B_ALG_PHOLD, "whole_loop_achieved": B_ALG_PHOLD * value / 1e9,
                     "whole_loop_frac": B_ALG_PHOLD * value / 1e9 / peak,
                     "note": "latency-bound, not bandwidth-bound: k_run_chain executes the busy LPs (a same-timestamp burst of up to "
                             "n_lps/64 events makes its receiver an LP with thousands of pending event lineages: the twin-seed "
                             "aliasing of the reference workload), ONE lane per LP at ~1.5 us per event, about a thousand such "
                             "lanes per launch; a pacing round lasts as long as its longest visit; see DESIGN.md section 5"},
        "kernel_ms": {k: round(v[0], 3) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])},
        "clocks": sampler.summary(),
        "parity_checksum_ok": parity_ok,
        "parity": {"checksum": "%016x" % e2e_steps[0]["checksum"], "golden": gold["checksum"] if gold else None,
                   "traced_events": e2e_steps[0]["traced"], "golden_traced_events": gold["traced_events_incl_init"] if gold else None,
                   "what": "order-independent 64-bit checksum over every LP's (gid, committed event count, FNV hash of its "
                           "(timestamp bits, type) sequence, last timestamp bits, RNG seed), computed from the device digests of "
                           "the end-to-end steps and compared with the CPU oracle's checksum for this configuration "
                           "(tests/golden/phold_checksums.json, tests/golden/make_checksums.py)"},
    }
    if sup:
        line["supplementary"] = sup
    if pcs:
        line["pcs"] = pcs
    if traffic_row:
        line["traffic_model"] = traffic_row
    if secondary:
        line["secondary"] = secondary
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(1)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_ok is False or (secondary and secondary["parity_checksum_ok"] is False) or (pcs and pcs["parity_checksum_ok"] is False) \
            or (traffic_row and traffic_row["parity_checksum_ok"] is False):
        print("bench.py: PARITY CHECKSUM MISMATCH against the oracle", file=sys.stderr)
        sys.exit(3)


def pcs_block(rs_host, rs_trace):
    """models/pcs (unmodified reference sources) on one GPU: 65,536 cells, --simulation-time 400, with the reference's
    default checkpoint period (--p 10) and with a checkpoint after every event (--p 1).  Every run is checked against
    the oracle's golden checksum (tests/golden/phold_checksums.json, key pcs_lp65536_T400).  B_alg follows SURVEY.md
    section 8(d): E + f*E + 2*S + S/P + f*H with E = 32-byte record + 48-byte payload slot, f = records sent per
    committed event (measured), S = measured average checkpoint size (model heap + 56-byte control block), H = 16."""
    lib_pcs = rs_host.build_model("pcs")
    n, T = 65536, 400.0
    gp = os.path.join(REPO, "tests", "golden", "phold_checksums.json")
    gold = json.load(open(gp)).get("pcs_lp%d_T%g" % (n, T)) if os.path.exists(gp) else None
    peak = measured_peak()[0]
    out = {"workload": "models/pcs (unmodified reference sources), %d cells, --simulation-time %g, master seed %d" % (n, T, MASTER_SEED),
           "unit": "events/s", "parity_checksum_ok": None}
    for p in (10, 1):
        best = None
        for _ in range(2):
            eng = rs_host.Engine(lib_pcs)
            eng.init(n_lps=n, master_seed=MASTER_SEED, simulation_time=T, window_events=1 << 20, arena_bytes=32768, max_payload=48,
                     bucket_width=0.25, n_buckets=4096, max_pending=256 * n + (1 << 20), gvt_snapshot_cycles=1 << 30,
                     ckpt_period=p, flags=rs_host.RS_F_TRACE)
            fin = eng.run()
            st = eng.stats()
            cs, traced = rs_trace.checksum_engine(eng, 0)
            eng.fini()
            assert fin and (st.status & ~0x1) == 0, st.as_dict()
            if gold is not None:
                ok = ("%016x" % cs) == gold["checksum"] and traced == gold["traced_events_incl_init"]
                out["parity_checksum_ok"] = ok if out["parity_checksum_ok"] is None else (out["parity_checksum_ok"] and ok)
            committed = st.committed_events - n
            v = committed / st.device_seconds
            if best is None or v > best["value"]:
                S = st.checkpoint_bytes / max(1, st.checkpoints)
                f = st.sent_records / max(1, committed)
                b_alg = 80.0 + f * 80.0 + 2.0 * S + S / p + f * 16.0
                best = {"value": v, "committed_events": committed, "device_s": st.device_seconds, "windows": st.windows, "rounds": st.relax_rounds,
                        "rollbacks": st.rollbacks, "silent_events": st.silent_events, "checkpoints": st.checkpoints,
                        "avg_checkpoint_bytes": S, "sends_per_event": f, "algorithmic_bytes_per_event": b_alg,
                        "hbm_frac_whole_loop": b_alg * v / 1e9 / peak, "window_cuts": st.as_dict()["window_cuts"]}
        out["p%d" % p] = best
    out["value"] = out["p10"]["value"]
    exe = os.path.join(REPO, "oracle", "_ref", "pcs")
    if os.path.exists(exe):
        cores = max(1, min(os.cpu_count() or 1, 32))
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, ".rootsim"))
            open(os.path.join(td, ".rootsim", "numerical.conf"), "w").write("%d\n" % MASTER_SEED)
            try:
                subprocess.run([exe, "--wt", str(cores), "--lp", "1024", "--simulation-time", "400", "--deterministic-seed",
                                "--output-dir", os.path.join(td, "o")], env=dict(os.environ, HOME=td), cwd=td, capture_output=True, text=True, timeout=300)
                txt = open(os.path.join(td, "o", "execution_stats")).read()
                committed = float([l for l in txt.split("\n") if l.startswith("TOTAL COMMITTED EVENTS")][0].split(":")[1])
                secs = float([l for l in txt.split("\n") if l.startswith("TOTAL SIMULATION TIME")][0].split(":")[1].split()[0])
                out["cpu_baseline"] = {"value": committed / secs, "unit": "events/s", "cores": cores, "kind": "reference",
                                       "sample": "unmodified ROOT-Sim 2.0.0 (oracle/_ref/pcs) --wt %d --lp 1024 --simulation-time 400" % cores}
            except Exception as ex:     # the baseline is a report, never a reason to lose the bench line
                out["cpu_baseline"] = {"unavailable": str(ex)[:200]}
    return out


def traffic_block(rs_host, rs_trace):
    """models/traffic (unmodified reference sources) on one GPU: the generated 16 x 16 road grid (256 junctions + 3840
    road segments = 4096 LPs, tests/golden/make_traffic_topology.py), one simulated hour.  INIT runs on the device and
    parses the topology image the host loaded.  Checked against the oracle's golden checksum (key traffic_lp4096_T3600);
    the unmodified reference with all host threads runs beside it.  The run is bounded by the event chain of the
    busiest junction (cars enter at junctions only), not by bandwidth: 4096 LPs cannot fill 148 SMs."""
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    import make_traffic_topology as mt
    name, T = "traffic_grid_lp4096", 3600.0
    text, n = mt.topology(name)
    lib = rs_host.build_model("traffic")
    gp = os.path.join(REPO, "tests", "golden", "phold_checksums.json")
    gold = json.load(open(gp)).get("traffic_lp%d_T%g" % (n, T)) if os.path.exists(gp) else None
    out = {"workload": "models/traffic (unmodified reference sources), generated road grid %s (%d LPs), --simulation-time %g, master seed %d"
                       % (name, n, T, MASTER_SEED), "unit": "events/s", "parity_checksum_ok": None}
    best = None
    for _ in range(2):
        eng = rs_host.Engine(lib)
        eng.model_file("topology.txt", text.encode())
        t0 = time.time()
        eng.init(n_lps=n, master_seed=MASTER_SEED, simulation_time=T, window_events=1 << 18, arena_bytes=1 << 17, max_payload=16,
                 bucket_width=4.0, n_buckets=4096, flags=rs_host.RS_F_TRACE)
        t_init = time.time() - t0
        fin = False
        while not fin and time.time() - t0 < 120.0:      # bounded: this block must never hold up the bench
            fin = eng.run(max_windows=2)
        st = eng.stats()
        cs, traced = rs_trace.checksum_engine(eng, 0)
        eng.fini()
        if not fin:
            out["unfinished"] = "not finished after 120 s (gvt %g of %g)" % (st.gvt, T)
            break
        assert (st.status & ~0x1) == 0, st.as_dict()
        if gold is not None:
            ok = ("%016x" % cs) == gold["checksum"] and traced == gold["traced_events_incl_init"]
            out["parity_checksum_ok"] = ok if out["parity_checksum_ok"] is None else (out["parity_checksum_ok"] and ok)
        committed = st.committed_events - n
        v = committed / st.device_seconds
        if best is None or v > best["value"]:
            # B_alg = E + f*E + 2*S + S/P + f*H, E = 32-byte record + 16-byte payload slot, f = records
            # sent per committed event (measured), S = measured average checkpoint size, P = 10 (default --p), H = 16
            S = st.checkpoint_bytes / max(1, st.checkpoints)
            f = st.sent_records / max(1, committed)
            b_alg = 48.0 + f * 48.0 + 2.0 * S + S / 10.0 + f * 16.0
            best = {"value": v, "committed_events": committed, "device_s": st.device_seconds, "init_s_incl_file_parse_on_device": t_init,
                    "sends_per_event": f, "algorithmic_bytes_per_event": b_alg, "hbm_frac_whole_loop": b_alg * v / 1e9 / measured_peak()[0],
                    "e2e_value": committed / (time.time() - t0), "windows": st.windows, "rounds": st.relax_rounds,
                    "rollbacks": st.rollbacks, "silent_events": st.silent_events, "checkpoints": st.checkpoints,
                    "avg_checkpoint_bytes": st.checkpoint_bytes / max(1, st.checkpoints), "kernel_launches": st.kernel_launches}
    if best:
        out.update(best)
    exe = os.path.join(REPO, "oracle", "_ref", "traffic")
    if os.path.exists(exe):
        cores = max(1, min(os.cpu_count() or 1, 32))
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, ".rootsim"))
            open(os.path.join(td, ".rootsim", "numerical.conf"), "w").write("%d\n" % MASTER_SEED)
            open(os.path.join(td, "topology.txt"), "w").write(text)
            try:
                subprocess.run([exe, "--wt", str(cores), "--lp", str(n), "--simulation-time", "%d" % T, "--deterministic-seed",
                                "--output-dir", os.path.join(td, "o")], env=dict(os.environ, HOME=td), cwd=td, capture_output=True, text=True, timeout=300)
                txt = open(os.path.join(td, "o", "execution_stats")).read()
                committed = float([l for l in txt.split("\n") if l.startswith("TOTAL COMMITTED EVENTS")][0].split(":")[1])
                secs = float([l for l in txt.split("\n") if l.startswith("TOTAL SIMULATION TIME")][0].split(":")[1].split()[0])
                out["cpu_baseline"] = {"value": committed / secs, "unit": "events/s", "cores": cores, "kind": "reference",
                                       "sample": "unmodified ROOT-Sim 2.0.0 (oracle/_ref/traffic) --wt %d --lp %d --simulation-time %d, same topology" % (cores, n, T)}
            except Exception as ex:     # the baseline is a report, never a reason to lose the bench line
                out["cpu_baseline"] = {"unavailable": str(ex)[:200]}
    return out


def cpu_baseline(runs):
    """The reference's own optimistic engine (--wt <host cores>) on a bounded sample of the workload:
    1024 of the LPs, --simulation-time 150 (the reference cannot hold more than 250000 LPs and needs
    4 MiB of stack per LP, SURVEY.md top)."""
    exe = os.path.join(REPO, "oracle", "_ref", "phold")
    cores = os.cpu_count() or 1
    if os.path.exists(exe):
        best = None
        for _ in range(runs):
            with tempfile.TemporaryDirectory() as td:
                os.makedirs(os.path.join(td, ".rootsim"))
                open(os.path.join(td, ".rootsim", "numerical.conf"), "w").write("%d\n" % MASTER_SEED)
                env = dict(os.environ, HOME=td)
                wt = max(1, min(cores, 32))
                subprocess.run([exe, "--wt", str(wt), "--lp", "1024", "--simulation-time", "150", "--deterministic-seed",
                                "--output-dir", os.path.join(td, "o")], env=env, cwd=td, capture_output=True, text=True, timeout=900)
                txt = open(os.path.join(td, "o", "execution_stats")).read()
                committed = float([l for l in txt.split("\n") if l.startswith("TOTAL COMMITTED EVENTS")][0].split(":")[1])
                secs = float([l for l in txt.split("\n") if l.startswith("TOTAL SIMULATION TIME")][0].split(":")[1].split()[0])
                v = committed / secs
                best = v if best is None else max(best, v)
        out = {"value": best, "unit": "events/s", "cores": wt, "kind": "reference",
               "sample": "unmodified ROOT-Sim 2.0.0 (oracle/_ref/phold) --wt %d --lp 1024 --simulation-time 150; the reference "
                         "cannot run 1M LPs (MAX_LPs 250000, 4 MiB stack per LP)" % wt}
        # the reference's own sequential engine on the same sample (one core).  Larger LP counts are
        # left out: `--wt 8 --lp 4096 --simulation-time 100` does not finish in two minutes (O(#LP) LP lookup per event,
        # src/scheduler/process.c:157-164; 16 GB of stacks)
        try:
            with tempfile.TemporaryDirectory() as td:
                os.makedirs(os.path.join(td, ".rootsim"))
                open(os.path.join(td, ".rootsim", "numerical.conf"), "w").write("%d\n" % MASTER_SEED)
                subprocess.run([exe, "--sequential", "--lp", "1024", "--simulation-time", "150", "--deterministic-seed",
                                "--output-dir", os.path.join(td, "o")], env=dict(os.environ, HOME=td), cwd=td, capture_output=True, text=True, timeout=120)
                txt = open(os.path.join(td, "o", "sequential_stats")).read()
                events = float([l for l in txt.split("\n") if l.startswith("TOTAL EXECUTED EVENTS")][0].split(":")[1])
                secs = float([l for l in txt.split("\n") if l.startswith("TOTAL SIMULATION TIME")][0].split(":")[1].split()[0])
                out["sequential"] = {"value": events / secs, "unit": "events/s", "cores": 1,
                                     "sample": "unmodified ROOT-Sim 2.0.0 --sequential --lp 1024 --simulation-time 150"}
        except Exception as ex:
            out["sequential"] = {"unavailable": str(ex)[:200]}
        return out
    import rs_trace
    _, _, out = rs_trace.run_oracle("phold", 65536, sim_time=60, horizon=60)
    v = float(out.strip().split("events_per_s=")[1].split()[0])
    return {"value": v, "unit": "events/s", "cores": 1, "kind": "port",
            "sample": "oracle/seqsim.c (sequential restatement) phold --lp 65536 --simulation-time 60"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    for _ in range(args.warmup):
        cpu_baseline(1)
    vals = [cpu_baseline(1) for _ in range(args.steps)]
    t1 = time.perf_counter()
    v = sum(x["value"] for x in vals) / len(vals)
    cb = dict(vals[0], value=v)
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "events/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (t1 - t0) / max(1, args.steps + args.warmup),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "models/phold on the host CPU: " + cb["sample"]},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else the process prints (the model's own printf
    from INIT runs on the device and lands on fd 1) is diverted to stderr for the duration of the run."""
    sys.stdout.flush()
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lps", type=int, default=1 << 20)
    ap.add_argument("--sim-time", type=float, default=100.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-supplementary", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the 16M-LP secondary run of the multi-GPU bench")
    ap.add_argument("--c4-sim-time", type=float, default=60.0, help="horizon of the 16M-LP secondary run (golden checksums exist for 30 and 60)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the models/traffic block (SURVEY 8(f) rank 3) of the 1-GPU bench")
    ap.add_argument("--no-pcs", action="store_true", help="skip the models/pcs block (BASELINE configs[2]) of the 1-GPU bench")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
