"""Engine LOGIC on the CPU: the same unity source that nvcc compiles for sm_100a is built with g++ in
thread-emulation mode (rootsim-cc --emu; kernels become serial loops) and its committed trace is
compared BIT-EXACTLY (timestamps included) with the CPU oracle built with the same portable log.
This is a test harness for the window/rollback/cancellation logic in a container without a GPU; it is
not a product path (the shipped library has no CPU execution path, see test_abi.py)."""
import os

import pytest

import rs_host
import rs_trace

PHOLD_MASK = [0, 4, 5, 6, 7]


def run_emu(model, n, T, wev, margs=(), state_bytes=8, seed=rs_trace.MASTER_SEED, **cfg):
    lib = rs_host.build_model(model, emu=True)
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(list(margs))
        eng.init(n_lps=n, master_seed=seed, simulation_time=T, window_events=wev, **cfg)
        assert eng.run()
        st = eng.stats()
        rows = rs_trace.engine_rows(eng, state_bytes)
    return rows, st


@pytest.mark.parametrize("n,T,wev", [(64, 100, 100000), (1024, 60, 4096), (257, 80, 50), (4096, 40, 100000), (1, 50, 100), (2, 50, 100), (3, 150, 100)])
def test_phold_emu_vs_oracle(oracle_built, n, T, wev):
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    rows, st = run_emu("phold", n, T, wev, arena_bytes=128)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.committed_events == sum(r[0] for r in orows)
    assert st.status & ~0x1 == 0


@pytest.mark.parametrize("n,T,wev,margs", [
    (48, 200, 100000, ["--order-sensitive"]),
    (63, 300, 50, ["--order-sensitive", "--mean", "1.5"]),
    (1000, 100, 100000, []),
    (300, 150, 2000, ["--token-prob", "0.9"]),
])
def test_mixnet_emu_vs_oracle(oracle_built, n, T, wev, margs):
    rows, st = run_emu("mixnet", n, T, wev, margs, state_bytes=32, seed=42, arena_bytes=1024, max_payload=16, bucket_width=0.25)
    _, orows, _ = rs_trace.run_oracle("mixnet", n, sim_time=T, horizon=T, state_bytes=32, seed=42, extra=margs)
    assert rs_trace.compare(rows, orows) == []
    assert st.rollbacks > 0          # the cases are sized so that the optimistic machinery is exercised


@pytest.mark.parametrize("n,T,wev", [(16, 600, 4096), (64, 300, 256), (256, 120, 100000)])
def test_pcs_emu_vs_oracle(oracle_built, n, T, wev):
    """models/pcs unchanged: 40-byte payloads, malloc/free inside events (channel lists), hexagonal
    topology, zero-delay HANDOFF_RECV between neighbours.  First 56 state bytes = counters, lvt, ta."""
    if not rs_host.model_sources("pcs") or not os.path.exists(os.path.join(oracle_built, "seqp_pcs")):
        pytest.skip("pcs sources live in the reference tree")
    rows, st = run_emu("pcs", n, T, wev, state_bytes=56, arena_bytes=32768, max_payload=48, bucket_width=0.25, n_buckets=4096)
    _, orows, _ = rs_trace.run_oracle("pcs", n, sim_time=T, horizon=T, state_bytes=56)
    assert rs_trace.compare(rows, orows) == []
    assert st.rollbacks > 0 and st.antimessages > 0


def test_pcs_mutual_zero_delay_cycle(oracle_built):
    """Regression: at 256 x 256 cells, cell 33 and its neighbour 289 share a seed (gid % 64) and hand a call
    off to EACH OTHER at the bit-identical timestamp 29.087...; with eager cancellation the two LPs roll each
    other back forever.  The receiver-side lazy cancellation makes the identical re-sent copy a no-op."""
    if not rs_host.model_sources("pcs") or not os.path.exists(os.path.join(oracle_built, "seqp_pcs")):
        pytest.skip("pcs sources live in the reference tree")
    n, T = 65536, 30
    rows, st = run_emu("pcs", n, T, 1 << 20, state_bytes=56, arena_bytes=8192, max_payload=48, bucket_width=0.25, n_buckets=4096,
                       max_pending=16 * n)
    _, orows, _ = rs_trace.run_oracle("pcs", n, sim_time=T, horizon=T, state_bytes=56)
    assert rs_trace.compare(rows, orows) == []


def test_hot_lp_path(oracle_built):
    """Twin LPs (same seed every 64 gids, src/lib/numerical.c:399-409) fire together and create LPs with
    hundreds of events per window: exercises the chain-LP path (own warp, one pacing step behind), the
    cooperative sort of large groups, the heap for in-window self events and long in-window arrival lists."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    n, T = 16384, 30
    rows, st = run_emu("phold", n, T, 1 << 20, arena_bytes=128, hot_threshold=40)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    d = st.as_dict()
    assert d["big_groups"] > 0 and d["spill_rounds"] > 0


def test_own_event_queue_overflow_is_replayed_identically(oracle_built):
    """A burst of in-window ARRIVALS (4096 twin senders, small group) overflows the LP's queue of own
    in-window events, so some of them travel through the output buffer; a rollback into that stretch
    restores a checkpoint record and replays silently -- the replay has to route every own event exactly
    as the first execution did (regression: a larger queue after the restore executed one event twice)."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    n, T = 262144, 16
    rows, st = run_emu("phold", n, T, 1 << 20, arena_bytes=128, max_pending=4 << 20)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.rollbacks > 0 and st.as_dict()["spill_rounds"] > 0


def test_window_size_does_not_change_the_committed_trace(oracle_built):
    a, sa = run_emu("mixnet", 200, 120, 64, state_bytes=32, seed=7, arena_bytes=1024, max_payload=16)
    b, sb = run_emu("mixnet", 200, 120, 1 << 20, state_bytes=32, seed=7, arena_bytes=1024, max_payload=16, bucket_width=1.0)
    assert rs_trace.compare(a, b) == []
    assert sa.windows != sb.windows


def test_model_termination_by_ongvt(oracle_built):
    """Without --simulation-time the run ends once every LP's OnGVT holds on committed state
    (src/gvt/ccgs.c:107-179); the sequential reference stops at the first instant where all agree, the
    optimistic engine at the first window boundary after it => at least as many events per LP."""
    lib = rs_host.build_model("mixnet", emu=True)
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--tick-limit", "30"])
        eng.init(n_lps=20, master_seed=5, simulation_time=0.0, window_events=500, arena_bytes=1024, max_payload=16)
        assert eng.run()
        st = eng.stats()
        states = eng.lp_state(4)
    ticks = [int.from_bytes(states[4 * i:4 * i + 4], "little") for i in range(20)]
    assert st.all_lps_agree == 1 and min(ticks) >= 30


def test_pool_exhaustion_is_reported_not_fatal_to_the_process(oracle_built):
    """Capacity errors surface as RS_ERR_CAPACITY with the status bit set (no memory corruption, no hang)."""
    lib = rs_host.build_model("mixnet", emu=True)
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--token-prob", "1.0", "--mean", "0.5"])
        eng.init(n_lps=3000, master_seed=3, simulation_time=400.0, window_events=4096, arena_bytes=1024, max_payload=16,
                 bucket_width=4.0, n_buckets=64, max_pending=512)
        with pytest.raises(rs_host.RsError) as ei:
            eng.run()
        assert ei.value.code == 6
        assert set(eng.stats().as_dict()["status_names"]) & {"POOL_FULL", "OUT_FULL", "WINDOW_FULL"}


# ---------------------------------------------------------------- error paths and degenerate inputs
def run_edge(fault, n=16, T=30.0, build=None, **cfg):
    lib = (build or (lambda: rs_host.build_model("edgecases", emu=True)))()
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--fault", str(fault)])
        eng.init(n_lps=n, master_seed=9, simulation_time=T, window_events=256, arena_bytes=512, **cfg)
        try:
            fin = eng.run()
            err = None
        except rs_host.RsError as ex:
            fin, err = None, ex
        return fin, err, eng.stats().as_dict(), rs_trace.engine_rows(eng, 8)


EDGE = [
    # fault, expected error code (None = completes), expected status flag
    (0, None, None),                      # healthy run
    (1, 5, "PAST_EVENT"),                 # src/communication/communication.c:285-289: fatal
    (2, 5, "RESERVED_TYPE"),              # :291-294: fatal
    (3, None, "BAD_RECEIVER"),            # :280-283: warning, event dropped, run goes on
    (4, 5, "MODEL_EXIT"),                 # model called exit()
    (5, 6, "ARENA_FULL"),                 # per-LP heap exhausted: malloc returns NULL, capacity error reported
    (6, None, "NO_EVENTS"),               # nothing was ever scheduled: the run ends with an empty calendar
    (7, None, None),                      # one active LP, fifteen that only ever saw INIT (ragged input)
]


@pytest.mark.parametrize("fault,code,flag", EDGE, ids=["ok", "past", "reserved-type", "bad-receiver", "exit", "arena", "no-events", "one-active"])
def test_edge_cases_emu(oracle_built, fault, code, flag):
    fin, err, st, rows = run_edge(fault)
    if code is None:
        assert err is None and fin
    else:
        assert err is not None and err.code == code
    if flag:
        assert flag in st["status_names"]
    if fault in (0, 3, 7):
        extra = ["--fault", str(fault)]
        _, orows, _ = rs_trace.run_oracle("edgecases", 16, sim_time=30.0, horizon=30.0, seed=9, extra=extra)
        assert rs_trace.compare(rows, orows) == []
    if fault == 6:
        assert st["committed_events"] == 16 and [r[0] for r in rows] == [1] * 16      # INIT only
    if fault == 7:
        assert [r[0] for r in rows][1:] == [1] * 15 and rows[0][0] > 10


def test_cli_executable_emu(oracle_built, tmp_path):
    """host/rs_main.c: ROOT-Sim command line (--lp, --simulation-time, --seed, model options chained through
    argp like src/core/init.c:385-390), execution_stats with the reference's labels, trace dump."""
    import subprocess
    import sys
    exe = tmp_path / "mixnet_cli"
    srcs = rs_host.model_sources("mixnet")
    subprocess.run([sys.executable, rs_host.ROOTSIM_CC, "--emu"] + srcs + ["-o", str(exe)], check=True, capture_output=True)
    out = tmp_path / "out"
    tr = tmp_path / "trace.txt"
    r = subprocess.run([str(exe), "--lp", "200", "--simulation-time", "60", "--seed", "42", "--wt", "4", "--no-core-binding",
                        "--b200-arena-bytes", "1024", "--b200-max-payload", "16", "--b200-bucket-width", "0.25",
                        "--output-dir", str(out), "--b200-trace", str(tr), "--mean", "3.0"], capture_output=True, text=True)
    assert r.returncode == 0 and "SIMULATION CORRECTLY COMPLETED" in r.stdout
    gvt = open(out / "thread_0_0" / "gvt").read().split("\n")          # statistics.c:408-431
    assert gvt[0].split() == ['#', '"WCT"', '"GVT', 'VALUE"', '"EVENTS"', '"CUMUL', 'EVENTS"'] or '"GVT VALUE"' in gvt[0]
    rows_g = [l.split() for l in gvt[1:] if l.strip()]
    assert rows_g and all(len(r) == 4 for r in rows_g) and int(rows_g[-1][3]) == sum(int(r[2]) for r in rows_g)
    stats = open(out / "execution_stats").read()
    for label in ("TOTAL EXECUTED EVENTS", "TOTAL COMMITTED EVENTS", "TOTAL ROLLBACKS EXECUTED", "TOTAL ANTIMESSAGES", "EFFICIENCY", "LAST COMMITTED GVT"):
        assert label in stats
    local = open(out / "thread_0_0" / "local_stats").read()                # statistics.c:518-521
    assert "THREAD STATISTICS" in local and "LPs HOSTED BY THREAD ...... : 200" in local and "TOTAL COMMITTED EVENTS" in local
    _, rows = rs_trace.parse_digest(str(tr))
    _, orows, _ = rs_trace.run_oracle("mixnet", 200, sim_time=60, horizon=60, seed=42, extra=["--mean", "3.0"])
    assert rs_trace.compare(rows, orows) == []
    lps = [l.split() for l in open(out / "thread_0_0" / "lps").read().split("\n") if l.strip() and not l.startswith("#")]   # statistics.c:483-505
    assert len(lps) == 200 and all(len(r) == 11 for r in lps)
    assert [int(r[3]) for r in lps] == [r[0] for r in orows]              # committed events per LP
    # the sequential engine is the reference's own: asking this runtime for it is an error, like an unknown option
    r = subprocess.run([str(exe), "--lp", "8", "--sequential"], capture_output=True, text=True)
    assert r.returncode != 0


# ---------------------------------------------------------------- checkpoint ring and chunked execution
@pytest.mark.parametrize("chunk,period", [(1, 1), (2, 3), (5, 2), (3, 1000000)])
def test_chunked_execution_and_checkpoint_ring(oracle_built, chunk, period):
    """Tiny chunk_events / ckpt_period force every LP to stop after a few events per visit and to take in-window
    checkpoints constantly: going on from the window-local position, rollback to a checkpoint record, rollback of
    an LP that stopped in the middle, the own-pending-events snapshot in the records, and annihilation of
    self-addressed events at rollback are all exercised on three models; the committed trace must not change."""
    cases = [("mixnet", 300, 150, ["--token-prob", "0.9"], 32, dict(seed=42),
              dict(master_seed=42, window_events=2000, arena_bytes=1024, max_payload=16, bucket_width=0.25), None)]
    if rs_host.model_sources("phold") and os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        cases.append(("phold", 1024, 60, [], 8, {}, dict(master_seed=rs_trace.MASTER_SEED, window_events=1 << 20, arena_bytes=128, run_ahead=2), PHOLD_MASK))
        cases.append(("pcs", 64, 300, [], 56, {}, dict(master_seed=rs_trace.MASTER_SEED, window_events=256, arena_bytes=32768, max_payload=48,
                                                      bucket_width=0.25, n_buckets=4096), None))
    for model, n, T, margs, sb, okw, cfg, mask in cases:
        lib = rs_host.build_model(model, emu=True)
        with rs_host.Engine(lib, private=True) as eng:
            eng.model_args(margs)
            eng.init(n_lps=n, simulation_time=float(T), chunk_events=chunk, ckpt_period=period, **cfg)
            assert eng.run()
            st = eng.stats()
            rows = rs_trace.engine_rows(eng, sb)
        _, orows, _ = rs_trace.run_oracle(model, n, sim_time=T, horizon=T, state_bytes=sb, extra=margs, **okw)
        assert rs_trace.compare(rows, orows, state_mask=mask) == [], (model, chunk, period)
        assert st.rollbacks > 0


@pytest.mark.parametrize("substeps,bw,trail,hot", [(4, 0.0, 1, 0), (8, 0.125, 2, 40), (3, 0.5, 1, 8), (64, 0.0, 1, 100), (1, 0.0, 1, 0)])
def test_paced_rounds(oracle_built, substeps, bw, trail, hot):
    """Pacing inside a window: the time limit moves from the window start to its end in `substeps` rounds, every LP
    stops in front of the first event beyond it and goes on in a later round from its window-local position; chain
    LPs (hot_threshold) stay one step behind (pace_trail).  A straggler below the position an LP stopped at rolls it
    back; long in-window arrival lists are kept ordered across visits and only the new arrivals are merged in.  The
    committed trace must not depend on any of it."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    n, T = 65536, 30
    cfg = dict(arena_bytes=128, substeps=substeps, ckpt_period=3, pace_trail=trail, hot_threshold=hot)
    if bw:
        cfg["bucket_width"] = bw
    rows, st = run_emu("phold", n, T, 1 << 20, **cfg)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.rollbacks > 0
    if bw == 0.0 and substeps > 1:
        assert st.as_dict()["run_ahead_lps"] > 0    # 1024 LPs per seed: same-timestamp bursts of up to 1024 events ran ahead of the limit
    if substeps > 1:
        base_rows, base = run_emu("phold", n, T, 1 << 20, arena_bytes=128, substeps=1, **({"bucket_width": bw} if bw else {}))
        assert rs_trace.compare(rows, base_rows, state_mask=PHOLD_MASK) == []


def test_wide_window_with_lagging_lps(oracle_built):
    """One window spanning the whole run (16 simulated time units, 64 buckets of 0.25), paced in many rounds, tiny
    event budget per visit: the LPs that receive the twin bursts fall far behind the limit while the others go on;
    what they send later arrives as stragglers everywhere."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    n, T = 32768, 16
    rows, st = run_emu("phold", n, T, 1 << 22, arena_bytes=128, bucket_width=0.25, chunk_events=16, hot_threshold=32, max_pending=4 << 20)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.windows <= 3 and st.rollbacks > 0 and st.relax_rounds > 32


# ---------------------------------------------------------------- windows cut short by resource pressure
@pytest.mark.parametrize("model,n,T,cfg,margs,sb,okw", [
    ("phold", 4096, 60, dict(window_events=1 << 16, max_out=1 << 15, arena_bytes=128), [], 8, {}),
    ("mixnet", 300, 150, dict(window_events=4000, max_out=6000, arena_bytes=1024, max_payload=16, bucket_width=0.25, master_seed=42),
     ["--token-prob", "0.9"], 32, dict(seed=42)),
])
def test_window_cut_by_resource_pressure(oracle_built, model, n, T, cfg, margs, sb, okw):
    """The output buffer is far too small for what a whole window sends: when it is half full the window is ended at
    the current pacing limit (nothing at or beyond the limit has been executed or delivered; records held back beyond it
    become future events; own events left in the LPs' queues beyond the cut are flushed to the calendar).  The committed
    trace must not change."""
    if model == "phold" and (not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold"))):
        pytest.skip("phold sources live in the reference tree")
    lib = rs_host.build_model(model, emu=True)
    kw = dict(cfg)
    kw.setdefault("master_seed", rs_trace.MASTER_SEED)
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(list(margs))
        eng.init(n_lps=n, simulation_time=float(T), **kw)
        assert eng.run()
        st = eng.stats()
        rows = rs_trace.engine_rows(eng, sb)
    _, orows, _ = rs_trace.run_oracle(model, n, sim_time=T, horizon=T, state_bytes=sb, extra=margs, **okw)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK if model == "phold" else None) == []
    assert st.as_dict()["window_cuts"] > 0
    assert st.status & ~0x1 == 0


def test_two_live_engines_of_one_library(oracle_built):
    """Two engines created from ONE loaded library and run in alternation (a few windows each): the model API reaches
    the engine through one per-library descriptor, which every rs_dev_run claims for its own engine first
    (a second live engine silently redirected the first one's sends)."""
    if not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold")):
        pytest.skip("phold sources live in the reference tree")
    lib = rs_host.build_model("phold", emu=True)
    a, b = rs_host.Engine(lib), rs_host.Engine(lib)
    assert a.lib is b.lib or a.lib._handle == b.lib._handle
    a.init(n_lps=96, master_seed=rs_trace.MASTER_SEED, simulation_time=60.0, window_events=200, arena_bytes=128)
    b.init(n_lps=200, master_seed=rs_trace.MASTER_SEED, simulation_time=40.0, window_events=300, arena_bytes=128)
    fa = fb = False
    for _ in range(10000):
        if not fa:
            fa = a.run(max_windows=2)
        if not fb:
            fb = b.run(max_windows=3)
        if fa and fb:
            break
    assert fa and fb
    for eng, n, T in ((a, 96, 60), (b, 200, 40)):
        rows = rs_trace.engine_rows(eng, 8)
        _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
        assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
        assert eng.stats().status & ~0x1 == 0
    a.fini()
    b.fini()


@pytest.mark.parametrize("model,n,T,nb,bw,margs", [
    ("phold", 257, 80, 64, 0.0625, []),          # ring = 4 time units, mean increment 5: most events start on the far list
    ("phold", 1024, 40, 128, 0.03125, []),
    ("mixnet", 200, 120, 64, 0.25, ["--mean", "30.0"]),     # payloads travel through the far list too
])
def test_far_future_events_wait_on_the_far_list(oracle_built, model, n, T, nb, bw, margs):
    """An event beyond the calendar ring's horizon (n_buckets x bucket_width) is not an error: it waits on the far
    list and is re-filed when the ring has advanced (the reference's list_insert takes any timestamp,
    src/datatypes/list.h:163-208)."""
    if model == "phold" and (not rs_host.model_sources("phold") or not os.path.exists(os.path.join(oracle_built, "seqp_phold"))):
        pytest.skip("phold sources live in the reference tree")
    if model == "phold":
        rows, st = run_emu("phold", n, T, 4096, arena_bytes=128, n_buckets=nb, bucket_width=bw)
        _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
        assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    else:
        rows, st = run_emu("mixnet", n, T, 4096, margs, state_bytes=32, seed=42, arena_bytes=1024, max_payload=16, n_buckets=nb, bucket_width=bw)
        _, orows, _ = rs_trace.run_oracle("mixnet", n, sim_time=T, horizon=T, state_bytes=32, seed=42, extra=margs)
        assert rs_trace.compare(rows, orows) == []
    assert st.status & ~0x1 == 0
    assert st.committed_events == sum(r[0] for r in orows)


# ---------------------------------------------------------------------------------------------------
# models/traffic: INIT reads ./topology.txt (models/traffic/init.c:280-313)
def traffic_topology(name):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_traffic_topology
    return make_traffic_topology.topology(name)


def run_traffic(emu, name, T, wev, **cfg):
    text, n = traffic_topology(name)
    lib = rs_host.build_model("traffic", emu=emu)
    with rs_host.Engine(lib) as eng:
        eng.model_file("topology.txt", text.encode())
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=wev, arena_bytes=1 << 17,
                 max_payload=16, bucket_width=4.0, n_buckets=4096, **cfg)
        assert eng.run()
        st = eng.stats()
        rows = rs_trace.engine_rows(eng, 8)
        eng.model_file("topology.txt", None)
    _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text})
    return rows, orows, st


@pytest.mark.parametrize("name,T,wev,cfg", [
    ("traffic_grid_lp16", 20000, 4096, {}),
    ("traffic_grid_lp960", 3000, 1 << 16, {}),
    ("traffic_grid_lp960", 1500, 512, dict(ckpt_period=3)),
    ("traffic_grid_lp4096", 600, 1 << 18, {}),
])
def test_traffic_emu_vs_oracle(oracle_built, name, T, wev, cfg):
    """The reference's models/traffic, sources unchanged: INIT on the engine parses the topology image the host loaded
    (in-memory fopen/fgets/rewind/strtok_r/strtod), cars are malloc'ed/freed in events, RandomRange, integer abs() of
    a double, model-local Gaussian through the portable log, exp in the accident probability.  Bit-exact."""
    if not rs_host.model_sources("traffic") or not os.path.exists(os.path.join(oracle_built, "seqp_traffic")):
        pytest.skip("traffic sources live in the reference tree")
    rows, orows, st = run_traffic(True, name, T, wev, **cfg)
    assert rs_trace.compare(rows, orows) == []
    assert st.committed_events == sum(r[0] for r in orows)
    if name != "traffic_grid_lp16":
        assert st.rollbacks > 0


def test_traffic_emu_vs_reference_golden(oracle_built):
    """Against the UNMODIFIED reference's --sequential digest (libm): integers bit-exact, timestamps within 64 ULP."""
    if not rs_host.model_sources("traffic"):
        pytest.skip("traffic sources live in the reference tree")
    rows, _, _ = run_traffic(True, "traffic_grid_lp960", 3000, 1 << 16)
    _, grows = rs_trace.parse_digest(os.path.join(os.path.dirname(__file__), "golden", "traffic_grid_lp960_T3000.trace"))
    assert rs_trace.compare(rows, grows, state_mask=[], ts_ulp=64, check_hash=False) == []


def test_missing_input_file_fails_init(oracle_built, tmp_path, monkeypatch):
    """models/traffic/init.c:289-294: fopen fails -> perror + exit(EXIT_FAILURE).  Here the host notices before INIT."""
    if not rs_host.model_sources("traffic"):
        pytest.skip("traffic sources live in the reference tree")
    monkeypatch.chdir(tmp_path)
    lib = rs_host.build_model("traffic", emu=True)
    with rs_host.Engine(lib) as eng:
        with pytest.raises(rs_host.RsError) as ei:
            eng.init(n_lps=16, master_seed=1, simulation_time=10.0, arena_bytes=1 << 16, max_payload=16)
        assert ei.value.code == 5 and "topology.txt" in str(ei.value)
    # ... and found in the working directory when it is there (what the reference's fopen does)
    text, n = traffic_topology("traffic_grid_lp16")
    (tmp_path / "topology.txt").write_text(text)
    with rs_host.Engine(lib) as eng:
        eng.init(n_lps=n, master_seed=1, simulation_time=500.0, arena_bytes=1 << 16, max_payload=16, bucket_width=4.0)
        assert eng.run()
        assert eng.stats().committed_events > n


def test_traffic_makefile_flow_and_cli_emu(oracle_built, tmp_path):
    """The reference's own build recipe for the model (models/traffic/Makefile: one `rootsim-cc -c` per source, then
    a link of the objects) and its way of running it: the executable finds ./topology.txt like the reference's."""
    import subprocess
    import sys
    srcs = rs_host.model_sources("traffic")
    if not srcs:
        pytest.skip("traffic sources live in the reference tree")
    cc = [sys.executable, rs_host.ROOTSIM_CC, "--emu"]
    objs = []
    for s in srcs:
        o = str(tmp_path / (os.path.basename(s)[:-2] + ".o"))
        subprocess.run(cc + ["-Wall", "-Wextra", "-c", s, "-o", o], check=True, capture_output=True, cwd=tmp_path)
        objs.append(o)
    subprocess.run(cc + ["-Wall", "-Wextra", "-o", "traffic"] + objs, check=True, capture_output=True, cwd=tmp_path)
    text, n = traffic_topology("traffic_grid_lp16")
    (tmp_path / "topology.txt").write_text(text)
    tr = tmp_path / "trace.txt"
    r = subprocess.run([str(tmp_path / "traffic"), "--lp", str(n), "--simulation-time", "5000", "--seed", str(rs_trace.MASTER_SEED),
                        "--wt", "2", "--output-dir", str(tmp_path / "out"), "--b200-trace", str(tr)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "SIMULATION CORRECTLY COMPLETED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # no device sizing on the command line, as with the reference: the per-LP heap defaults to a share of 8 GiB, and the
    # run that found its 0-byte payload slots too small was started again with room for the model's 16-byte payloads
    assert "restarting with --b200-max-payload 16" in r.stderr
    assert "LP 0 (j0.0) is an intersection with 2 neighbours" in r.stdout      # models/traffic/init.c:315
    _, rows = rs_trace.parse_digest(str(tr))
    _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=5000, horizon=5000, files={"topology.txt": text})
    assert rs_trace.compare(rows, orows) == []
    os.remove(tmp_path / "topology.txt")
    r = subprocess.run([str(tmp_path / "traffic"), "--lp", str(n), "--simulation-time", "50"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "topology.txt" in r.stdout + r.stderr


@pytest.mark.parametrize("model,n,T,wev", [("packet", 256, 20000, 4096), ("packet", 5000, 2000, 1 << 16), ("collector", 256, 400, 4096),
                                           ("collector", 4000, 25, 1 << 16)])
def test_small_reference_models_emu_vs_oracle(oracle_built, model, n, T, wev):
    """models/packet (TOPOLOGY_GRAPH, 24-byte payload with a pointer in it, two mallocs in INIT) and models/collector
    (TOPOLOGY_STAR: every LP sends to LP 0, which becomes one long event chain) -- reference sources unchanged,
    horizons below the models' own OnGVT termination."""
    if not rs_host.model_sources(model) or not os.path.exists(os.path.join(oracle_built, "seqp_" + model)):
        pytest.skip("%s sources live in the reference tree" % model)
    rows, st = run_emu(model, n, T, wev, state_bytes=4, arena_bytes=256, max_payload=32, bucket_width=8.0, n_buckets=4096)
    _, orows, _ = rs_trace.run_oracle(model, n, sim_time=T, horizon=T, state_bytes=4)
    assert rs_trace.compare(rows, orows) == []
    assert st.committed_events == sum(r[0] for r in orows)


def test_traffic_shipped_topology_emu_vs_oracle_and_reference(oracle_built):
    """The topology the reference ships (models/traffic/topology.txt: the Italian motorway graph, 62 junctions + 962
    one-km segments = 1024 LPs), where the reference tree is present: engine = oracle bit for bit, oracle (libm) = the
    unmodified reference's --sequential run bit for bit."""
    from conftest import have_reference_binary
    path = os.path.join(os.environ.get("RS_REFERENCE", "/root/reference"), "models", "traffic", "topology.txt")
    if not os.path.exists(path) or not rs_host.model_sources("traffic"):
        pytest.skip("the reference tree is not here")
    text = open(path).read()
    n, T = 1024, 3600
    lib = rs_host.build_model("traffic", emu=True)
    with rs_host.Engine(lib) as eng:
        eng.model_file("topology.txt", text.encode())
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=1 << 16, arena_bytes=1 << 17,
                 max_payload=16, bucket_width=4.0, n_buckets=4096)
        assert eng.run()
        rows = rs_trace.engine_rows(eng, 8)
        eng.model_file("topology.txt", None)
    _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text})
    assert rs_trace.compare(rows, orows) == []
    assert sum(r[0] for r in orows) > 300000
    if have_reference_binary("traffic"):
        _, lrows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text}, portable=False)
        _, rrows, _ = rs_trace.run_reference("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text})
        assert rs_trace.compare(lrows, rrows) == []
        assert rs_trace.compare(rows, rrows, state_mask=[], ts_ulp=64, check_hash=False) == []


@pytest.mark.parametrize("n,T,wev,cfg", [(200, 300, 4096, {}), (1000, 100, 100000, {}), (50, 3000, 512, dict(ckpt_period=1)), (64, 400, 64, dict(ckpt_period=1000))])
def test_heap_stress_emu_vs_oracle(oracle_built, n, T, wev, cfg):
    """After the reference's allocator unit test (tests/dymelor.c): random malloc / calloc / realloc / free over a set of
    bins in the LP's rollbackable heap, every block pattern-checked before it is touched again, under thousands of
    rollbacks.  First 36 state bytes: ops, failures (must be 0), hash of everything read back, live bytes, call counts."""
    import struct
    rows, st = run_emu("heapstress", n, T, wev, state_bytes=36, arena_bytes=1 << 15, bucket_width=0.5, n_buckets=1024, **cfg)
    _, orows, _ = rs_trace.run_oracle("heapstress", n, sim_time=T, horizon=T, state_bytes=36)
    assert rs_trace.compare(rows, orows) == []
    assert st.rollbacks > 0
    tot = [sum(struct.unpack("<IIQIIIII", r[4])[i] for r in rows) for i in range(8)]
    assert tot[1] == 0 and min(tot[4:8]) > 1000         # no failed check; mallocs, callocs, reallocs and frees all happened


def test_traffic_randomized_configurations_emu(oracle_built):
    """Eight seeded random draws of (grid, horizon, master seed, window size, bucket width, checkpoint period): the
    committed trace is the oracle's whatever the window partition and the checkpoint spacing (zero-delay hand-offs
    between neighbours, 16-byte payloads, up to 100,000 rollbacks in the larger draws of the same generator)."""
    import random
    if not rs_host.model_sources("traffic") or not os.path.exists(os.path.join(oracle_built, "seqp_traffic")):
        pytest.skip("traffic sources live in the reference tree")
    rng = random.Random(7)
    lib = rs_host.build_model("traffic", emu=True)
    for _ in range(8):
        name = rng.choice(["traffic_grid_lp16", "traffic_grid_lp960"])
        T = rng.choice([1000, 2000]) if name != "traffic_grid_lp16" else rng.choice([20000, 40000])
        seed = rng.getrandbits(62) | 1
        wev, bw, p = rng.choice([64, 512, 4096, 1 << 16, 1 << 20]), rng.choice([0.25, 1.0, 4.0, 16.0]), rng.choice([1, 3, 10, 1000])
        text, n = traffic_topology(name)
        _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, seed=seed, files={"topology.txt": text})
        with rs_host.Engine(lib) as eng:
            eng.model_file("topology.txt", text.encode())
            eng.init(n_lps=n, master_seed=seed, simulation_time=float(T), window_events=wev, arena_bytes=1 << 17, max_payload=16,
                     bucket_width=bw, n_buckets=4096, ckpt_period=p)
            assert eng.run()
            rows = rs_trace.engine_rows(eng, 8)
            eng.model_file("topology.txt", None)
        assert rs_trace.compare(rows, orows) == [], (name, T, seed, wev, bw, p)


def test_traffic_ragged_topology_file(oracle_built):
    """A topology file with blank lines, comments between the data lines and NO trailing newline: the model strips the
    last character of every line it reads (models/traffic/init.c:44,302), so the last segment loses its length, gets
    zero queue slots and answers every arrival with "Object queue full" -- on the device exactly as in the reference."""
    from conftest import have_reference_binary
    if not rs_host.model_sources("traffic") or not os.path.exists(os.path.join(oracle_built, "seqp_traffic")):
        pytest.skip("traffic sources live in the reference tree")
    text, n = traffic_topology("traffic_grid_lp16")
    out = []
    for i, line in enumerate(text.rstrip("\n").split("\n")):
        out.append(line)
        if i % 5 == 2:
            out += ["", "# a comment, with, commas"]
    ragged = "\n".join(out)
    T = 20000
    lib = rs_host.build_model("traffic", emu=True)
    with rs_host.Engine(lib) as eng:
        eng.model_file("topology.txt", ragged.encode())
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=float(T), window_events=4096, arena_bytes=1 << 17,
                 max_payload=16, bucket_width=4.0, n_buckets=4096)
        assert eng.run()
        rows = rs_trace.engine_rows(eng, 8)
        eng.model_file("topology.txt", None)
    _, orows, so = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": ragged})
    assert rs_trace.compare(rows, orows) == []
    assert so.count("Object queue full") > 100
    if have_reference_binary("traffic"):
        _, lrows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": ragged}, portable=False)
        _, rrows, _ = rs_trace.run_reference("traffic", n, sim_time=T, horizon=T, files={"topology.txt": ragged})
        assert rs_trace.compare(lrows, rrows) == []
