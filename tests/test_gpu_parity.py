"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every run goes through the C-ABI of the
product library (lib<model>.so, nvcc sm_100a) and is compared with the CPU oracle built with the same
portable log => BIT-EXACT on every field, timestamps and the per-LP event hash included; against the
real reference's golden digests (libm log) integers are bit-exact and timestamps within 64 ULP."""
import os

import pytest

import rs_host
import rs_trace

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
PHOLD_MASK = [0, 4, 5, 6, 7]


def run_gpu(model, n, T, margs=(), state_bytes=8, seed=rs_trace.MASTER_SEED, **cfg):
    lib = rs_host.build_model(model)
    # model options live in file-scope variables of the library: only a run that sets them needs a private
    # copy; everything else loads the in-tree root-sim_b200/_built/lib<model>.so itself
    with rs_host.Engine(lib, private=bool(margs)) as eng:
        assert eng.info().engine == b"cuda-sm100a"
        if margs:
            eng.model_args(list(margs))
        eng.init(n_lps=n, master_seed=seed, simulation_time=T, **cfg)
        assert eng.run()
        st = eng.stats()
        rows = rs_trace.engine_rows(eng, state_bytes)
    return rows, st


def need(model, oracle_built):
    try:
        rs_host.build_model(model)
    except FileNotFoundError:
        pytest.skip("no library for model %s" % model)
    if not os.path.exists(os.path.join(oracle_built, "seqp_" + model)):
        pytest.skip("no oracle for model %s" % model)


@pytest.mark.parametrize("n,T,wev", [(64, 100, 100000), (1024, 60, 4096), (1024, 100, 1 << 20), (257, 80, 50),
                                     (4096, 40, 100000), (1, 50, 100), (2, 50, 100), (65536, 25, 1 << 20),
                                     (262144, 40, 1 << 20)])
def test_phold_vs_oracle(oracle_built, n, T, wev):
    need("phold", oracle_built)
    rows, st = run_gpu("phold", n, T, window_events=wev, arena_bytes=128)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.committed_events == sum(r[0] for r in orows)
    assert st.kernel_launches > 0


def test_phold_vs_reference_golden(oracle_built):
    """Against the UNMODIFIED reference's --sequential digest (tests/golden, libm log)."""
    need("phold", oracle_built)
    for name, n, T in (("phold_lp1024_T50", 1024, 50), ("phold_lp100_T120", 100, 120)):
        _, grows = rs_trace.parse_digest(os.path.join(GOLD, name + ".trace"))
        rows, _ = run_gpu("phold", n, T, window_events=8192, arena_bytes=128)
        assert rs_trace.compare(rows, grows, state_mask=PHOLD_MASK, ts_ulp=64, check_hash=False) == []


def test_phold_1m_lps_vs_oracle(oracle_built):
    """BASELINE config 2 size (1,048,576 LPs) on a short horizon the oracle finishes in seconds."""
    need("phold", oracle_built)
    n, T = 1 << 20, 12.0
    rows, st = run_gpu("phold", n, T, window_events=1 << 20, arena_bytes=128, max_pending=16 << 20)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.committed_events == sum(r[0] for r in orows)


@pytest.mark.parametrize("n,T,wev,margs", [
    (48, 200, 100000, ["--order-sensitive"]),
    (63, 300, 50, ["--order-sensitive", "--mean", "1.5"]),
    (1000, 100, 100000, []),
    (300, 150, 2000, ["--token-prob", "0.9"]),
    (20000, 60, 1 << 18, []),
])
def test_mixnet_vs_oracle(oracle_built, n, T, wev, margs):
    need("mixnet", oracle_built)
    rows, st = run_gpu("mixnet", n, T, margs, state_bytes=32, seed=42, window_events=wev, arena_bytes=1024, max_payload=16, bucket_width=0.25)
    _, orows, _ = rs_trace.run_oracle("mixnet", n, sim_time=T, horizon=T, state_bytes=32, seed=42, extra=margs)
    assert rs_trace.compare(rows, orows) == []
    assert st.rollbacks > 0


@pytest.mark.parametrize("n,T", [(16, 600), (1024, 150), (65536, 40)])
def test_pcs_vs_oracle(oracle_built, n, T):
    """models/pcs (BASELINE config 3 is 65536 cells = 256 x 256 hexagons): unchanged sources, 40-byte payloads,
    malloc/free in events, zero-delay hand-offs.  Compared: counts, event hash, timestamps, RNG seeds and the
    first 56 bytes of the state (all counters, lvt, ta); the per-call power/fading doubles further down in
    the heap go through pow(), for which the device uses CUDA's libm (not part of the trace)."""
    need("pcs", oracle_built)
    rows, st = run_gpu("pcs", n, T, state_bytes=56, window_events=1 << 18, arena_bytes=32768, max_payload=48,
                       bucket_width=0.25, n_buckets=4096, max_pending=8 * n + (1 << 20))
    _, orows, _ = rs_trace.run_oracle("pcs", n, sim_time=T, horizon=T, state_bytes=56)
    assert rs_trace.compare(rows, orows) == []
    assert st.committed_events == sum(r[0] for r in orows)


def test_size_independent_properties_at_full_size(oracle_built):
    """1M LPs, longer horizon than the oracle can check quickly: the committed trace must not depend on
    how the run is cut into windows (a checksum of per-LP digests), committed == sum of per-LP counts,
    every rolled back execution is accounted for."""
    need("phold", oracle_built)
    n, T = 1 << 20, 30.0
    a, sa = run_gpu("phold", n, T, window_events=1 << 20, arena_bytes=128, max_pending=24 << 20)
    b, sb = run_gpu("phold", n, T, window_events=1 << 18, bucket_width=0.125, arena_bytes=128, max_pending=24 << 20)
    assert sa.windows != sb.windows
    assert rs_trace.compare(a, b, state_mask=PHOLD_MASK) == []
    for st, rows in ((sa, a), (sb, b)):
        assert st.committed_events == sum(r[0] for r in rows)
        assert st.executed_events + st.silent_events >= st.committed_events
        assert st.status & ~0x1 == 0


def test_model_termination_by_ongvt(oracle_built):
    need("mixnet", oracle_built)
    lib = rs_host.build_model("mixnet")
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--tick-limit", "30"])
        eng.init(n_lps=500, master_seed=5, simulation_time=0.0, window_events=5000, arena_bytes=1024, max_payload=16)
        assert eng.run()
        st = eng.stats()
        states = eng.lp_state(4)
    ticks = [int.from_bytes(states[4 * i:4 * i + 4], "little") for i in range(500)]
    assert st.all_lps_agree == 1 and min(ticks) >= 30


def test_fatal_model_error_is_reported(oracle_built):
    """An event scheduled in the past aborts the run (src/communication/communication.c:285-289)."""
    need("mixnet", oracle_built)
    lib = rs_host.build_model("mixnet")
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--mean", "-1"])        # Expent() with a negative mean: the reference aborts too
        with pytest.raises(rs_host.RsError):
            eng.init(n_lps=64, master_seed=5, simulation_time=50.0, arena_bytes=1024, max_payload=16)
            eng.run()


EDGE = [
    (0, None, None), (1, 5, "PAST_EVENT"), (2, 5, "RESERVED_TYPE"), (3, None, "BAD_RECEIVER"), (4, 5, "MODEL_EXIT"),
    (5, 6, "ARENA_FULL"), (6, None, "NO_EVENTS"), (7, None, None),
]


@pytest.mark.parametrize("fault,code,flag", EDGE, ids=["ok", "past", "reserved-type", "bad-receiver", "exit", "arena", "no-events", "one-active"])
def test_edge_cases(oracle_built, fault, code, flag):
    """Error behaviour of the event API (src/communication/communication.c:280-294), model exit, heap
    exhaustion, a model that schedules nothing, and fifteen LPs that only ever see INIT."""
    need("edgecases", oracle_built)
    lib = rs_host.build_model("edgecases")
    with rs_host.Engine(lib, private=True) as eng:
        eng.model_args(["--fault", str(fault)])
        eng.init(n_lps=16, master_seed=9, simulation_time=30.0, window_events=256, arena_bytes=512)
        err = None
        try:
            fin = eng.run()
        except rs_host.RsError as ex:
            fin, err = None, ex
        st = eng.stats().as_dict()
        rows = rs_trace.engine_rows(eng, 8)
    if code is None:
        assert err is None and fin
    else:
        assert err is not None and err.code == code
    if flag:
        assert flag in st["status_names"]
    if fault in (0, 3, 7):
        _, orows, _ = rs_trace.run_oracle("edgecases", 16, sim_time=30.0, horizon=30.0, seed=9, extra=["--fault", str(fault)])
        assert rs_trace.compare(rows, orows) == []
    if fault == 6:
        assert [r[0] for r in rows] == [1] * 16


def test_cli_executable(oracle_built, tmp_path):
    """The ROOT-Sim style executable (host/rs_main.c + lib<model>.so, built by __graft_entry__.build())."""
    import subprocess
    need("mixnet", oracle_built)
    exe = os.path.join(rs_host.BUILT, "mixnet")
    if not os.path.exists(exe):
        pytest.skip("executable not built")
    out, tr = tmp_path / "out", tmp_path / "trace.txt"
    r = subprocess.run([exe, "--lp", "2000", "--simulation-time", "60", "--seed", "42", "--wt", "4", "--b200-arena-bytes", "1024",
                        "--b200-max-payload", "16", "--b200-bucket-width", "0.25", "--output-dir", str(out), "--b200-trace", str(tr),
                        "--mean", "3.0"], capture_output=True, text=True)
    assert r.returncode == 0 and "SIMULATION CORRECTLY COMPLETED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    stats = open(out / "execution_stats").read()
    assert "TOTAL COMMITTED EVENTS" in stats and "EFFICIENCY" in stats
    _, rows = rs_trace.parse_digest(str(tr))
    _, orows, _ = rs_trace.run_oracle("mixnet", 2000, sim_time=60, horizon=60, seed=42, extra=["--mean", "3.0"])
    assert rs_trace.compare(rows, orows) == []


def test_time_stepped_big_lps_gpu(oracle_built):
    """substeps (pause/resume of big LPs at checkpoint records, ordered arrival arrays kept across passes)"""
    need("phold", oracle_built)
    n, T = 262144, 30
    rows, st = run_gpu("phold", n, T, window_events=1 << 20, arena_bytes=128, substeps=4, bucket_width=0.125)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []


def test_event_latency_probe(oracle_built):
    """rs_dev_probe_event: the handler's serial latency on one device thread"""
    need("phold", oracle_built)
    lib = rs_host.build_model("phold")
    with rs_host.Engine(lib) as eng:
        eng.init(n_lps=1024, master_seed=1, simulation_time=10.0, max_out=1 << 18)
        ns = eng.probe_event(5, 3, 2000, True)
        assert 50.0 < ns < 20000.0


def test_two_live_engines_of_one_library(oracle_built):
    """Two device engines from ONE loaded library, run in alternation a few windows at a time: every rs_dev_run claims
    the library's __constant__ engine descriptor for its own engine before it launches model code."""
    need("phold", oracle_built)
    lib = rs_host.build_model("phold")
    a, b = rs_host.Engine(lib), rs_host.Engine(lib)
    a.init(n_lps=4096, master_seed=rs_trace.MASTER_SEED, simulation_time=40.0, window_events=8192, arena_bytes=128)
    b.init(n_lps=1000, master_seed=rs_trace.MASTER_SEED, simulation_time=60.0, window_events=4096, arena_bytes=128)
    fa = fb = False
    for _ in range(10000):
        if not fa:
            fa = a.run(max_windows=1)
        if not fb:
            fb = b.run(max_windows=2)
        if fa and fb:
            break
    assert fa and fb
    for eng, n, T in ((a, 4096, 40), (b, 1000, 60)):
        rows = rs_trace.engine_rows(eng, 8)
        _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
        assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    a.fini()
    b.fini()


@pytest.mark.parametrize("n,T,nb,bw", [(4096, 40, 64, 0.0625), (65536, 20, 128, 0.03125)])
def test_far_future_events_wait_on_the_far_list(oracle_built, n, T, nb, bw):
    """calendar ring of 4 time units: most PHOLD events (mean increment 5) are committed beyond the ring's horizon,
    wait on the far list and are re-filed when the ring has advanced (src/datatypes/list.h:163-208 takes any timestamp)"""
    need("phold", oracle_built)
    rows, st = run_gpu("phold", n, T, window_events=1 << 20, arena_bytes=128, n_buckets=nb, bucket_width=bw)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.status & ~0x1 == 0


@pytest.mark.parametrize("cfg", [dict(flags=rs_host.RS_F_TRACE | rs_host.RS_F_SERIAL_RUN), dict(helper_lanes=2), dict(flags=rs_host.RS_F_TRACE | rs_host.RS_F_KTIME, helper_lanes=2),
                                 dict(hot_threshold=8), dict(chunk_events=64)])
def test_execution_modes_commit_the_same_trace(oracle_built, cfg):
    """The scheduling choices of the engine -- ordinary LPs beside or after the chain LPs (two streams / one), arrival
    merge by the helper lanes or by lane 0 alone, which LPs get a warp of their own, the visit budget -- change how the
    work is laid out on the GPU, never the committed trace."""
    need("phold", oracle_built)
    n, T = 65536, 25
    rows, st = run_gpu("phold", n, T, window_events=1 << 20, arena_bytes=128, **cfg)
    _, orows, _ = rs_trace.run_oracle("phold", n, sim_time=T, horizon=T)
    assert rs_trace.compare(rows, orows, state_mask=PHOLD_MASK) == []
    assert st.status & ~0x1 == 0


# ---------------------------------------------------------------------------------------------------
# models/traffic (BASELINE config 5): INIT reads ./topology.txt
def run_traffic_gpu(name, T, wev, **cfg):
    import sys
    sys.path.insert(0, GOLD)
    import make_traffic_topology
    text, n = make_traffic_topology.topology(name)
    lib = rs_host.build_model("traffic")
    with rs_host.Engine(lib) as eng:
        assert eng.info().engine == b"cuda-sm100a"
        eng.model_file("topology.txt", text.encode())
        eng.init(n_lps=n, master_seed=rs_trace.MASTER_SEED, simulation_time=T, window_events=wev, arena_bytes=1 << 17,
                 max_payload=16, bucket_width=4.0, n_buckets=4096, **cfg)
        assert eng.run()
        st = eng.stats()
        rows = rs_trace.engine_rows(eng, 8)
        eng.model_file("topology.txt", None)
    return rows, st, text, n


@pytest.mark.parametrize("name,T,wev,cfg", [
    ("traffic_grid_lp16", 20000, 4096, {}),
    ("traffic_grid_lp960", 3000, 1 << 16, {}),
    ("traffic_grid_lp960", 1500, 512, dict(ckpt_period=3)),
    ("traffic_grid_lp4096", 2000, 1 << 18, {}),
])
def test_traffic_vs_oracle(oracle_built, name, T, wev, cfg):
    """The reference's models/traffic, sources unchanged.  INIT runs on the device and parses the topology image the
    host loaded (in-memory fopen/fgets/rewind/strtok_r/strtod of csrc/rs_prelude.cuh); cars are malloc'ed and freed
    inside events, the queue is a linked list in the LP's heap (up to 1000 cars of 64 bytes at a junction).
    Compared bit-exactly: counts, event hash, timestamps, RNG seeds, lvt in the state."""
    need("traffic", oracle_built)
    rows, st, text, n = run_traffic_gpu(name, T, wev, **cfg)
    _, orows, _ = rs_trace.run_oracle("traffic", n, sim_time=T, horizon=T, files={"topology.txt": text})
    assert rs_trace.compare(rows, orows) == []
    assert st.committed_events == sum(r[0] for r in orows)
    assert st.kernel_launches > 0


def test_traffic_vs_reference_golden(oracle_built):
    """Against the UNMODIFIED reference's --sequential digest (tests/golden, libm log/exp): integers bit-exact,
    timestamps within 64 ULP."""
    need("traffic", oracle_built)
    for name, gold, T in (("traffic_grid_lp16", "traffic_grid_lp16_T20000", 20000), ("traffic_grid_lp960", "traffic_grid_lp960_T3000", 3000)):
        rows, _, _, _ = run_traffic_gpu(name, T, 1 << 16)
        _, grows = rs_trace.parse_digest(os.path.join(GOLD, gold + ".trace"))
        assert rs_trace.compare(rows, grows, state_mask=[], ts_ulp=64, check_hash=False) == []


@pytest.mark.parametrize("model,n,T,wev", [("packet", 256, 20000, 4096), ("packet", 65536, 1500, 1 << 18), ("collector", 256, 400, 4096),
                                           ("collector", 4000, 25, 1 << 16)])
def test_small_reference_models_vs_oracle(oracle_built, model, n, T, wev):
    """models/packet (TOPOLOGY_GRAPH, 24-byte payloads) and models/collector (TOPOLOGY_STAR: LP 0 receives everything),
    reference sources unchanged, horizons below the models' own OnGVT termination.  Bit-exact."""
    need(model, oracle_built)
    rows, st = run_gpu(model, n, T, state_bytes=4, window_events=wev, arena_bytes=256, max_payload=32, bucket_width=8.0, n_buckets=4096)
    _, orows, _ = rs_trace.run_oracle(model, n, sim_time=T, horizon=T, state_bytes=4)
    assert rs_trace.compare(rows, orows) == []
    assert st.committed_events == sum(r[0] for r in orows)


@pytest.mark.parametrize("n,T,wev,cfg", [(200, 300, 4096, {}), (20000, 60, 1 << 18, {}), (50, 3000, 512, dict(ckpt_period=1))])
def test_heap_stress_vs_oracle(oracle_built, n, T, wev, cfg):
    """After the reference's allocator unit test (tests/dymelor.c): random malloc / calloc / realloc / free in the LP's
    rollbackable heap with pattern checks, under rollbacks (tests/models/heapstress).  Bit-exact, no failed check."""
    import struct
    need("heapstress", oracle_built)
    rows, st = run_gpu("heapstress", n, T, state_bytes=36, window_events=wev, arena_bytes=1 << 15, bucket_width=0.5, n_buckets=1024, **cfg)
    _, orows, _ = rs_trace.run_oracle("heapstress", n, sim_time=T, horizon=T, state_bytes=36)
    assert rs_trace.compare(rows, orows) == []
    assert sum(struct.unpack("<IIQIIIII", r[4])[1] for r in rows) == 0
