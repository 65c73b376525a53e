"""Helpers shared by the tests: run the oracle / reference binaries, parse and compare trace digests."""
import os
import struct
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))

ORACLE_BUILD = os.path.join(REPO, "oracle", "_build")
ORACLE_REF = os.path.join(REPO, "oracle", "_ref")
MASTER_SEED = 12345678901234567


def parse_digest(path):
    """-> (header dict, list of (count, hash, last_ts_bits, seed, state_bytes))"""
    rows = []
    hdr = {}
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                for tok in line.split():
                    if "=" in tok:
                        k, v = tok.split("=", 1)
                        hdr[k] = v
                continue
            p = line.split()
            rows.append((int(p[1]), int(p[2], 16), int(p[3], 16), int(p[4], 16), bytes.fromhex(p[5]) if len(p) > 5 else b""))
    return hdr, rows


def run_oracle(model, n_lps, sim_time=0.0, horizon=None, portable=True, state_bytes=8, seed=MASTER_SEED, extra=(), files=None):
    """Run oracle/_build/seq[p]_<model>; returns (header, rows, stdout).  files: {name: text} of the input files a
    model reads inside INIT (written to the working directory of the run)."""
    exe = os.path.join(ORACLE_BUILD, ("seqp_" if portable else "seq_") + model)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "trace.txt")
        env = dict(os.environ, RS_TRACE_OUT=out, RS_TRACE_STATE_BYTES=str(state_bytes))
        if horizon is not None:
            env["RS_TRACE_HORIZON"] = repr(float(horizon))
        cmd = [exe, "--lp", str(n_lps), "--seed", str(seed)] + list(extra)
        if sim_time:
            cmd += ["--simulation-time", repr(float(sim_time))]
        for fn, text in (files or {}).items():
            with open(os.path.join(td, fn), "w") as f:
                f.write(text)
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True, cwd=td)
        hdr, rows = parse_digest(out)
    return hdr, rows, r.stdout


def run_reference(model, n_lps, sim_time=0.0, horizon=None, state_bytes=8, seed=MASTER_SEED, extra=(), mode=("--sequential",), files=None):
    """Run the REAL reference built by oracle/build_ref.sh with the tracing shim."""
    exe = os.path.join(ORACLE_REF, model + "_trace")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, ".rootsim"))
        with open(os.path.join(td, ".rootsim", "numerical.conf"), "w") as f:
            f.write("%d\n" % seed)
        out = os.path.join(td, "trace.txt")
        env = dict(os.environ, HOME=td, RS_TRACE_OUT=out, RS_TRACE_STATE_BYTES=str(state_bytes))
        if horizon is not None:
            env["RS_TRACE_HORIZON"] = repr(float(horizon))
        cmd = [exe] + list(mode) + ["--lp", str(n_lps), "--deterministic-seed", "--output-dir", os.path.join(td, "out")] + list(extra)
        if sim_time:
            cmd += ["--simulation-time", str(int(sim_time))]
        for fn, text in (files or {}).items():
            with open(os.path.join(td, fn), "w") as f:
                f.write(text)
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=td)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        hdr, rows = parse_digest(out)
    return hdr, rows, r.stdout


def engine_rows(eng, state_bytes=8):
    """Digest rows of a device engine in the same shape as parse_digest()."""
    tr = eng.trace()
    n = len(tr)
    st = eng.lp_state(state_bytes, n) if state_bytes else b""
    rows = []
    for i in range(n):
        bits = struct.unpack("<Q", struct.pack("<d", tr[i].last_ts))[0]
        rows.append((tr[i].count, tr[i].hash, bits, tr[i].seed, st[i * state_bytes:(i + 1) * state_bytes]))
    return rows


def ulp_diff(a_bits, b_bits):
    return abs(a_bits - b_bits)


def compare(rows_a, rows_b, state_mask=None, ts_ulp=0, check_hash=True):
    """Return a list of human-readable mismatches (empty = equal).  state_mask: iterable of byte
    offsets to compare (None = all).  ts_ulp: tolerance on last_ts in ULPs (0 = bit-exact)."""
    bad = []
    if len(rows_a) != len(rows_b):
        return ["row count %d vs %d" % (len(rows_a), len(rows_b))]
    for i, (a, b) in enumerate(zip(rows_a, rows_b)):
        if a[0] != b[0]:
            bad.append("lp %d count %d vs %d" % (i, a[0], b[0]))
        if check_hash and a[1] != b[1]:
            bad.append("lp %d hash %016x vs %016x" % (i, a[1], b[1]))
        if ulp_diff(a[2], b[2]) > ts_ulp:
            bad.append("lp %d last_ts bits %016x vs %016x" % (i, a[2], b[2]))
        if a[3] != b[3]:
            bad.append("lp %d seed %016x vs %016x" % (i, a[3], b[3]))
        sa, sb = a[4], b[4]
        idx = range(min(len(sa), len(sb))) if state_mask is None else state_mask
        for k in idx:
            if k < len(sa) and k < len(sb) and sa[k] != sb[k]:
                bad.append("lp %d state byte %d: %02x vs %02x" % (i, k, sa[k], sb[k]))
                break
        if len(bad) > 20:
            bad.append("...")
            break
    return bad


# ---------------------------------------------------------------------------------------------------
# Order-independent 64-bit checksum of a committed trace: sum over LPs (mod 2^64) of a splitmix64 chain over
# (global LP id, event count, event hash, last timestamp bits, RNG seed).  Partial sums over disjoint LP
# blocks add up, so every rank of a multi-GPU run checksums its own block and the sums are all-reduced.
# bench.py compares it with tests/golden/phold_checksums.json (made by tests/golden/make_checksums.py from
# the CPU oracle) for the exact configuration it times.
def _splitmix64(x):
    import numpy as np
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def checksum_arrays(counts, hashes, ts_bits, seeds, first_gid=0):
    """-> (checksum, total count); all inputs are array-likes of unsigned 64-bit integers of one LP block."""
    import numpy as np
    with np.errstate(over="ignore"):
        c = np.asarray(counts, dtype=np.uint64)
        g = np.arange(first_gid, first_gid + len(c), dtype=np.uint64)
        x = _splitmix64(np.asarray(seeds, dtype=np.uint64))
        x = _splitmix64(x ^ np.asarray(ts_bits, dtype=np.uint64))
        x = _splitmix64(x ^ np.asarray(hashes, dtype=np.uint64))
        x = _splitmix64(x ^ c)
        x = _splitmix64(x ^ g)
        return int(x.sum(dtype=np.uint64)), int(c.sum(dtype=np.uint64))


def checksum_rows(rows, first_gid=0):
    """checksum of digest rows as returned by parse_digest() / engine_rows()"""
    return checksum_arrays([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows], first_gid)


def checksum_digests(tr, first_gid=0):
    """checksum of the per-LP digest array rs_dev_trace returned (32 bytes per LP)"""
    import numpy as np
    a = np.frombuffer(tr, dtype=np.dtype([("count", "<u8"), ("hash", "<u8"), ("ts", "<u8"), ("seed", "<u8")]))
    return checksum_arrays(a["count"], a["hash"], a["ts"], a["seed"], first_gid)


def checksum_engine(eng, first_gid=0):
    """checksum of a device engine's per-LP digests (one D2H copy of 32 bytes per LP through rs_dev_trace)"""
    return checksum_digests(eng.trace(), first_gid)


def checksum_digest_file(path):
    """checksum of a digest file written by the oracle, without building Python tuples (16M-LP files)"""
    import numpy as np
    cols = [[], [], [], []]
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            p = line.split()
            cols[0].append(int(p[1]))
            cols[1].append(int(p[2], 16))
            cols[2].append(int(p[3], 16))
            cols[3].append(int(p[4], 16))
    return checksum_arrays(*[np.array(c, dtype=np.uint64) for c in cols])
