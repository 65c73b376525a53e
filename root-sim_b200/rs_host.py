"""ctypes binding of the C-ABI in include/rootsim_b200.h (host-side harness for tests and bench.py).

The reference is compiled C and so is this project's host runtime (host/rs_main.c); this module is
only the thin foreign-function stub a Python tool needs: it mirrors the structs of the header
field by field and adds no logic of its own.  Every call goes to lib<model>.so produced by
bin/rootsim-cc (CUDA, sm_100a).  There is no CPU path: on a machine without a GPU
`Engine.init` raises RsError(RS_ERR_NO_DEVICE).
"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
BUILT = os.path.join(HERE, "_built")
ROOTSIM_CC = os.path.join(HERE, "bin", "rootsim-cc")

RS_ABI_VERSION = 3
RS_F_TRACE = 0x1
RS_F_KTIME = 0x2
RS_F_SERIAL_RUN = 0x4

STATUS_BITS = {
    0x0001: "BAD_RECEIVER", 0x0002: "PAST_EVENT", 0x0004: "RESERVED_TYPE", 0x0008: "MODEL_EXIT",
    0x0010: "ARENA_FULL", 0x0020: "POOL_FULL", 0x0040: "WINDOW_FULL", 0x0080: "OUT_FULL",
    0x0100: "HORIZON", 0x0200: "LATE_FULL", 0x0400: "PAYLOAD", 0x0800: "NO_EVENTS", 0x1000: "XCHG_FULL", 0x2000: "PEER_FATAL",
}
ERRORS = {0: "RS_OK", 1: "RS_ERR_NO_DEVICE", 2: "RS_ERR_BAD_CONFIG", 3: "RS_ERR_OOM", 4: "RS_ERR_CUDA",
          5: "RS_ERR_MODEL", 6: "RS_ERR_CAPACITY"}


class RsError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("n_lps", C.c_uint32), ("lp_first", C.c_uint32), ("lp_count", C.c_uint32),
        ("master_seed", C.c_uint64), ("simulation_time", C.c_double),
        ("ckpt_period", C.c_uint32), ("gvt_snapshot_cycles", C.c_uint32),
        ("bucket_width", C.c_double), ("n_buckets", C.c_uint32), ("window_events", C.c_uint32),
        ("arena_bytes", C.c_uint32), ("max_payload", C.c_uint32), ("flags", C.c_uint32),
        ("max_pending", C.c_uint64), ("max_window", C.c_uint32), ("max_out", C.c_uint32),
        ("device", C.c_int32), ("topology_geometry", C.c_uint32), ("topology_out_of", C.c_uint32),
        ("hot_threshold", C.c_uint32), ("run_ahead", C.c_uint32), ("rank", C.c_uint32), ("nranks", C.c_uint32),
        ("xchg_cap", C.c_uint32), ("chunk_events", C.c_uint32), ("substeps", C.c_uint32), ("pace_trail", C.c_uint32),
        ("window_span", C.c_double), ("burst_events", C.c_uint32), ("helper_lanes", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("committed_events", C.c_uint64), ("executed_events", C.c_uint64), ("silent_events", C.c_uint64),
        ("rollbacks", C.c_uint64), ("antimessages", C.c_uint64), ("checkpoints", C.c_uint64),
        ("checkpoint_bytes", C.c_uint64), ("windows", C.c_uint64), ("relax_rounds", C.c_uint64),
        ("late_events", C.c_uint64), ("pending_events", C.c_uint64), ("peak_window_events", C.c_uint64),
        ("gvt", C.c_double), ("loop_seconds", C.c_double), ("device_seconds", C.c_double),
        ("status", C.c_uint32), ("all_lps_agree", C.c_uint32), ("kernel_launches", C.c_uint64),
        ("chain_events", C.c_uint64), ("sent_records", C.c_uint64), ("reserved", C.c_uint64 * 4),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["big_groups"] = self.reserved[0]
        d["spill_rounds"] = self.reserved[1]
        d["run_ahead_lps"] = self.reserved[2]
        d["window_cuts"] = self.reserved[3]
        d["status_names"] = [n for b, n in STATUS_BITS.items() if self.status & b]
        return d


class LpTrace(C.Structure):
    _fields_ = [("count", C.c_uint64), ("hash", C.c_uint64), ("last_ts", C.c_double), ("seed", C.c_uint64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p)
ALLTOALL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class CommOps(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("allreduce_i64", ALLREDUCE_FN), ("alltoall", ALLTOALL_FN)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("total_ms", C.c_double), ("launches", C.c_uint64)]


class ModelInfo(C.Structure):
    _fields_ = [("model_argp", C.c_void_p), ("has_topology", C.c_int32), ("topology_type", C.c_uint32),
                ("topology_geometry", C.c_uint32), ("topology_out_of", C.c_uint32),
                ("topology_path", C.c_char_p), ("engine", C.c_char_p)]


EXPORTS = ["rs_model_args", "rs_model_info", "rs_config_defaults", "rs_host_lp_seeds", "rs_dev_init", "rs_dev_run", "rs_dev_stats",
           "rs_dev_trace", "rs_dev_lp_state", "rs_dev_set_comm", "rs_nccl_unique_id", "rs_dev_comm_init_nccl", "rs_nccl_comm_create", "rs_dev_set_nccl_comm",
           "rs_nccl_comm_destroy", "rs_host_lp_block",
           "rs_dev_kernel_times", "rs_dev_probe_event", "rs_dev_ipc_export", "rs_dev_ipc_import", "rs_set_debug", "rs_dev_error", "rs_dev_fini", "rs_model_file"]


def load(path):
    lib = C.CDLL(path)
    lib.rs_model_info.argtypes = [C.POINTER(ModelInfo)]
    lib.rs_model_info.restype = None
    lib.rs_model_args.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.rs_model_args.restype = C.c_int
    lib.rs_model_file.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    lib.rs_model_file.restype = C.c_int
    lib.rs_config_defaults.argtypes = [C.POINTER(Config)]
    lib.rs_config_defaults.restype = None
    lib.rs_host_lp_seeds.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.rs_host_lp_seeds.restype = None
    lib.rs_dev_init.argtypes = [C.POINTER(Config), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    lib.rs_dev_init.restype = C.c_int
    lib.rs_dev_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    lib.rs_dev_run.restype = C.c_int
    lib.rs_dev_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.rs_dev_stats.restype = C.c_int
    lib.rs_dev_trace.argtypes = [C.c_void_p, C.POINTER(LpTrace), C.c_uint32]
    lib.rs_dev_trace.restype = C.c_int
    lib.rs_dev_lp_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.rs_dev_lp_state.restype = C.c_int
    lib.rs_dev_set_comm.argtypes = [C.c_void_p, C.POINTER(CommOps)]
    lib.rs_dev_set_comm.restype = C.c_int
    lib.rs_nccl_unique_id.argtypes = [C.c_void_p, C.c_char_p]
    lib.rs_nccl_unique_id.restype = C.c_int
    lib.rs_dev_comm_init_nccl.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    lib.rs_dev_comm_init_nccl.restype = C.c_int
    lib.rs_nccl_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.rs_nccl_comm_create.restype = C.c_int
    lib.rs_dev_set_nccl_comm.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_dev_set_nccl_comm.restype = C.c_int
    lib.rs_nccl_comm_destroy.argtypes = [C.c_void_p]
    lib.rs_nccl_comm_destroy.restype = C.c_int
    lib.rs_host_lp_block.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rs_host_lp_block.restype = None
    lib.rs_dev_kernel_times.argtypes = [C.c_void_p, C.POINTER(KernelTime), C.c_int]
    lib.rs_dev_kernel_times.restype = C.c_int
    lib.rs_dev_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_dev_ipc_export.restype = C.c_int
    lib.rs_dev_ipc_import.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_dev_ipc_import.restype = C.c_int
    lib.rs_dev_probe_event.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double)]
    lib.rs_dev_probe_event.restype = C.c_int
    lib.rs_dev_error.argtypes = [C.c_void_p]
    lib.rs_dev_error.restype = C.c_char_p
    lib.rs_dev_fini.argtypes = [C.c_void_p]
    lib.rs_dev_fini.restype = C.c_int
    return lib


class Engine:
    """One device engine (one GPU, one block of LPs)."""

    def __init__(self, lib_path, private=False):
        """private=True loads a private copy of the library, so that the model's file-scope option
        variables (set through model_args) start from their defaults."""
        self._tmp = None
        if private:
            import shutil
            import tempfile
            fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix="rs_private_")
            os.close(fd)
            shutil.copy(lib_path, self._tmp)
            lib_path = self._tmp
        self.lib = load(lib_path)
        self.h = C.c_void_p()
        self.cfg = Config()
        self.lib.rs_config_defaults(C.byref(self.cfg))

    def info(self):
        mi = ModelInfo()
        self.lib.rs_model_info(C.byref(mi))
        return mi

    def model_args(self, args):
        """Pass model-specific options (e.g. ["--tau", "3"]) to the model's argp parser."""
        argv = [b"model"] + [a.encode() for a in args]
        arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
        rc = self.lib.rs_model_args(len(argv), arr)
        if rc != 0:
            raise RsError(2, "model rejected its options %r" % (args,))

    def model_file(self, name, data):
        """Hand the bytes of an input file the model opens inside INIT to the library (before init();
        data=None forgets the name).  Without it the engine looks for the file in the working directory."""
        rc = self.lib.rs_model_file(name.encode(), data, len(data) if data is not None else 0)
        if rc != 0:
            raise RsError(rc, "rs_model_file(%s)" % name)

    def init(self, host_seeds=None, **kw):
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        seeds = None
        if host_seeds is not None:
            seeds = host_seeds if isinstance(host_seeds, C.Array) else (C.c_uint64 * len(host_seeds))(*host_seeds)
        rc = self.lib.rs_dev_init(C.byref(self.cfg), seeds, C.byref(self.h))
        if rc != 0:
            msg = self.lib.rs_dev_error(self.h).decode() if self.h else ""
            self.fini()
            raise RsError(rc, msg)
        return self

    def run(self, max_windows=0):
        fin = C.c_int(0)
        rc = self.lib.rs_dev_run(self.h, max_windows, C.byref(fin))
        if rc != 0:
            raise RsError(rc, self.lib.rs_dev_error(self.h).decode())
        return bool(fin.value)

    def stats(self):
        st = Stats()
        self.lib.rs_dev_stats(self.h, C.byref(st))
        return st

    def trace(self, n=None):
        n = n or (self.cfg.lp_count or self.cfg.n_lps)
        arr = (LpTrace * n)()
        rc = self.lib.rs_dev_trace(self.h, arr, n)
        if rc != 0:
            raise RsError(rc, "trace")
        return arr

    def lp_state(self, nbytes, n=None):
        n = n or (self.cfg.lp_count or self.cfg.n_lps)
        buf = (C.c_ubyte * (nbytes * n))()
        rc = self.lib.rs_dev_lp_state(self.h, buf, nbytes, n)
        if rc != 0:
            raise RsError(rc, "lp_state")
        return bytes(buf)

    def set_comm(self, ops):
        """Install caller-provided collectives (tests: torch.distributed/gloo on host memory)."""
        self._comm = ops        # keep the callbacks alive
        rc = self.lib.rs_dev_set_comm(self.h, C.byref(ops))
        if rc != 0:
            raise RsError(rc, "set_comm")

    def comm_init_nccl(self, id128, libpath=None):
        rc = self.lib.rs_dev_comm_init_nccl(self.h, id128, libpath.encode() if libpath else None)
        if rc != 0:
            raise RsError(rc, self.lib.rs_dev_error(self.h).decode())

    def set_nccl_comm(self, comm):
        rc = self.lib.rs_dev_set_nccl_comm(self.h, comm)
        if rc != 0:
            raise RsError(rc, "set_nccl_comm")

    def ipc_export(self):
        """64-byte handle of this rank's inbox (direct NVLink exchange)"""
        buf = (C.c_ubyte * 64)()
        rc = self.lib.rs_dev_ipc_export(self.h, buf)
        if rc != 0:
            raise RsError(rc, self.lib.rs_dev_error(self.h).decode())
        return bytes(buf)

    def ipc_import(self, handles):
        """handles: the nranks exported handles in rank order"""
        if handles is None:
            buf = None
        else:
            blob = b"".join(handles)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        rc = self.lib.rs_dev_ipc_import(self.h, buf)
        if rc != 0:
            raise RsError(rc, self.lib.rs_dev_error(self.h).decode())

    def probe_event(self, lp, event_type, iters=10000, silent=False):
        """ns per handler call on one device thread (destructive diagnostic)"""
        ns = C.c_double(0)
        rc = self.lib.rs_dev_probe_event(self.h, lp, event_type, iters, 1 if silent else 0, C.byref(ns))
        if rc != 0:
            raise RsError(rc, self.lib.rs_dev_error(self.h).decode())
        return ns.value

    def kernel_times(self):
        arr = (KernelTime * 32)()
        n = self.lib.rs_dev_kernel_times(self.h, arr, 32)
        return {arr[i].name.decode(): (arr[i].total_ms, arr[i].launches) for i in range(n)}

    def fini(self):
        if self.h:
            self.lib.rs_dev_fini(self.h)
            self.h = C.c_void_p()
        if self._tmp:
            try:
                os.unlink(self._tmp)
            except OSError:
                pass
            self._tmp = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.fini()


def lp_block(lib, n_lps, nranks, rank):
    first, count = C.c_uint32(), C.c_uint32()
    lib.rs_host_lp_block(n_lps, nranks, rank, C.byref(first), C.byref(count))
    return first.value, count.value


def lp_seeds(lib, master_seed, lp_first, lp_count):
    arr = (C.c_uint64 * lp_count)()
    lib.rs_host_lp_seeds(master_seed, lp_first, lp_count, arr)
    return arr


def model_sources(name):
    """Sources of a model by name: the repo's own tests/models/<name>, else the reference's
    models/<name> when the reference tree is present (this container only)."""
    own = os.path.join(REPO, "tests", "models", name)
    ref = os.path.join(os.environ.get("RS_REFERENCE", "/root/reference"), "models", name)
    for d in (own, ref):
        if os.path.isdir(d):
            return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".c"))
    return []


def lib_path(name, emu=False):
    return os.path.join(BUILT, "lib%s%s.so" % (name, "_emu" if emu else ""))


def build_model(name, emu=False, force=False, extra=()):
    """Compile a model with rootsim-cc into root-sim_b200/_built/.  Returns the library path, or the
    prebuilt library when the sources are not available (GPU box: no /root/reference)."""
    out = lib_path(name, emu)
    srcs = model_sources(name)
    if not srcs:
        if os.path.exists(out):
            return out
        raise FileNotFoundError("no sources and no prebuilt library for model %r" % name)
    deps = srcs + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + \
        [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))] + \
        [ROOTSIM_CC, os.path.join(REPO, "include", "rootsim_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(BUILT, exist_ok=True)
    cmd = [sys.executable, ROOTSIM_CC, "--shared"] + (["--emu"] if emu else []) + list(extra) + srcs + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("rootsim-cc failed:\n" + r.stdout + r.stderr)
    return out
