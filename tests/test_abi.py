"""The product library (nvcc, sm_100a) loads, exports every symbol of include/rootsim_b200.h, and
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import rs_host
from conftest import have_gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(REPO, "include", "rootsim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert sorted(rs_host.EXPORTS) == declared_symbols()


def test_product_library_exports_the_abi():
    lib = ctypes.CDLL(rs_host.build_model("mixnet"))
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    mi = rs_host.Engine(rs_host.lib_path("mixnet")).info()
    assert mi.engine == b"cuda-sm100a" and mi.has_topology == 1


def test_no_cpu_fallback():
    if have_gpu():
        pytest.skip("GPU present")
    eng = rs_host.Engine(rs_host.build_model("mixnet"))
    with pytest.raises(rs_host.RsError) as ei:
        eng.init(n_lps=16, master_seed=1, simulation_time=10.0)
    assert ei.value.code == 1      # RS_ERR_NO_DEVICE


def test_kernel_resources_of_the_product_library():
    """Static checks on the sm_100a cubin (cuobjdump --dump-resource-usage): one sm_100a image; the LP-visit kernels
    keep the register budgets the launch geometry was tuned for; k_route does not carry the per-thread contexts'
    shared memory (it did: 40,480 bytes, five blocks per SM -- VERDICT round 1)."""
    import re
    import shutil
    import subprocess
    lib = rs_host.lib_path("phold")
    if not os.path.exists(lib) or not shutil.which("cuobjdump"):
        pytest.skip("no product library / cuobjdump")
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        mm = re.match(r"_Z\d+([a-z_]+?)\d", m.group(1))            # _Z7k_route6rs_dev -> k_route
        res[mm.group(1) if mm else m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    assert res["k_route"][2] < 4096
    assert res["k_run"][0] <= 128 and res["k_run_chain"][0] <= 168
    for k in ("k_scatter", "k_hist", "k_commit_scatter", "k_arr_fill", "k_wake"):
        assert res[k][0] <= 64 and res[k][2] < 8192, (k, res[k])
