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
