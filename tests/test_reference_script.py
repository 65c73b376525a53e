"""The reference's own acceptance script (scripts/tests.sh:124-137) restated for this runtime: every model it lists
that does not need the ABM layer is compiled with rootsim-cc and run with the script's own command lines -- to the
model's own termination, no device sizing on the command line -- and must exit 0 with the completion banner.  Where
the script runs `--sequential`, this runtime refuses (it is the optimistic engine; the sequential one is the
reference's).  CPU: thread-emulation build of the same unity source (the command line of the product executable is exercised on
the GPU by test_gpu_parity.py::test_cli_executable)."""
import os
import subprocess
import sys

import pytest

import rs_host

# scripts/tests.sh:131-134 (stupid_model needs src/lib/abm_layer.c: out of scope) and :137
SCRIPT = [
    ("pcs", ["--wt", "2", "--lp", "16"]),
    ("packet", ["--wt", "2", "--lp", "4"]),
    ("collector", ["--wt", "2", "--lp", "4"]),
    ("pcs", ["--wt", "2", "--lp", "16", "--output-dir", "dummy", "--npwd", "--gvt", "500", "--gvt-snapshot-cycles", "3", "--verbose", "info",
             "--seed", "12345", "--scheduler", "stf", "--cktrm-mode", "normal", "--simulation-time", "1000"]),
]


def build_exe(model, tmp_path, emu):
    srcs = rs_host.model_sources(model)
    if not srcs:
        pytest.skip("%s sources live in the reference tree" % model)
    exe = tmp_path / "model"
    subprocess.run([sys.executable, rs_host.ROOTSIM_CC] + (["--emu"] if emu else []) + srcs + ["-o", str(exe)], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("model,argv", SCRIPT, ids=["pcs", "packet", "collector", "pcs-cust"])
def test_reference_script_lines_emu(tmp_path, model, argv):
    exe = build_exe(model, tmp_path, emu=True)
    env = dict(os.environ, HOME=str(tmp_path))
    r = subprocess.run([str(exe)] + argv, capture_output=True, text=True, cwd=tmp_path, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "SIMULATION CORRECTLY COMPLETED" in r.stdout
    out = tmp_path / ("dummy" if "dummy" in argv else "outputs")
    stats = open(out / "execution_stats").read()
    committed = int([l for l in stats.split("\n") if l.startswith("TOTAL COMMITTED EVENTS")][0].split(":")[1])
    assert committed > 1000
    r = subprocess.run([str(exe), "--sequential", "--lp", "1"], capture_output=True, text=True, cwd=tmp_path, env=env)
    assert r.returncode != 0

