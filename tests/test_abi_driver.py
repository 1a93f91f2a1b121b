"""tests/abi_driver.c: the C ABI driven from plain C99 (gcc -Wall -Wextra -Werror against include/b2o.h), no Python between
the caller and libb2o.so.  CPU box: it must compile, link, load and be refused a context ("no CPU fallback").  GPU box
(-m gpu): diag -> index -> L-BFGS / inverse / L-SR1 push!+apply -> fused cfg3 tree -> kron, each checked against the oracle .so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "abi_driver")


def build_driver(orc):
    lib_dir, orc_dir = os.path.join(ROOT, "linearoperators.jl_b200"), os.path.join(ROOT, "oracle")
    from linearoperators_jl_b200 import _lib
    _lib.load()                                   # libb2o.so must exist
    cmd = ["/usr/bin/gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-o", EXE, os.path.join(ROOT, "tests", "abi_driver.c"),
           "-L" + lib_dir, "-lb2o", "-L" + orc_dir, "-lb2o_oracle", "-lm",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + orc_dir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def run_driver():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    return subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_abi_driver_builds_and_is_refused_without_gpu(lo, orc):
    import torch
    build_driver(orc)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = run_driver()
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_abi_driver_plain_c_against_oracle(lo, orc):
    build_driver(orc)
    r = run_driver()
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "ABI_DRIVER_OK" in r.stdout
