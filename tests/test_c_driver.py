"""A pure-C program (tests/c/abi_driver.c) linked against libcsic.so: the boundary is a plain C ABI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "chroma-subsampling-image-compressor_b200")
SRC = os.path.join(ROOT, "tests", "c", "abi_driver.c")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cdrv") / "abi_driver")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include", SRC, f"-L{PKG}", "-lcsic",
                           f"-Wl,-rpath,{PKG}", "-o", exe])
    return exe


def test_c_driver_host_checks(driver):
    r = subprocess.run([driver, "host"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host checks ok" in r.stdout


def fnv1a(buf):
    h = 1469598103934665603
    for b in bytes(buf):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,a,b,f,fmt", [(256, 32, 2, 0, 2, 3), (64, 16, 2, 0, 1, 0), (40, 12, 1, 1, 4, 1)])
def test_c_driver_gpu_matches_oracle(driver, W, H, a, b, f, fmt):
    import oracle
    r = subprocess.run([driver, "gpu", *map(str, (W, H, a, b, f, fmt))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    n = 3 * W * H * 3
    s, vals = 12345, np.empty(n, np.uint8)
    for i in range(n):                                   # the driver's LCG
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        vals[i] = s >> 24
    want = oracle.process(oracle.make_params(W, H, a, b, (6, 5, 5), f, "CSQ", out_format=fmt), vals.reshape(3, H, W, 3))
    assert f"fnv {fnv1a(want.tobytes()):016x}" in r.stdout, r.stdout
    assert "generic same" in r.stdout
