"""The C ABI driven from plain C (gcc, include/astr_gpu.h only -- no Python between the caller and
libastr_gpu.so), the way the Fortran bind(C) shim calls it; result compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi", "abi_driver.c")
EXE = os.path.join(ROOT, "tests", "abi", "abi_driver")


def build_driver():
    so_dir = os.path.join(ROOT, "astr_b200")
    cmd = ["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", so_dir, "-lastr_gpu", f"-Wl,-rpath,{so_dir}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_c_driver_compiles_against_the_header():
    import astr_b200
    astr_b200.build()
    build_driver()


@pytest.mark.gpu
def test_c_driver_matches_the_oracle(oracle, tmp_path):
    exe = build_driver()
    n, nsteps = 32, 2
    out = tmp_path / "q.bin"
    r = subprocess.run([exe, str(n), str(nsteps), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernel launches=" in r.stdout and int(r.stdout.split("kernel launches=")[1]) > 0
    m = n + 11
    q = np.fromfile(out).reshape((m, m, m, 5), order="F")
    c = oracle.Case(n, n, n)
    c.gridgeom(); c.tgvini(); c.run(nsteps)
    for k in range(5):
        ref = c.get(f"q{k + 1}")[5:-5, 5:-5, 5:-5]
        got = q[5:-5, 5:-5, 5:-5, k]
        scale = np.abs(c.get("q2")).max() if k in (1, 2, 3) else np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-12 * scale, k
    c.close()
