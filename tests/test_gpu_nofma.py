"""Why some parity tolerances exceed north_star's 1e-12: the production library lets nvcc contract a*b+c into
FMAs (and the line solves use explicit FMAs), the oracle -- like the reference's golden run -- does not.  The
same sources built with -fmad=false -DASTR_NO_FMA (`make -C astr_b200/csrc nofma`) are run here on exactly
the cases whose tolerance is loosened in tests/test_gpu_parity.py.  What this shows, case by case, is recorded
in the assertions below: where the no-contraction build collapses to rounding level the looseness is contraction
amplified by the conditioning of the case; where it does not, the residual is re-association (partitioned line
solves, reduction trees), bounded here as well."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(which):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "nofma_worker.py"), which, "40"],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("NOFMA_JSON ")][-1]
    return json.loads(line[len("NOFMA_JSON "):])


@pytest.fixture(scope="module")
def errors():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return {"fma": _run("fma"), "nofma": _run("nofma")}


def test_report(errors):
    for k in sorted(errors["fma"]):
        print(f"{k:32s} fma {errors['fma'][k]:.3e}   nofma {errors['nofma'][k]:.3e}", file=sys.stderr)


# (case, bound of the production build as used in test_gpu_parity.py, bound of the no-contraction build)
CASES = [
    ("upwind_periodic:rhscal", 2e-11, 2e-12),
    ("upwind_walls:rhscal", 2e-11, 2e-12),
    ("upwind_periodic:3stages", 1e-12, 1e-12),
    ("upwind_walls:3stages", 1e-12, 1e-12),
    ("upwind288:rhscal", 2e-9, 2e-10),
    ("explicit_recon1:rhscal", 2e-10, 2e-11),
    ("explicit_recon2:rhscal", 2e-10, 2e-11),
    ("explicit_recon6:rhscal", 2e-10, 2e-11),
    ("device_gridgeom", 1e-11, 1e-11),
    ("golden:ke", 1e-12, 1e-12),
    ("golden:enstrophy", 1e-11, 1e-11),
]


@pytest.mark.parametrize("case,tol_fma,tol_nofma", CASES)
def test_loosened_cases_without_contraction(errors, case, tol_fma, tol_nofma):
    assert errors["fma"][case] <= tol_fma, (case, errors["fma"][case])
    assert errors["nofma"][case] <= tol_nofma, (case, errors["nofma"][case])
