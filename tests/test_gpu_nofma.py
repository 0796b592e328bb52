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


# (case, bound of the production build = the tolerance tests/test_gpu_parity.py uses, bound of the no-contraction
# build).  Measured on a B200 (round 2): rhscal of every case collapses to <= 8e-14 without contraction -- 2.4e-15,
# i.e. the last bit, for the explicit WENO / WENO-Z / ROUND reconstructions -- so the production build's 4.6e-12
# (WENO-Z) and 1.3e-12 (device gridgeom -> 7e-14) are FMA contraction amplified by stiff nonlinear weights / nested
# derivatives.  What does NOT collapse is the state after several stages of the WENO-Z case (3.5e-12 -> 2.3e-12):
# there the re-associated line solves (partitioned Thomas) feed the same stiff weights.
CASES = [
    ("upwind_periodic:rhscal", 2e-12, 2e-13),
    ("upwind_walls:rhscal", 2e-12, 2e-13),
    ("upwind_periodic:3stages", 1e-12, 1e-12),
    ("upwind_walls:3stages", 1e-12, 1e-12),
    ("upwind_explicit:rhscal", 1e-11, 1e-11),
    ("upwind288:rhscal", 1e-11, 2e-13),
    ("upwind288:3stages", 1e-11, 5e-12),
    ("explicit_recon1:rhscal", 2e-11, 2e-14),
    ("explicit_recon2:rhscal", 2e-11, 2e-14),
    ("explicit_recon6:rhscal", 2e-11, 2e-14),
    ("explicit_recon2:2stages", 1e-11, 1e-11),
    ("device_gridgeom", 5e-12, 2e-13),
    ("golden:ke", 1e-12, 1e-12),
    ("golden:enstrophy", 2e-12, 2e-12),
]


@pytest.mark.parametrize("case,tol_fma,tol_nofma", CASES)
def test_loosened_cases_without_contraction(errors, case, tol_fma, tol_nofma):
    assert errors["fma"][case] <= tol_fma, (case, errors["fma"][case])
    assert errors["nofma"][case] <= tol_nofma, (case, errors["nofma"][case])
