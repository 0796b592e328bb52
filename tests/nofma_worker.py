"""Runs the parity cases whose tolerance exceeds north_star's 1e-12 in the production build against the
oracle with ONE library (argv[1]: "fma" = astr_b200/libastr_gpu.so, "nofma" = libastr_gpu_nofma.so, built with
-fmad=false -DASTR_NO_FMA) and prints one JSON object {case: max relative error}.  A separate process because a
process binds exactly one build of the library.  TEST INFRASTRUCTURE (tests/test_gpu_nofma.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import astr_b200  # noqa: E402
from astr_b200 import lib as L  # noqa: E402

if sys.argv[1] == "nofma":
    L.use_debug_library(L.build_nofma())
import pyoracle  # noqa: E402
from gpu_common import PRIMS, QS, assert_fields_close, core, make_pair  # noqa: E402

QRHS = [f"qrhs{n + 1}" for n in range(5)]
out = {}


def rhs_case(name, steps=0, **kw):
    c, eng = make_pair(pyoracle, **kw)
    c.qswap(); eng.qswap(); c.gradcal(); eng.gradcal()
    c.zero_qrhs(); c.rhscal(); eng.rhscal()
    out[name + ":rhscal"] = max(assert_fields_close(c, eng, QRHS, 1.0, what=name).values())
    if steps:
        for rk in range(1, steps + 1):
            c.rk_stage(rk); eng.rk_stage(rk)
        out[name + f":{steps}stages"] = max(assert_fields_close(c, eng, QS + PRIMS, 1.0, what=name).values())
    eng.close(); c.close()


# compact upwind path, UPWIND_TOL = 2e-11 in tests/test_gpu_parity.py
rhs_case("upwind_periodic", steps=3, n=(32, 36, 40), homo=(True, True, True), stretch="skew", perturb=1e-2,
         upwind=dict(lchardecomp=True, shkcrt="auto"))
rhs_case("upwind_walls", steps=3, n=(36, 32, 40), homo=(False, False, True), stretch=True, perturb=1e-2,
         upwind=dict(lchardecomp=True, shkcrt="auto"))
rhs_case("upwind_explicit", n=(32, 32, 32), homo=(True, True, False), stretch=True, perturb=1e-2, explicit=True,
         upwind=dict(lchardecomp=True, shkcrt="auto"))
# long, fine-in-i lines: 2e-9 there
rhs_case("upwind288", steps=3, n=(288, 16, 12), homo=(True, True, True), stretch="skew", perturb=1e-2,
         upwind=dict(lchardecomp=True, shkcrt="auto"))
# explicit upwind family: WENO (1), WENO-Z (2), ROUND (6) at 2e-10 / 1e-11
for rs in (1, 2, 6):
    rhs_case(f"explicit_recon{rs}", steps=2, n=(36, 32, 24), homo=(False, True, True), stretch=True, perturb=1e-2,
             upwind=dict(lchardecomp=True, shkcrt="auto", recon_schem=rs))
# device gridgeom: 1e-11
worst = 0.0
for kw in (dict(n=(40, 32, 36), stretch=True), dict(n=(32, 48, 32), homo=(True, False, True), stretch=True)):
    c, eng = make_pair(pyoracle, device_metrics=True, **kw)
    for name in ["jacob"] + [f"dxi{a + 1}{b + 1}" for a in range(3) for b in range(3)]:
        got, ref = eng.get(name), c.get(name)
        scale = np.abs(c.get("dxi11")).max() if name != "jacob" else np.abs(ref).max()
        worst = max(worst, float(np.abs(got - ref).max() / scale))
    eng.close(); c.close()
out["device_gridgeom"] = worst
# golden enstrophy history through the GPU path: 1e-11 (mini-app stage order, Sutherland 110.4)
golden = np.loadtxt(os.path.join(ROOT, "tests", "golden", "state.ref_128"))
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 40
c, eng = make_pair(pyoracle, n=(128, 128, 128), perturb=0.0, sutherland_s=110.4)
hist = []
for step in range(rows):
    for rk in (1, 2, 3):
        eng.qswap(); eng.gradcal()
        if rk == 1:
            hist.append(eng.tgv_stats())
        eng.rhscal(); eng.rk_update(rk); eng.filterq(); eng.updatefvar()
hist = np.array(hist)
out["golden:ke"] = float(np.abs(hist[:, 0] - golden[:rows, 2]).max() / golden[0, 2])
out["golden:enstrophy"] = float(np.abs(hist[:, 1] - golden[:rows, 3]).max() / golden[0, 3])
eng.close(); c.close()
print("NOFMA_JSON " + json.dumps(out))
