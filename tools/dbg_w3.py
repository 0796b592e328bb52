import os, sys
os.environ["ASTR_SWEEP_W3"] = os.environ.get("W3", "1")
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np, pyoracle
from gpu_common import *
DVEL = [f"dvel{m + 1}{n + 1}" for m in range(3) for n in range(3)]
DTMP = [f"dtmp{n + 1}" for n in range(3)]
QRHS = [f"qrhs{n + 1}" for n in range(5)]
def worst(c, eng, names):
    out = {}
    for nm in names:
        r = core(c.get(nm)); g = core(eng.get(nm))
        d = np.abs(g - r); prof = d.max(axis=(1, 2))
        out[nm] = (float(d.max() / max(np.abs(r).max(), 1e-300)), int(prof.argmax()))
    return out
c, eng = make_pair(pyoracle, n=(300, 16, 16))
c.filterq(); eng.filterq(); print("filterq", worst(c, eng, QS))
c.qswap(); eng.qswap(); c.gradcal(); eng.gradcal(); print("gradcal", worst(c, eng, ["dvel11", "dvel21", "dvel31", "dtmp1", "dvel12"]))
c.zero_qrhs(); c.rhscal(); eng.rhscal(); print("rhscal", worst(c, eng, QRHS))
