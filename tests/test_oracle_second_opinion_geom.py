"""oracle/solver.cpp's `gridgeom` (Jacobian and the nine conservative-form metrics, with their halo exchanges and
shared-node averages) against the independent NumPy restatement tests/second_opinion_geom.py, which evaluates the
Levi-Civita identity behind the reference's 18 hand-written terms.  SURVEY.md 8c: "curvilinear metrics" are pinned by
no stored number of the reference."""
import numpy as np
import pytest

import second_opinion_geom as G
import second_opinion_rhs as R
from gpu_common import stretched_x

HM = 5
# measured: 6e-15 periodic, 1e-14 multi-block, 6e-14 .. 9e-14 with wall closures (two nested one-sided derivative levels)
TOL = 1e-12


@pytest.mark.parametrize("n,blocks,homo", [((16, 14, 12), (1, 1, 1), (True, True, True)),
                                           ((16, 14, 12), (1, 1, 1), (True, False, True)),
                                           ((24, 14, 24), (2, 1, 2), (True, True, True)),
                                           ((28, 12, 12), (2, 1, 1), (False, True, True)),
                                           ((24, 20, 0), (1, 1, 1), (False, False, True)),       # ka = 0: 2-D metrics
                                           ((32, 24, 0), (2, 2, 1), (True, False, True))])
def test_gridgeom(oracle, n, blocks, homo):
    c = oracle.Case(*n, blocks=blocks, homo=homo)
    x = stretched_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom()
    blk = [R.Fields(c, ib) for ib in range(c.nblocks)]
    xs = [[c.get(f"x{m + 1}", ib) for m in range(3)] for ib in range(c.nblocks)]
    jac, dxi = G.gridgeom(xs, blk, homo)
    for ib, F in enumerate(blk):
        ref = R.core(c.get("jacob", ib))
        assert np.abs(R.core(jac[ib]) - ref).max() <= TOL * np.abs(ref).max()
        scale = max(np.abs(R.core(F.dxi[a][m])).max() for a in range(3) for m in range(3))
        for a in range(3):
            for m in range(3):
                assert np.abs(R.core(dxi[ib][a][m]) - R.core(F.dxi[a][m])).max() <= TOL * scale, (ib, a, m)
                for ax in range(3):       # halos of directions with a neighbour (or periodic onto itself)
                    lo, hi = F.nb[2 * ax], F.nb[2 * ax + 1]
                    if lo < 0 and hi < 0 and homo[ax]:
                        lo = hi = ib
                    idx = [slice(HM, -HM)] * 3
                    for present, sl in ((lo >= 0, slice(0, HM)), (hi >= 0, slice(-HM, None))):
                        if present:
                            idx[ax] = sl
                            d = dxi[ib][a][m][tuple(idx)] - F.dxi[a][m][tuple(idx)]
                            assert np.abs(d).max() <= TOL * scale, (ib, a, m, ax)
    c.close()
