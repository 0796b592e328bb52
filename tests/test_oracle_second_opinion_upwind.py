"""oracle/upwind.hpp's `convrsdcmp` (543c: Steger-Warming split, compact upwind interface flux, Roe-averaged
characteristic projection, MP5 with the Ducros flags) against the independent NumPy restatement
tests/second_opinion_upwind.py, which computes the split fluxes from the eigen-decomposition of the flux Jacobian
and the left eigenvectors by inverting the right ones.  SURVEY.md 8c: "the whole upwind/shock path" is pinned by no
stored number of the reference.

Grids: the sheared lattice of tests/gpu_common.py (`skewed_x`), whose metric normals are O(1) everywhere.  On grids
where a metric component passes through zero the reference's pivot rule (`abs(var1) > 1.d-12` on the RAW metric, then
a division by the normalised component, src/solver.F90:2051-2052) makes the eigenvector matrices ill conditioned
(1e7 and more): an inverse by LU and the reference's closed form then differ by cond x eps (measured 1e-6 on
`stretched_x`) -- conditioning of the reference's formulation, not a transcription error, and the reason the GPU parity
grids are sanitised too (conditioning note)."""
import numpy as np
import pytest

import second_opinion_rhs as R
import second_opinion_upwind as U
from gpu_common import auto_shkcrt, clean_metrics, skewed_x

GAMMA = 1.4
# measured: 3e-14 .. 1.4e-13 (five 5x5 projections and a 5x5 inverse per interface)
TOL = 2e-12


def _grid(c, n, homo):
    x = skewed_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom()
    clean_metrics(c)


def _compare(c, lchardecomp, mach=0.1):
    shk = auto_shkcrt(c, 0.3) if lchardecomp else 0.01
    c.set_upwind(543, lchardecomp, 0.3, shk)
    c.qswap(); c.gradcal()
    if lchardecomp:
        c.ducrossensor()
    c.zero_qrhs(); c.convrsdcmp()
    worst, flagged = 0.0, []
    for ib in range(c.nblocks):
        F = R.Fields(c, ib)
        lsh = c.get("lshock", ib) if lchardecomp else np.ones(F.prs.shape)       # lshock not allocated: lsh = .true.
        flagged.append(R.core(lsh).mean())
        got = U.convrsdcmp(F, GAMMA, mach, lsh, lchardecomp, 0.3)
        for m in range(5):
            ref = R.core(c.get(f"qrhs{m + 1}", ib))
            worst = max(worst, np.abs(got[m] - ref).max() / np.abs(ref).max())
    return worst, flagged


@pytest.mark.parametrize("n,homo,blocks,lchardecomp", [
    ((16, 14, 12), (True, True, True), (1, 1, 1), True),
    ((16, 14, 12), (True, True, True), (1, 1, 1), False),
    ((16, 14, 12), (True, False, True), (1, 1, 1), True),        # walls in j: ntype 4 closures of the compact flux
    ((14, 16, 12), (False, True, False), (1, 1, 1), True),
    ((28, 12, 12), (False, True, True), (2, 1, 1), True),        # ntype 1 | 2: mplimiter's unlimited edge interfaces (Q8)
    ((24, 14, 24), (True, True, True), (2, 1, 2), True)])
def test_convrsdcmp(oracle, n, homo, blocks, lchardecomp):
    c = oracle.Case(*n, blocks=blocks, homo=homo)
    _grid(c, n, homo)
    c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 2e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar()
    worst, flagged = _compare(c, lchardecomp)
    if lchardecomp:          # both branches of MP5 (flagged / unflagged interfaces) are exercised
        assert 0.2 < min(flagged) and max(flagged) < 0.8, flagged
    assert worst < TOL
    c.close()


def test_convrsdcmp_with_supersonic_pockets(oracle):
    """Local Mach numbers beyond +-1 in the i and j directions: the full-flux branches of the Steger-Warming split."""
    n, homo, mach = (16, 14, 12), (True, True, True), 0.9
    c = oracle.Case(*n, homo=homo, mach=mach)
    _grid(c, n, homo)
    c.tgvini()
    X = [c.get(f"x{d + 1}") for d in range(3)]
    fields = dict(rho=1.0 + 0.1 * np.sin(X[0]) * np.cos(X[1]), tmp=1.0 + 0.05 * np.cos(X[1]) * np.sin(X[2]),
                  u=1.5 * np.sin(X[0]) * np.cos(X[1]) * np.cos(X[2]), v=-1.5 * np.cos(X[0]) * np.sin(X[1]) * np.cos(X[2]),
                  w=0.4 * np.sin(X[2]))
    for nm, a in fields.items():
        c.set(nm, np.asfortranarray(a))
    c.updateq(); c.updatefvar()
    F = R.Fields(c)
    for ax in (0, 1):
        uu = sum(F.dxi[ax][k] * F.vel[k] for k in range(3))
        mag = np.sqrt(sum(F.dxi[ax][k] ** 2 for k in range(3)))
        lmach = R.core(uu)/ (np.sqrt(R.core(F.tmp)) / mach * R.core(mag))
        assert (lmach >= 1.0).mean() > 0.01 and (lmach <= -1.0).mean() > 0.01
    worst, _ = _compare(c, True, mach)
    assert worst < TOL
    c.close()


@pytest.mark.parametrize("n,homo,blocks", [((16, 14, 12), (True, True, True), (1, 1, 1)),
                                           ((24, 14, 24), (True, True, True), (2, 1, 2)),
                                           ((28, 12, 12), (False, True, True), (2, 1, 1)),     # clamped at the walls
                                           ((16, 14, 12), (True, False, True), (1, 1, 1))])
def test_ducrossensor(oracle, n, homo, blocks):
    """ssf to rounding, lshock flag for flag (the threshold sits in the widest gap of the sensor values).  For a
    direction with both walls in one block (npdc 4) the reference neither clamps the index nor fills the halo of
    ssf (allocated, never set: src/commcal.F90:212); like the oracle this restatement reads zeros there."""
    c = oracle.Case(*n, blocks=blocks, homo=homo)
    _grid(c, n, homo)
    c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 2e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar()
    shk = auto_shkcrt(c, 0.3)
    c.set_upwind(543, True, 0.3, shk)
    c.qswap(); c.gradcal(); c.ducrossensor()
    blk = [R.Fields(c, ib) for ib in range(c.nblocks)]
    ssf = []
    for ib, F in enumerate(blk):
        dvel, _ = R.gradcal(F)
        s = U.ducros_ssf(F, dvel)
        ref = R.core(c.get("ssf", ib))
        assert np.abs(s - ref).max() <= 1e-13 * ref.max()
        full = np.zeros(F.prs.shape)
        R.core(full)[...] = s
        ssf.append(full)
    ssf = R.exchange_halos(ssf, blk, homo)
    for ib, F in enumerate(blk):
        ref = R.core(c.get("lshock", ib))
        assert 0.2 < ref.mean() < 0.8
        assert np.array_equal(U.ducros_flags(ssf[ib], F.npdc, shk), ref)
    c.close()
