"""oracle/solver.cpp's gradcal and rhscal (central compact path) against the independent NumPy restatement
tests/second_opinion_rhs.py on curvilinear grids: periodic, wall-bounded (ranges is:ie ...) and multi-block
(interface closures ntype 1 / 2, exchanged stress halos).  SURVEY.md 8c lists exactly these as not pinned by any
stored number of the reference."""
import numpy as np
import pytest

import second_opinion_rhs as R
from gpu_common import stretched_x

# measured: <= 9.5e-15 (continuity, whose scale is 1e-2 of the others), <= 3.5e-15 for the other fields
TOL = 1e-13
GAMMA, PRANDTL = 1.4, 0.72


def _case(oracle, n, homo, blocks=(1, 1, 1), reynolds=1600.0, mach=0.1):
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=reynolds, mach=mach)
    x = stretched_x(n, homo)
    if blocks == (1, 1, 1):
        c.set_x(x)
    else:      # every block gets its part of the global grid (blocks share their end nodes)
        for ib in range(c.nblocks):
            info = c.block_info(ib)
            g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
            c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom()
    c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 1e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar(); c.qswap(); c.zero_qrhs(); c.gradcal(); c.rhscal()
    th = dict(tempconst=110.3 / 273.15, reynolds=reynolds, prandtl=PRANDTL, const5=(GAMMA - 1.0) * mach ** 2)
    return c, th


def _check(c, th, homo):
    blocks = [R.Fields(c, ib) for ib in range(c.nblocks)]
    worst = 0.0
    for ib, F in enumerate(blocks):
        dvel, dtmp = R.gradcal(F)
        ref = [[R.core(c.get(f"dvel{a + 1}{b + 1}", ib)) for b in range(3)] for a in range(3)]
        scale = max(np.abs(r).max() for row in ref for r in row)
        worst = max(worst, max(np.abs(dvel[a][b] - ref[a][b]).max() for a in range(3) for b in range(3)) / scale)
        reft = [R.core(c.get(f"dtmp{a + 1}", ib)) for a in range(3)]
        scale = max(np.abs(r).max() for r in reft)
        worst = max(worst, max(np.abs(dtmp[a] - reft[a]).max() for a in range(3)) / scale)
    got = R.rhscal_blocks(blocks, th, homo)
    for ib in range(c.nblocks):
        for m in range(5):
            ref = R.core(c.get(f"qrhs{m + 1}", ib))
            worst = max(worst, np.abs(got[ib][m] - ref).max() / np.abs(ref).max())
    return worst


@pytest.mark.parametrize("n,homo", [((16, 14, 12), (True, True, True)),
                                    ((16, 14, 12), (True, False, True)),       # channel-like: walls in j (ntype 4)
                                    ((14, 16, 12), (False, True, False))])
def test_rhs_single_block(oracle, n, homo):
    c, th = _case(oracle, n, homo)
    assert _check(c, th, homo) < TOL
    c.close()


@pytest.mark.parametrize("n,homo,blocks", [((28, 12, 12), (False, True, True), (2, 1, 1)),   # ntype 1 | 2 in i
                                           ((12, 28, 14), (True, False, True), (1, 2, 1)),
                                           ((24, 12, 24), (True, True, True), (2, 1, 2))])   # periodic pairs
def test_rhs_multi_block(oracle, n, homo, blocks):
    c, th = _case(oracle, n, homo, blocks)
    assert _check(c, th, homo) < TOL
    c.close()


@pytest.mark.parametrize("n,blocks", [((16, 20, 12), (1, 1, 1)), ((24, 28, 12), (2, 2, 1))])
def test_rhs_channel_with_bulk_forcing(oracle, n, blocks):
    """examples/Channel option set: walls in y (one block: ntype 4; split in y: ntype 1 | 2), grichan stretching,
    src_chan with the bulk velocity summed over blocks."""
    from gpu_common import channel_state, channel_x
    homo, lengths, force = (True, False, True), (2 * np.pi, 2.0, np.pi), (2.5e-3, 1e-4, 3e-4)
    reynolds, mach = 3000.0, 0.3
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=reynolds, mach=mach, lengths=lengths)
    c.set_bc((1, 1, 41, 41, 1, 1), (0, 0, 1.0, 1.0, 0, 0))
    c.set_flow(1, force)
    x = channel_x(n, lengths)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom()
    th = dict(tempconst=110.3 / 273.15, reynolds=reynolds, prandtl=PRANDTL, const5=(GAMMA - 1.0) * mach ** 2,
              gamma=GAMMA, mach=mach, const1=1.0 / (GAMMA * (GAMMA - 1.0) * mach ** 2), const2=GAMMA * mach ** 2)
    for ib in range(c.nblocks):
        channel_state(c, th, ib=ib)
    c.updatefvar(); c.qswap(); c.zero_qrhs(); c.gradcal(); c.rhscal()
    blk = [R.Fields(c, ib) for ib in range(c.nblocks)]
    got = R.rhscal_blocks(blk, th, homo)
    src = R.src_chan(blk, [c.get("x2", ib) for ib in range(c.nblocks)], force)
    for ib in range(c.nblocks):
        for m in range(5):
            ref = R.core(c.get(f"qrhs{m + 1}", ib))
            # measured 5.6e-14 / 5.9e-14
            assert np.abs(got[ib][m] + src[ib][m] - ref).max() <= 1e-12 * np.abs(ref).max(), (ib, m)
        # the forcing itself is visible: without it the x-momentum row is off by force * J
        ref2 = R.core(c.get("qrhs2", ib))
        assert np.abs(got[ib][1] - ref2).max() > 1e-4 * np.abs(src[ib][1]).max()
    # massfluxchan / fbcxchan sums
    mf, fb = R.channel_sums(blk, [c.get("x2", ib) for ib in range(c.nblocks)], th, blocks[1])
    want = c.reduce(2)
    assert abs(mf - want[0]) <= 1e-13 * abs(want[0]) and abs(fb - want[1]) <= 1e-12 * abs(want[1]), (mf, fb, want)
    c.close()


def test_rhs_dimensional_gas(oracle):
    """nondimen = .false. (the HBL input): SI Sutherland law, hcc = cp miu / Pr, rgas = 287.1
    (src/solver.F90:128-139, src/fludyna.F90:806-809, src/solver.F90:2497-2501)."""
    from gpu_common import dimensional_state
    from astr_b200 import refcal_dimensional
    n, homo = (16, 14, 12), (True, True, True)
    ref = (226.65, 900.0, 1.0, 0.0180119)              # examples/Hypersonic_Boundary_Layer/datin/input.M3
    c = oracle.Case(*n, homo=homo, deltat=1e-5)
    c.set_dimensional(*ref)
    c.set_x(stretched_x(n, homo))
    c.gridgeom()
    dimensional_state(c, refcal_dimensional(*ref))
    c.updatefvar(); c.qswap(); c.zero_qrhs(); c.gradcal(); c.rhscal()
    rgas = 287.1
    th = dict(dimensional=True, prandtl=PRANDTL, cp=GAMMA / (GAMMA - 1.0) * rgas)
    F = R.Fields(c)
    got = R.rhscal(F, th, homo)
    for m in range(5):
        want = R.core(c.get(f"qrhs{m + 1}"))
        assert np.abs(got[m] - want).max() <= TOL * np.abs(want).max(), m
    # the viscous part is visible at this Reynolds number: without it the energy row moves by far more than TOL
    inviscid = R.rhscal(F, th, homo, diffterm=False)
    assert np.abs(inviscid[4] - R.core(c.get("qrhs5"))).max() > 1e-6 * np.abs(R.core(c.get("qrhs5"))).max()
    c.close()


@pytest.mark.parametrize("n,blocks", [((16, 14, 12), (1, 1, 1)), ((24, 14, 24), (2, 1, 2))])
def test_step_diagnostics(oracle, n, blocks):
    """kenergycal / enstophycal / diss_rate_cal sums and the CFL maxima."""
    homo = (True, True, True)
    c, th = _case(oracle, n, homo, blocks)
    th["mach"] = 0.1
    blk = [R.Fields(c, ib) for ib in range(c.nblocks)]
    got, want = R.tgv_sums(blk, th), c.reduce(0)
    assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want)), (got, want)
    got, want = R.cfl_maxima(blk, th), c.reduce(1)
    assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want)), (got, want)
    c.close()
