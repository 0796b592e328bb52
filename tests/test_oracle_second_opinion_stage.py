"""Three Runge-Kutta stages of oracle/solver.cpp (filterq -> qswap -> gradcal -> rhscal -> RK3 -> updatefvar) against
the independent NumPy restatement tests/second_opinion_stage.py on curvilinear periodic block grids: pins the stage
ORDER, the copy-vs-average exchange semantics (quirk Q3) and the pre-/post-filter primitive split (quirk Q2) of the
main-solver mode, which the reference's golden vector (mini-app order) does not cover."""
import numpy as np
import pytest

import second_opinion_rhs as R
import second_opinion_stage as S
from gpu_common import stretched_x

GAMMA, PRANDTL, REYNOLDS, MACH, DT = 1.4, 0.72, 1600.0, 0.1, 1e-3
# measured 3.5e-14 .. 5.5e-14 after each of the three stages (the filter at alfa = 0.49 is the least well conditioned
# operator: dense LU and Thomas differ by ~3e-15 per application)
TOL = 5e-13
NAMES = [f"q{m + 1}" for m in range(5)] + ["rho", "u", "v", "w", "prs", "tmp"]


@pytest.mark.parametrize("n,blocks", [((16, 14, 12), (1, 1, 1)), ((24, 14, 24), (2, 1, 2)), ((14, 26, 12), (1, 2, 1))])
def test_three_stages(oracle, n, blocks):
    homo = (True, True, True)
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=REYNOLDS, mach=MACH, deltat=DT)
    x = stretched_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 1e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar()
    th = dict(tempconst=110.3 / 273.15, reynolds=REYNOLDS, prandtl=PRANDTL, const5=(GAMMA - 1.0) * MACH ** 2,
              const6=1.0 / (GAMMA - 1.0), const2=GAMMA * MACH ** 2)
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    qsave = [None] * c.nblocks
    for rk in (1, 2, 3):
        c.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, DT, qsave)
        for ib, F in enumerate(state):
            got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
            ref = {nm: R.core(c.get(nm, ib)) for nm in NAMES}
            for grp in (["q2", "q3", "q4"], ["u", "v", "w"]):        # vector components share the vector's scale
                s = max(np.abs(ref[nm]).max() for nm in grp)
                for nm in grp:
                    assert np.abs(R.core(got[nm]) - ref[nm]).max() <= TOL * s, (rk, ib, nm)
            for nm in ("q1", "q5", "rho", "prs", "tmp"):
                assert np.abs(R.core(got[nm]) - ref[nm]).max() <= TOL * np.abs(ref[nm]).max(), (rk, ib, nm)
    c.close()


@pytest.mark.parametrize("n,blocks", [((16, 20, 12), (1, 1, 1)), ((24, 28, 12), (2, 2, 1))])
def test_three_stages_of_the_channel(oracle, n, blocks):
    """examples/Channel option set (config 2): isothermal no-slip walls 41 at jmin / jmax through boucon, wall
    closures of filter and derivative (one block: ntype 4; split in y: ntype 1 | 2), src_chan."""
    from gpu_common import channel_state, channel_x
    homo, lengths, force = (True, False, True), (2 * np.pi, 2.0, np.pi), (2.5e-3, 0.0, 1e-4)
    bctype, twall = (1, 1, 41, 41, 1, 1), (0.0, 0.0, 1.0, 1.0, 0.0, 0.0)
    reynolds, mach = 3000.0, 0.3
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=reynolds, mach=mach, lengths=lengths, deltat=DT)
    c.set_bc(bctype, twall)
    c.set_flow(1, force)
    x = channel_x(n, lengths)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom()
    th = dict(tempconst=110.3 / 273.15, reynolds=reynolds, prandtl=PRANDTL, const5=(GAMMA - 1.0) * mach ** 2,
              const6=1.0 / (GAMMA - 1.0), const2=GAMMA * mach ** 2, gamma=GAMMA, mach=mach,
              const1=1.0 / (GAMMA * (GAMMA - 1.0) * mach ** 2))
    for ib in range(c.nblocks):
        channel_state(c, th, ib=ib)
    c.updatefvar()
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    ys = [c.get("x2", ib) for ib in range(c.nblocks)]
    qsave = [None] * c.nblocks
    for rk in (1, 2, 3):
        c.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, DT, qsave, bctype=bctype, twall=twall, force=force, ys=ys)
        for ib, F in enumerate(state):
            got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
            ref = {nm: R.core(c.get(nm, ib)) for nm in NAMES}
            for grp in (["q2", "q3", "q4"], ["u", "v", "w"]):
                s = max(np.abs(ref[nm]).max() for nm in grp)
                for nm in grp:
                    assert np.abs(R.core(got[nm]) - ref[nm]).max() <= TOL * s, (rk, ib, nm)
            for nm in ("q1", "q5", "rho", "prs", "tmp"):
                assert np.abs(R.core(got[nm]) - ref[nm]).max() <= TOL * np.abs(ref[nm]).max(), (rk, ib, nm)
    c.close()


@pytest.mark.parametrize("n,blocks", [((16, 14, 10), (1, 1, 1)), ((24, 16, 10), (2, 2, 1)),
                                      ((24, 20, 0), (1, 1, 1)), ((32, 24, 0), (2, 2, 1))])      # ka = 0: true 2-D blocks
def test_three_stages_of_a_swbli_like_case(oracle, n, blocks):
    """The SWLBI option set (config 4) in 3-D: conschm 543c with characteristic MP5 and the Ducros sensor, inflow 11 at
    imin, outflow 21 at imax, slip wall 421 at jmin, far field 51 at jmax, a sponge layer at imax, periodic in k (3-D) or
    ka = 0 (the input's own 2-D mode).  Inside a stage the split
    fluxes see pre-filter primitives next to post-filter q (quirk Q2): this is the test that exercises the reference's
    mixed use of the two in flux_steger_warming (tests/second_opinion_upwind.py::steger_warming)."""
    import second_opinion_bc  # noqa: F401  (used through S.rk_stage)
    from gpu_common import auto_shkcrt, clean_metrics, skewed_x
    homo, bctype, twall, mach = (False, False, True), (11, 21, 421, 51, 1, 1), (0.0,) * 6, 0.5
    c = oracle.Case(*n, homo=homo, blocks=blocks, mach=mach, reynolds=REYNOLDS, deltat=DT)
    c.set_bc(bctype, twall)
    x = skewed_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); clean_metrics(c); c.tgvini()
    th = dict(mach=mach, gamma=GAMMA, const1=1.0 / (GAMMA * (GAMMA - 1.0) * mach ** 2), const2=GAMMA * mach ** 2,
              const6=1.0 / (GAMMA - 1.0), const5=(GAMMA - 1.0) * mach ** 2, tempconst=110.3 / 273.15,
              reynolds=REYNOLDS, prandtl=PRANDTL)
    inflow = []
    for ib in range(c.nblocks):
        X = [c.get(f"x{d + 1}", ib) for d in range(3)]
        fields = dict(rho=1.0 + 0.1 * np.sin(X[0]) * np.cos(X[1]), tmp=1.0 + 0.05 * np.cos(X[1]) * np.sin(X[2] + 0.3),
                      u=1.0 + 0.3 * np.sin(X[1]) * np.cos(X[2]), v=0.3 * np.sin(X[0] + 0.4) * np.cos(X[2]),
                      w=0.2 * np.sin(X[0]) * np.cos(X[1] + 0.2))
        for nm, a in fields.items():
            c.set(nm, np.asfortranarray(a), ib)
        info = c.block_info(ib)
        jm, km = info["jm"], info["km"]
        yy = (np.arange(jm + 1) + info["g0"][1]) / n[1]
        vel_in = np.zeros((jm + 1, km + 1, 3), order="F")
        vel_in[:, :, 0] = (0.6 + 1.0 * yy)[:, None]
        vel_in[:, :, 2] = 0.02
        tmp_in = np.asfortranarray(1.0 + 0.05 * yy[:, None] * np.ones((1, km + 1)))
        tmp_prof = 1.0 + 0.05 * yy
        c.set_inflow(vel_in, tmp_in, tmp_prof, ib)
        inflow.append((vel_in, tmp_in, tmp_prof))
    c.updateq(); c.updatefvar()
    shk = auto_shkcrt(c, 0.3)
    c.set_upwind(543, True, 0.3, shk)
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    # examples/SWLBI/datin/input.2d has a sponge layer at imax: 8 nodes deep on the blocks that own the face
    layers, rng = {1: []}, np.random.default_rng(11)
    for ib, F in enumerate(state):
        info = c.block_info(ib)
        if info["g0"][0] + info["im"] == n[0]:
            beg, end = info["im"] - 8, F.hi[0]
            coef = 0.05 * rng.random((end - beg + 1, F.hi[1] - F.lo[1] + 1, F.hi[2] - F.lo[2] + 1))
            layers[1].append((beg, end, coef))
            c.set_sponge(1, beg, end, np.asfortranarray(coef), ib)
        else:
            layers[1].append(None)
            c.set_sponge(1, -1, -1, None, ib)
    qsave = [None] * c.nblocks
    extra = dict(free=(1.0, 0.0, 0.0, 1.0, c.pinf), inflow_data=inflow)
    for rk in (1, 2, 3):
        c.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, DT, qsave, bctype=bctype, twall=twall,
                   upwind=dict(shkcrt=shk, lchardecomp=True, bfacmpld=0.3), bc_extra=extra,
                   sponge=lambda blk: S.spongefilter_layer(blk, homo, layers))
        for ib, F in enumerate(state):
            got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
            for nm in NAMES:
                ref = R.core(c.get(nm, ib))
                # measured <= 5.2e-14; velocities / momenta relative to the free-stream speed 1
                assert np.abs(R.core(got[nm]) - ref).max() <= TOL * max(np.abs(ref).max(), 0.3), (rk, ib, nm)
    c.close()


@pytest.mark.parametrize("n,blocks,bctype,homo", [
    ((16, 14, 10), (1, 1, 1), (11, 21, 41, 51, 1, 1), (False, False, True)),
    ((24, 16, 10), (2, 2, 1), (11, 21, 41, 51, 1, 1), (False, False, True)),
    ((24, 20, 0), (1, 1, 1), (11, 21, 41, 51, 1, 1), (False, False, True)),       # ka = 0: the input's own 2-D mode
    ((32, 24, 0), (2, 2, 1), (11, 21, 41, 51, 1, 1), (False, False, True)),
    # far field with the dimensional free stream (uinf = ref_vel, roinf = ref_den, src/solver.F90:131-139) on the
    # faces whose characteristic branches use it, outflow at jmax
    ((16, 14, 10), (1, 1, 1), (11, 21, 51, 21, 51, 51), (False, False, False))])
def test_three_stages_of_an_hbl_like_case(oracle, n, blocks, bctype, homo):
    """The HBL option set (config 3, examples/Hypersonic_Boundary_Layer/datin/input.M3) in 3-D: dimensional gas
    (nondimen = f), conschm 543c + difschm 642e (fds%central is the explicit ladder in gradcal and diffrsdcal6), no
    filter, inflow 11 / outflow 21 / isothermal wall 41 (568.89 K) / far field 51."""
    from gpu_common import auto_shkcrt, clean_metrics, skewed_x
    ref = (226.65, 900.0, 1.0, 0.0180119)
    dt, twall = 1e-6, (0, 0, 568.89, 0, 0, 0)
    c = oracle.Case(*n, homo=homo, blocks=blocks, deltat=dt)
    c.set_dimensional(*ref)
    c.set_flags(lfilter=False, diffterm=True)
    c.set_scheme(True)
    c.set_bc(bctype, twall)
    x = skewed_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); clean_metrics(c); c.tgvini()
    th = dict(dimensional=True, gamma=GAMMA, prandtl=PRANDTL, rgas=287.1, cp=GAMMA / (GAMMA - 1.0) * 287.1)
    inflow = []
    for ib in range(c.nblocks):
        X = [c.get(f"x{d + 1}", ib) for d in range(3)]
        fields = dict(rho=ref[3] * (1.0 + 0.1 * np.sin(X[0]) * np.cos(X[1])),
                      tmp=ref[0] * (1.0 + 0.05 * np.cos(X[1]) * np.sin(X[2] + 0.3)),
                      u=ref[1] * (0.6 + 0.2 * np.sin(X[1]) * np.cos(X[2])),
                      v=ref[1] * 0.1 * np.sin(X[0] + 0.4) * np.cos(X[2]),
                      w=ref[1] * 0.05 * np.sin(X[0]) * np.cos(X[1] + 0.2))
        for nm, a in fields.items():
            c.set(nm, np.asfortranarray(a), ib)
        info = c.block_info(ib)
        jm, km = info["jm"], info["km"]
        yy = (np.arange(jm + 1) + info["g0"][1]) / n[1]
        vel_in = np.zeros((jm + 1, km + 1, 3), order="F")
        vel_in[:, :, 0] = ref[1] * (0.2 + 0.8 * yy)[:, None]          # Mach 0.6 .. 3 against ~302 m/s
        vel_in[:, :, 2] = ref[1] * 0.01
        tmp_in = np.asfortranarray(ref[0] * (1.0 + 0.05 * yy[:, None] * np.ones((1, km + 1))))
        tmp_prof = ref[0] * (1.0 + 0.05 * yy)
        c.set_inflow(vel_in, tmp_in, tmp_prof, ib)
        inflow.append((vel_in, tmp_in, tmp_prof))
    c.updateq(); c.updatefvar()
    shk = auto_shkcrt(c, 0.1)
    c.set_upwind(543, True, 0.1, shk)
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    for F in state:
        F.explicit = True
    qsave = [None] * c.nblocks
    extra = dict(free=(ref[1], 0.0, 0.0, ref[3], c.pinf), inflow_data=inflow)
    scale = dict(u=ref[1], v=ref[1], w=ref[1], q2=ref[1] * ref[3], q3=ref[1] * ref[3], q4=ref[1] * ref[3])
    before = {nm: R.core(c.get(nm, 0)).copy() for nm in ("q1", "q5")}
    for rk in (1, 2, 3):
        c.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, dt, qsave, bctype=bctype, twall=twall,
                   upwind=dict(shkcrt=shk, lchardecomp=True, bfacmpld=0.1), bc_extra=extra, lfilter=False)
        for ib, F in enumerate(state):
            got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
            for nm in NAMES:
                want = R.core(c.get(nm, ib))
                # measured <= 5e-14
                assert np.abs(R.core(got[nm]) - want).max() <= TOL * scale.get(nm, np.abs(want).max()), (rk, ib, nm)
    # the step moved the state by far more than the tolerance (the comparison is not trivially satisfied)
    for nm in before:
        now = R.core(c.get(nm, 0))
        assert np.abs(now - before[nm]).max() > 1e-5 * np.abs(now).max()
    c.close()


def test_rk4_step(oracle):
    """rkscheme = 'rk4' (src/mainloop.F90:368-381, :452-476): four stages with the accumulated right-hand side."""
    n, blocks, homo = (16, 14, 12), (1, 1, 1), (True, True, True)
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=REYNOLDS, mach=MACH, deltat=DT)
    c.set_x(stretched_x(n, homo))
    c.gridgeom(); c.tgvini()
    c.set_rkscheme(4)
    rng = np.random.default_rng(5)
    for m in range(5):
        a = c.get(f"q{m + 1}")
        a *= 1.0 + 1e-2 * rng.standard_normal(a.shape)
        c.set(f"q{m + 1}", a)
    c.updatefvar()
    th = dict(tempconst=110.3 / 273.15, reynolds=REYNOLDS, prandtl=PRANDTL, const5=(GAMMA - 1.0) * MACH ** 2,
              const6=1.0 / (GAMMA - 1.0), const2=GAMMA * MACH ** 2)
    state = [R.Fields(c)]
    qsave, rhsav = [None], [None]
    for rk in (1, 2, 3, 4):
        c.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, DT, qsave, rk4=rhsav)
        got = dict(zip(NAMES, state[0].q + [state[0].rho] + state[0].vel + [state[0].prs, state[0].tmp]))
        for nm in NAMES:
            want = R.core(c.get(nm))
            assert np.abs(R.core(got[nm]) - want).max() <= TOL * max(np.abs(want).max(), 0.5), (rk, nm)
    c.close()


def _periodic_case(oracle, n, blocks):
    homo = (True, True, True)
    c = oracle.Case(*n, blocks=blocks, homo=homo, reynolds=REYNOLDS, mach=MACH, deltat=DT)
    x = stretched_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 1e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar()
    th = dict(tempconst=110.3 / 273.15, reynolds=REYNOLDS, prandtl=PRANDTL, const5=(GAMMA - 1.0) * MACH ** 2,
              const6=1.0 / (GAMMA - 1.0), const2=GAMMA * MACH ** 2)
    return c, th, homo


def _assert_state(c, state, what):
    for ib, F in enumerate(state):
        got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
        for nm in NAMES:
            want = R.core(c.get(nm, ib))
            assert np.abs(R.core(got[nm]) - want).max() <= TOL * max(np.abs(want).max(), 0.5), (what, ib, nm)


@pytest.mark.parametrize("form", ["layer", "circl"])
def test_three_stages_with_a_sponge(oracle, form):
    """spongefilter between the RK update and updatefvar (src/mainloop.F90:478; src/sponge_layer.F90:55-364): the layer
    form (imax and kmax layers, 6 nodes deep, on the blocks that own those faces) and the global 'circl' form with the
    coefficients of spongelayer_define_circle (maximum over blocks), 2 x 1 x 2 periodic blocks."""
    n, blocks = (24, 14, 24), (2, 1, 2)
    c, th, homo = _periodic_case(oracle, n, blocks)
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    if form == "layer":
        layers, rng = {1: [], 5: []}, np.random.default_rng(9)
        for ib, F in enumerate(state):
            info = c.block_info(ib)
            ext = [info["im"], info["jm"], info["km"]]
            for face in (1, 5):
                ax = face // 2
                if info["g0"][ax] + ext[ax] == n[ax]:          # the block at the high end of that direction
                    beg, end = ext[ax] - 6, F.hi[ax]
                    shape = [F.hi[a] - F.lo[a] + 1 for a in range(3)]
                    shape[ax] = end - beg + 1
                    coef = 0.05 * rng.random(shape)
                    layers[face].append((beg, end, coef))
                    c.set_sponge(face, beg, end, np.asfortranarray(coef), ib)
                else:
                    layers[face].append(None)
                    c.set_sponge(face, -1, -1, None, ib)
        sponge = lambda b: S.spongefilter_layer(b, homo, layers)
    else:
        centre, radius, dampfac = (3.0, 3.0, 3.0), 2.5, 0.05
        c.set_sponge_circle(centre, radius, dampfac)
        xs = [[c.get(f"x{m + 1}", ib) for m in range(3)] for ib in range(c.nblocks)]
        coefs = S.sponge_circle_coefficients(state, xs, centre, radius, dampfac)
        for ib in range(c.nblocks):
            np.testing.assert_allclose(coefs[ib], c.sponge_circle_coef(ib), rtol=1e-14, atol=0)
        assert max(co.max() for co in coefs) == dampfac
        sponge = lambda b: S.spongefilter_global(b, homo, coefs)
    # the same stages without the sponge end up somewhere else: the damping is visible at the tolerance
    c0, _, _ = _periodic_case(oracle, n, blocks)
    qsave = [None] * c.nblocks
    for rk in (1, 2, 3):
        c.rk_stage(rk); c0.rk_stage(rk)
        S.rk_stage(state, rk, th, homo, DT, qsave, sponge=sponge)
        _assert_state(c, state, f"{form} stage {rk}")
    assert max(np.abs(R.core(c.get("q2", ib)) - R.core(c0.get("q2", ib))).max() for ib in range(c.nblocks)) > 1e-6
    c.close(); c0.close()
