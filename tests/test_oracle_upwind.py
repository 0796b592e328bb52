"""Building blocks of the upwind compact convection path in the oracle (oracle/upwind.hpp):
parity is UNPINNED by any stored number of the reference, so these tests hold the
restatement to the identities the reference's own self-tests use (src/solver.F90:2878-2944
`LEV*REV=I`; src/test.F90 `fluxtest`: convergence order) and to conservation."""
import numpy as np
import pytest

HM = 5


def _state(rng):
    g, M = 1.4, 0.1 + 2.0 * rng.random()
    rho = 0.8 + 0.4 * rng.random()
    vel = rng.standard_normal(3) * (1.5 if rng.random() < 0.5 else 0.3)
    T = 0.9 + 0.2 * rng.random()
    p = rho * T / (g * M * M)
    E = p / (g - 1.0) + 0.5 * rho * vel @ vel
    return g, M, rho, vel, T, p, E


def test_left_and_right_eigenvectors_are_inverse(oracle):
    rng = np.random.default_rng(7)
    for trial in range(40):
        g, M, rho, vel, T, p, E = _state(rng)
        ddi = rng.standard_normal(3)
        if trial % 4 == 1: ddi[0] = 0.0           # second pivot branch (solver.F90:2081)
        if trial % 4 == 2: ddi[0] = ddi[1] = 0.0  # third pivot branch (:2111)
        left = [rho, p, E, *vel, *ddi]
        g2, M2, rho2, vel2, T2, p2, E2 = _state(rng)
        right = [rho2, p2, E2, *vel2, *ddi]
        rev, lev = oracle.chardecomp(left, right, gamma=1.4)
        assert np.abs(lev @ rev - np.eye(5)).max() < 1e-8      # the reference's own tolerance
    with pytest.raises(ValueError):
        oracle.chardecomp([1, 1, 3, 0, 0, 0, 0, 0, 0], [1, 1, 3, 0, 0, 0, 0, 0, 0])


def test_steger_warming_split_sums_to_the_euler_flux(oracle):
    rng = np.random.default_rng(11)
    seen = set()
    for trial in range(200):
        g, M, rho, vel, T, p, E = _state(rng)
        q = [rho, *(rho * vel), E]
        dxi = rng.standard_normal(3) * 4.0
        jac = 0.01 + rng.random()
        fp, fm = oracle.steger_warming(rho, vel, p, T, q, dxi, jac, g, M)
        uu = dxi @ vel
        F = jac * np.array([rho * uu, q[1] * uu + dxi[0] * p, q[2] * uu + dxi[1] * p, q[3] * uu + dxi[2] * p, (E + p) * uu])
        assert np.abs(fp + fm - F).max() <= 1e-13 * np.abs(F).max()
        lmach = uu / (np.sqrt(T) / M * np.linalg.norm(dxi))
        if lmach >= 1.0:
            assert np.all(fm == 0.0); seen.add("+")
        elif lmach <= -1.0:
            assert np.all(fp == 0.0); seen.add("-")
        else:
            seen.add("0")
    assert seen == {"+", "-", "0"}


@pytest.mark.parametrize("plus", [True, False])
def test_compact_upwind_flux_is_fifth_order_and_telescopes(oracle, plus):
    errs = []
    for n in (32, 64):
        x = 2 * np.pi * np.arange(-HM, n + HM + 1) / n
        fh = oracle.flux_compact(np.sin(x) + 0.3 * np.cos(2 * x), 3, plus)
        d = (fh[1:] - fh[:-1]) * n / (2 * np.pi)
        errs.append(np.abs(d - (np.cos(x) - 0.6 * np.sin(2 * x))[HM:HM + n + 1]).max())
        # interface n is interface 0 of a periodic line, up to the truncation error of the explicit end rows
        assert abs(fh[-1] - fh[1]) < 2e-5
    assert errs[0] / errs[1] > 2 ** 4.5


@pytest.mark.parametrize("ntype", [1, 2, 4])
def test_compact_upwind_flux_wall_closures_converge(oracle, ntype):
    errs = []
    for n in (32, 64):
        x = np.arange(-HM, n + HM + 1) / n
        f = np.exp(x)
        fp, fm = oracle.flux_compact(f, ntype, True), oracle.flux_compact(f, ntype, False)
        lo = 1 if ntype in (1, 4) else 0
        hi = n - 1 if ntype in (2, 4) else n
        d = 0.5 * ((fp[1:] - fp[:-1]) + (fm[1:] - fm[:-1])) * n
        errs.append(np.abs(d - f[HM:HM + n + 1])[lo:hi + 1].max())
    assert errs[0] / errs[1] > 2 ** 1.8     # 3rd-order boundary fluxes -> 2nd-order flux difference


def test_mp5_keeps_smooth_data_and_clips_overshoots(oracle):
    u = np.array([1.0, 1.1, 1.2, 1.3, 1.4])
    ul = (2 * u[0] - 13 * u[1] + 47 * u[2] + 27 * u[3] - 3 * u[4]) / 60.0
    assert oracle.mp5(u, ul) == ul
    step = np.array([0.0, 0.0, 0.0, 1.0, 1.0])
    assert oracle.mp5(step, 1.7) <= 1.0 and oracle.mp5(step, -0.4) >= 0.0     # limited into [u3, u4]
    assert oracle.mp5(step, 1.7, discont=False) == 1.7                          # lshock false: no limiting


@pytest.mark.parametrize("lchardecomp", [True, False])
def test_convrsdcmp_is_conservative_on_a_periodic_block(oracle, lchardecomp):
    # qrhs(i) += Fh(i) - Fh(i-1) telescopes: over nodes 1..n of every periodic line the sum is
    # Fh(n) - Fh(0), which vanishes up to the truncation error of the explicit interface rows
    n = 24
    c = oracle.Case(n, n, n)
    c.set_upwind(543, lchardecomp, 0.3, 1e-5)
    c.gridgeom(); c.tgvini()
    rng = np.random.default_rng(5)
    for name in [f"q{m + 1}" for m in range(5)]:
        a = c.get(name)
        a[5:-5, 5:-5, 5:-5] *= 1.0 + 0.05 * rng.standard_normal((n + 1,) * 3)
        for ax in range(3):          # keep node n identical to node 0
            sl_n = [slice(5, -5)] * 3; sl_0 = [slice(5, -5)] * 3
            sl_n[ax] = 5 + n; sl_0[ax] = 5
            a[tuple(sl_n)] = a[tuple(sl_0)]
        c.set(name, a)
    c.updatefvar(); c.qswap(); c.gradcal(); c.zero_qrhs()
    if lchardecomp:
        c.ducrossensor()
        assert 0.0 < c.get("lshock")[5:-5, 5:-5, 5:-5].mean()
    assert c.convrsdcmp() == 0
    for m in range(5):
        r = c.get(f"qrhs{m + 1}")[5:-5, 5:-5, 5:-5]
        total = r[1:, 1:, 1:].sum()
        assert abs(total) < 1e-3 * np.abs(r).sum(), (m, total)
    c.close()


def test_upwind_run_is_stable_and_close_to_central(oracle):
    n = 24
    a, b = oracle.Case(n, n, n), oracle.Case(n, n, n)
    b.set_upwind(543, True, 0.3, 0.01)
    for c in (a, b):
        c.gridgeom(); c.tgvini(); c.run(3)
    ha, hb = a.history(), b.history()
    assert np.isfinite(hb).all() and np.abs(ha[:, 2] - hb[:, 2]).max() < 1e-5
    a.close(); b.close()
