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


# ---- explicit upwind family (oracle/recons.hpp) ---------------------------------------------------
def _poly_window(deg, rng):
    """point values f(j) of the derivative of a random polynomial primitive P of degree deg+1 on the 8
    nodes -3..4; a reconstruction that is exact for degree `deg` returns P(1/2)-P(-1/2) for the interface
    between nodes 0 and 1 (flux form: f(j) = P(j+1/2)-P(j-1/2))."""
    coef = rng.standard_normal(deg + 2)
    P = np.polynomial.Polynomial(coef)
    j = np.arange(-3, 5)
    f = P(j + 0.5) - P(j - 0.5)
    return f, P.deriv()(0.5)


@pytest.mark.parametrize("reschem,deg", [(0, 6), (3, 6), (5, 6)])
def test_linear_parts_of_the_explicit_schemes_are_exact_for_polynomials(oracle, reschem, deg):
    # suw7 (7th order), MP7 and MP7LD without limiting reproduce the flux-form identity for degree <= 6
    # (MP7LD with weight 1 is the 8th-order central scheme: degree 7)
    rng = np.random.default_rng(3)
    for _ in range(20):
        f, want = _poly_window(deg, rng)
        # scale down the non-linear content so that the MP switch stays off: (ul-u3)(ul-uMP) < 1e-10
        f = 1.0 + 1e-6 * f
        got = oracle.recons_exp(f, 10, 64, 3, reschem, shock=False, bfacmpld=0.3)
        assert abs(got - (1.0 + 1e-6 * want)) < 1e-13


def test_mp7ld_weight_one_is_the_eighth_order_central_scheme(oracle):
    rng = np.random.default_rng(4)
    f, want = _poly_window(7, rng)
    got = oracle.recons_exp(f, 10, 64, 3, 5, shock=False, bfacmpld=1.0)
    assert abs(got - want) < 1e-11 * max(1.0, abs(want))


@pytest.mark.parametrize("reschem", [1, 2])
def test_weno_weights_are_convex_and_order_is_high(oracle, reschem):
    errs = []
    for n in (16, 32):
        x = 2 * np.pi * np.arange(-6, n + 8) / n
        f = np.sin(x)
        fh = np.array([oracle.recons_exp(f[i + 3:i + 11], i, n, 3, reschem) for i in range(-1, n + 1)])
        d = (fh[1:] - fh[:-1]) * n / (2 * np.pi)
        errs.append(np.abs(d - np.cos(x[6:6 + n + 1])).max())
    assert errs[0] / errs[1] > 2 ** 4
    step = np.array([0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0])
    assert -1e-3 < oracle.recons_exp(step, 10, 64, 3, reschem) < 1.0 + 1e-3     # essentially non-oscillatory


def test_recons_exp_boundary_ladder(oracle):
    # flux.F90:276-296: first interface of an ntype-1 block is the plain average, the second SUW3
    f = np.arange(8, dtype=float) ** 2
    assert oracle.recons_exp(f, 0, 64, 1, 3) == 0.5 * (f[3] + f[4])
    assert abs(oracle.recons_exp(f, 1, 64, 1, 3) - (-f[2] / 6 + 5 * f[3] / 6 + f[4] / 3)) < 1e-14
    assert oracle.recons_exp(f, 63, 64, 2, 3) == 0.5 * (f[3] + f[4])
    assert oracle.recons_exp(f, 5, 64, 1, -1) == f[3]
    assert np.isnan(oracle.recons_exp(f, 5, 64, 3, 4))          # the reference stops on reschem 4


def test_convrsduwd_linear_scheme_matches_central_to_truncation(oracle):
    n = 32
    a, b = oracle.Case(n, n, n), oracle.Case(n, n, n)
    b.set_upwind_explicit(0, False, 0.3, 1.0)
    for c in (a, b):
        c.gridgeom(); c.tgvini(); c.qswap(); c.gradcal(); c.zero_qrhs(); c.rhscal()
    ra, rb = a.get("qrhs5")[5:-5, 5:-5, 5:-5], b.get("qrhs5")[5:-5, 5:-5, 5:-5]
    assert 0.0 < np.abs(ra - rb).max() < 0.05 * np.abs(ra).max()
    a.close(); b.close()
