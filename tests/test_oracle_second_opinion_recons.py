"""oracle/recons.hpp (`recons_exp` and its schemes) and oracle/upwind.hpp's `convrsduwd` against the second
restatement tests/second_opinion_recons.py, whose linear weights, WENO candidate / ideal weights and smoothness
stencils are DERIVED over the rationals instead of copied."""
from fractions import Fraction as Fr

import numpy as np
import pytest

import second_opinion_recons as X
import second_opinion_rhs as R
from gpu_common import auto_shkcrt, clean_metrics, skewed_x

SCHEMES = (-1, 0, 1, 2, 3, 5, 6)       # f(4), linear upwind, WENO-JS, WENO-Z, MP, MP-LD, ROUND


def test_derived_constants_are_the_ones_the_reference_hard_codes():
    assert X.recon_weights(range(-1, 2)) == [Fr(-1, 6), Fr(5, 6), Fr(1, 3)]                       # suw3, flux.F90:394
    assert X.recon_weights(range(-2, 3)) == [Fr(v, 60) for v in (2, -13, 47, 27, -3)]             # suw5, :406
    assert X.recon_weights(range(-3, 4)) == [Fr(v, 420) for v in (-3, 25, -101, 319, 214, -38, 4)]   # suw7, :418
    assert X.recon_weights(range(-2, 1)) == [Fr(1, 3), Fr(-7, 6), Fr(11, 6)]                      # WENO5 uh1, :771
    assert X.recon_weights(range(-3, 1)) == [Fr(-1, 4), Fr(13, 12), Fr(-23, 12), Fr(25, 12)]      # WENO7 uh1, :835
    assert X.deriv_weights(range(-3, 1), 1) == [Fr(v, 6) for v in (-2, 9, -18, 11)]               # WENO7 df1, :844
    assert X.deriv_weights(range(-3, 1), 2) == [-1, 4, -5, 2] and X.deriv_weights(range(-3, 1), 3) == [-1, 3, -3, 1]
    assert X.recon_weights(range(-2, 4)) == [Fr(v, 60) for v in (1, -8, 37, 37, -8, 1)]           # MP5LD at weight 1


@pytest.mark.parametrize("reschem", SCHEMES)
def test_recons_exp_ladder(oracle, reschem):
    """Interior interfaces and every rung of the near-boundary ladder (ntype 1: inode 0..3, ntype 2: dim-1..dim-4),
    on random and on near-monotone stencils (limiter inactive / active), both values of the shock flag."""
    rng = np.random.default_rng(3)
    worst = 0.0
    for ntype, dim, inode in ((3, 20, 7), (1, 20, 0), (1, 20, 1), (1, 20, 2), (1, 20, 3), (2, 20, 19), (2, 20, 18),
                              (2, 20, 17), (2, 20, 16), (4, 20, 0), (4, 20, 2)):
        for shock in (True, False):
            for t in range(20):
                f = rng.standard_normal(8) if t % 2 else np.sort(rng.standard_normal(8)) + 0.05 * rng.standard_normal(8)
                want = oracle.recons_exp(f, inode, dim, ntype, reschem, shock, 0.3)
                got = float(X.recons_exp(f.reshape(8, 1), inode, dim, ntype, reschem, np.array([shock]), 0.3)[0])
                worst = max(worst, abs(got - want) / max(abs(want), 1.0))
    # measured <= 7e-15 (WENO: the reference's constants are 14-15 digit decimals)
    assert worst < 1e-13


def _run(oracle, n, homo, reschem, lchardecomp, blocks=(1, 1, 1), mach=0.1):
    c = oracle.Case(*n, homo=homo, blocks=blocks, mach=mach)
    x = skewed_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); clean_metrics(c); c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(c.nblocks):
        for m in range(5):
            a = c.get(f"q{m + 1}", ib)
            a *= 1.0 + 2e-2 * rng.standard_normal(a.shape)
            c.set(f"q{m + 1}", a, ib)
    c.updatefvar()
    sensor = lchardecomp or reschem == 5            # solver.F90:221: the Ducros sensor runs for these
    shk = auto_shkcrt(c, 0.3) if sensor else 0.01
    c.set_upwind_explicit(reschem, lchardecomp, 0.3, shk)
    c.qswap(); c.gradcal()
    if sensor:
        c.ducrossensor()
    c.zero_qrhs(); c.convrsduwd()
    worst = 0.0
    for ib in range(c.nblocks):
        F = R.Fields(c, ib)
        lsh = c.get("lshock", ib) if sensor else np.ones(F.prs.shape)
        got = X.convrsduwd(F, 1.4, mach, lsh, reschem, lchardecomp, 0.3)
        for m in range(5):
            want = R.core(c.get(f"qrhs{m + 1}", ib))
            worst = max(worst, np.abs(got[m] - want).max() / np.abs(want).max())
    c.close()
    return worst


@pytest.mark.parametrize("lchardecomp", [True, False])
@pytest.mark.parametrize("reschem", SCHEMES)
def test_convrsduwd_periodic(oracle, reschem, lchardecomp):
    # measured 2e-14 .. 1e-13; WENO-Z 5e-13 (tau / (beta + eps)^2 amplifies the last digits of the reference's decimals)
    assert _run(oracle, (16, 14, 12), (True, True, True), reschem, lchardecomp) < (5e-12 if reschem == 2 else 1e-12)


@pytest.mark.parametrize("reschem,lchardecomp", [(3, True), (5, True), (1, False), (0, True)])
def test_convrsduwd_with_walls_and_block_interfaces(oracle, reschem, lchardecomp):
    """ntype 1 | 2 blocks: the whole near-boundary ladder of recons_exp runs inside the flux routine."""
    assert _run(oracle, (28, 12, 12), (False, True, True), reschem, lchardecomp, blocks=(2, 1, 1)) < 1e-12


def test_ideal_weights():
    assert X._ideal_weights(3) == (Fr(1, 10), Fr(6, 10), Fr(3, 10))                               # flux.F90:787-789
    assert X._ideal_weights(4) == (Fr(1, 35), Fr(12, 35), Fr(18, 35), Fr(4, 35))                  # :877-880
