"""Cross-check of the oracle's line operators (oracle/lineops.cpp) against an independent second restatement
of the same Fortran (tests/second_opinion.py: dense solves, exact-rational tables, node-indexed rows).
SURVEY.md 8c lists the ntype 1/2/4 closures of the derivative, the filter and the compact upwind flux as NOT
pinned by any stored number of the reference and names this cross-check as the mitigation."""
import numpy as np
import pytest

import second_opinion as so

HM = 5
# measured: derivative 3.6e-16, filter 8.5e-16 (2.9e-15 at alfa = 0.49), flux 5.6e-16
TOL = 1e-14


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("dim", [16, 37, 128])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_derivative_closures(oracle, ntype, dim):
    f = np.random.default_rng(100 + dim + ntype).standard_normal(dim + 1 + 2 * HM)
    assert _rel(oracle.df_compact(f, ntype), so.df_compact(f, ntype, dim)) < TOL


@pytest.mark.parametrize("alfa", [0.49, 0.45, 0.3])
@pytest.mark.parametrize("dim", [16, 37, 128])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_filter_closures(oracle, ntype, dim, alfa):
    f = np.random.default_rng(200 + dim + ntype).standard_normal(dim + 1 + 2 * HM)
    got = oracle.compact_filter(f, ntype, alfa)
    ref = so.compact_filter(f, ntype, dim, alfa)
    # alfa = 0.49 is close to the singular limit 1/2: the dense LU and the Thomas recurrence differ by cond * eps
    assert _rel(got, ref) < (5e-14 if alfa > 0.48 else TOL)


@pytest.mark.parametrize("bfac", [0.3, 1.0, 0.0])
@pytest.mark.parametrize("plus", [True, False])
@pytest.mark.parametrize("dim", [16, 37, 128])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_compact_upwind_flux_closures(oracle, ntype, dim, plus, bfac):
    f = np.random.default_rng(300 + dim + ntype).standard_normal(dim + 1 + 2 * HM)
    got = oracle.flux_compact(f, ntype, plus, bfac)
    ref = so.flux_compact(f, ntype, dim, plus, bfac)
    assert got.shape == ref.shape == (dim + 2,)
    assert _rel(got, ref) < TOL


def test_second_opinion_is_not_a_copy_of_the_oracle_tables(oracle):
    """The second restatement builds its tables from exact rationals; the oracle's factorisation tables must agree
    with them where they overlap (LHS off-diagonals), row by row."""
    for ntype in (1, 2, 3, 4):
        first, a, c, *_ = oracle.scheme_tables(False, ntype, 32)
        assert first == (0 if ntype in (1, 4) else -1)
        want = np.full(a.size, 1.0 / 3.0)
        if ntype in (1, 4):
            want[0], want[1] = 2.0, 0.25
        else:
            want[0] = 0.0
        if ntype in (2, 4):
            want[-1], want[-2] = 2.0, 0.25
        else:
            want[-1] = 0.0
        assert np.array_equal(a, want) and np.array_equal(c, want)
        first, a, c, *_ = oracle.scheme_tables(True, ntype, 32, 0.49)
        assert first == (0 if ntype in (1, 4) else -3)
        want = np.full(a.size, 0.49)
        want[0] = 0.98 if ntype in (1, 4) else 1.11
        want[-1] = 0.98 if ntype in (2, 4) else 1.11
        assert np.array_equal(a, want) and np.array_equal(c, want)


@pytest.mark.parametrize("dim", [16, 37])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_explicit_derivative_ladder(oracle, ntype, dim):
    """diff6ec against finite-difference weights DERIVED from polynomial exactness on the same stencils."""
    f = np.random.default_rng(400 + dim + ntype).standard_normal(dim + 1 + 2 * HM)
    assert _rel(oracle.diff6ec(f, ntype), so.diff6ec(f, ntype, dim)) < TOL
