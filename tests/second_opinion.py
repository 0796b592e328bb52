"""A SECOND, independent restatement of the reference's three compact line operators, used only by
tests/test_oracle_second_opinion.py to cross-check oracle/lineops.cpp where the reference ships no
stored numbers (SURVEY.md 8c "What is not pinned": the ntype 1/2/4 closures; mitigation "independent
second restatement in NumPy").  TEST INFRASTRUCTURE, like oracle/: nothing under astr_b200/ imports it.

Written from the Fortran alone, deliberately unlike the oracle in every choice that is free:
  * the system  A x = d  is assembled as a dense matrix and solved by numpy.linalg.solve (LU with partial
    pivoting), not by the pre-factored Thomas recurrences of src/commfunc.F90:752-813;
  * rows are addressed by NODE index through dictionaries, not by offset arithmetic;
  * right-hand sides are numpy dot products of coefficient vectors, not left-to-right scalar sums.
So agreement is to rounding (<= a few 1e-15 relative, measured), not bit for bit, and a transcription slip in
either restatement (a closure row, a table entry, a node range) shows up at O(1e-3 .. 1).

Reference: fd_scheme_initiate / compact_fd_rhs  src/derivative.F90:63-158, :210-306
           compact_filter_initiate / compact_filter / compact_filter_rhs / filter_coefficient_cal
                                                   src/filter.F90:31-100, :112-144, :156-285, :299-432
           compact_flux_initiate / flux_compact / compact_flux_rhs   src/flux.F90:32-123, :125-160, :163-266
"""
from fractions import Fraction as Fr

import numpy as np

HM = 5


class Pencil:
    """f(-hm:dim+hm) addressed by node index."""

    def __init__(self, f, dim):
        assert len(f) == dim + 1 + 2 * HM
        self.f, self.dim = np.asarray(f, dtype=np.float64), dim

    def __getitem__(self, node):
        assert -HM <= node <= self.dim + HM, node
        return self.f[node + HM]

    def window(self, nodes):
        return np.array([self[n] for n in nodes])


def _solve(rows, lower, upper, rhs):
    """rows: ordered node list; lower/upper/rhs: dicts by node.  Unit diagonal."""
    n = len(rows)
    A = np.eye(n)
    for r, node in enumerate(rows):
        if r > 0:
            A[r, r - 1] = lower[node]
        if r < n - 1:
            A[r, r + 1] = upper[node]
    x = np.linalg.solve(A, np.array([rhs[node] for node in rows]))
    return dict(zip(rows, x))


def _wall(ntype):
    """(physical boundary at the low end?, at the high end?)"""
    return ntype in (1, 4), ntype in (2, 4)


# ------------------------------------------------------------------------------------------------
# 6th-order compact first derivative, scheme '643c'
# ------------------------------------------------------------------------------------------------
def df_compact(f, ntype, dim):
    p = Pencil(f, dim)
    lo_wall, hi_wall = _wall(ntype)
    first = 0 if lo_wall else -1                      # derivative.F90:74-93
    last = dim if hi_wall else dim + 1
    rows = list(range(first, last + 1))
    third = float(Fr(1, 3))
    off = {n: third for n in rows}                    # a = c = 1/3
    d = {}
    for n in rows:                                    # interior rows first, closures overwrite
        d[n] = float(Fr(7, 9)) * (p[n + 1] - p[n - 1]) + float(Fr(1, 36)) * (p[n + 2] - p[n - 2]) \
            if -HM + 2 <= n <= dim + HM - 2 else 0.0
    explicit6 = np.array([-1 / 60, 0.15, -0.75, 0.0, 0.75, -0.15, 1 / 60])
    if lo_wall:                                       # derivative.F90:128-131, :230-248
        off[first], off[first + 1] = 2.0, 0.25
        d[first] = np.dot([-2.5, 2.0, 0.5], p.window([first, first + 1, first + 2]))
        d[first + 1] = 0.75 * (p[first + 2] - p[first])
    else:                                             # :111-113, :250-260
        off[first] = 0.0
        d[first] = np.dot(explicit6, p.window(range(first - 3, first + 4)))
    if hi_wall:                                       # :133-136, :264-281
        off[last], off[last - 1] = 2.0, 0.25
        d[last] = np.dot([2.5, -2.0, -0.5], p.window([last, last - 1, last - 2]))
        d[last - 1] = 0.75 * (p[last] - p[last - 2])
    else:
        off[last] = 0.0
        d[last] = np.dot(explicit6, p.window(range(last - 3, last + 4)))
    x = _solve(rows, off, off, d)
    return np.array([x[n] for n in range(dim + 1)])


# ------------------------------------------------------------------------------------------------
# 10th-order compact filter
# ------------------------------------------------------------------------------------------------
def _filter_tables(alfa, beter_halo, beter_bound):
    """filter.F90:299-432 as exact rationals in (alfa, beter): each entry is (c0 + c1*alfa)/den."""
    a, bh, bb = Fr(alfa), Fr(beter_halo), Fr(beter_bound)

    def lin(c0, c1, den, t=a):
        return float((Fr(c0) + Fr(c1) * t) / Fr(den))

    c6 = [lin(11, 10, 32), lin(15, 34, 64), lin(-3, 6, 32), lin(1, -2, 64)]
    c8 = [lin(93, 70, 256), lin(7, 18, 32), lin(-7, 14, 64), lin(1, -2, 32), lin(-1, 2, 256)]
    c10 = [lin(193, 126, 512), lin(105, 302, 512), lin(-15, 30, 128), lin(45, -90, 1024), lin(-5, 10, 512),
           lin(1, -2, 1024)]
    b0 = [lin(63, 1, 64, bb), lin(3, 29, 32, bb), lin(-15, 15, 64, bb), lin(5, -5, 16, bb), lin(-15, 15, 64, bb),
          lin(3, -3, 32, bb), lin(-1, 1, 64, bb)]
    b1 = [lin(1, 62, 64), lin(29, 6, 32), lin(15, 34, 64), lin(-5, 10, 16), lin(15, -30, 64), lin(-3, 6, 32),
          lin(1, -2, 64)]
    b2 = [lin(-1, 2, 64), lin(3, 26, 32), lin(49, 30, 64), lin(5, 6, 16), lin(-15, 30, 64), lin(3, -6, 32),
          lin(-1, 2, 64)]
    h0 = [lin(-1, 1, 1024, bh), lin(5, -5, 512, bh), lin(979, 45, 1024, bh), lin(15, 113, 128, bh),
          lin(-105, 105, 512, bh), lin(63, -63, 256, bh), lin(-105, 105, 512, bh), lin(15, -15, 128, bh),
          lin(-45, 45, 1024, bh), lin(5, -5, 512, bh), lin(-1, 1, 1024, bh)]
    h1 = [lin(1, -2, 1024), lin(-5, 10, 512), lin(45, 934, 1024), lin(113, 30, 128), lin(105, 302, 512),
          lin(-63, 126, 256), lin(105, -210, 512), lin(-15, 30, 128), lin(45, -90, 1024), lin(-5, 10, 512),
          lin(1, -2, 1024)]
    h2 = [lin(-1, 2, 1024), lin(5, -10, 512), lin(-45, 90, 1024), lin(15, 98, 128), lin(407, 210, 512),
          lin(63, 130, 256), lin(-105, 210, 512), lin(15, -30, 128), lin(-45, 90, 1024), lin(5, -10, 512),
          lin(-1, 2, 1024)]
    return c6, c8, c10, [b0, b1, b2], [h0, h1, h2]


def _symmetric(p, n, coef):
    """sum_k coef[k] * (f(n+k) + f(n-k)), k = 0 counted twice as the reference does (var0 = f+f)."""
    return sum(c * (p[n + k] + p[n - k]) for k, c in enumerate(coef))


def compact_filter(f, ntype, dim, alfa=0.49, beter_halo=1.11, beter_bound=0.98):
    p = Pencil(f, dim)
    c6, c8, c10, cb, ch = _filter_tables(alfa, beter_halo, beter_bound)
    lo_wall, hi_wall = _wall(ntype)
    first = 0 if lo_wall else -3                      # filter.F90:44-71
    last = dim if hi_wall else dim + 3
    rows = list(range(first, last + 1))
    off = {n: alfa for n in rows}
    off[first] = beter_bound if lo_wall else beter_halo          # :92-96
    off[last] = beter_bound if hi_wall else beter_halo
    d = {}
    lo_done = 5 if lo_wall else 3
    hi_done = 5 if hi_wall else 3
    for n in rows[lo_done:len(rows) - hi_done]:       # :271-283
        d[n] = _symmetric(p, n, c10)
    if lo_wall:                                       # :180-207: 7-node one-sided rows, then 6th and 8th order
        w = p.window(range(first, first + 7))
        for k in range(3):
            d[first + k] = np.dot(cb[k], w)
        d[first + 3] = _symmetric(p, first + 3, c6)
        d[first + 4] = _symmetric(p, first + 4, c8)
    else:                                             # :209-221: fixed 11-node window f(first-2 .. first+8)
        w = p.window(range(first - 2, first + 9))
        for k in range(3):
            d[first + k] = np.dot(ch[k], w)
    if hi_wall:                                       # :225-250 (mirror image)
        w = p.window(range(last, last - 7, -1))
        for k in range(3):
            d[last - k] = np.dot(cb[k], w)
        d[last - 3] = _symmetric(p, last - 3, c6)
        d[last - 4] = _symmetric(p, last - 4, c8)
    else:                                             # :252-263
        w = p.window(range(last + 2, last - 9, -1))
        for k in range(3):
            d[last - k] = np.dot(ch[k], w)
    x = _solve(rows, off, off, d)
    out = np.array([x[n] for n in range(dim + 1)])
    if lo_wall:                                       # :140-141: the boundary node is not filtered
        out[0] = p[0]
    if hi_wall:
        out[dim] = p[dim]
    return out


# ------------------------------------------------------------------------------------------------
# 5th-order compact upwind interface flux, scheme 543 ('+' / '-' wind)
# ------------------------------------------------------------------------------------------------
def flux_compact(f, ntype, dim, plus, bfacmpld=0.3):
    """fh(-1:dim): fh(i) is the value at interface i+1/2."""
    p = Pencil(f, dim)
    b = Fr(bfacmpld)
    lo_wall, hi_wall = _wall(ntype)
    first = -1 if lo_wall else -2                     # flux.F90:45-62
    last = dim if hi_wall else dim + 1
    rows = list(range(first, last + 1))
    big, small = float(Fr(1, 2) - b / 6), float(Fr(1, 6) + b / 6)
    lower = {n: (big if plus else small) for n in rows}        # :74-80
    upper = {n: (small if plus else big) for n in rows}
    w4 = [float(Fr(1, 18) - b / 36), float(Fr(19, 18) - 9 * b / 36), float(Fr(5, 9) + 9 * b / 36), float(b / 36)]
    d = {}
    explicit6 = np.array([1 / 60, -2 / 15, 37 / 60, 37 / 60, -2 / 15, 1 / 60])
    for n in rows[1:-1]:                              # :249-263
        nodes = [n - 1, n, n + 1, n + 2] if plus else [n + 2, n + 1, n, n - 1]
        d[n] = np.dot(w4, p.window(nodes))
    if lo_wall:                                       # :88-91, :193-204
        lower[first] = upper[first] = 2.0
        lower[first + 1] = upper[first + 1] = 0.25
        d[first] = 2.5 * p[first + 1] + 0.5 * p[first + 2]
        d[first + 1] = 0.75 * (p[first + 1] + p[first + 2])
    else:                                             # :84, :206-214
        lower[first] = upper[first] = 0.0
        d[first] = np.dot(explicit6, p.window(range(first - 2, first + 4)))
    if hi_wall:                                       # :93-96, :218-229
        lower[last] = upper[last] = 2.0
        lower[last - 1] = upper[last - 1] = 0.25
        d[last] = 2.5 * p[last] + 0.5 * p[last - 1]
        d[last - 1] = 0.75 * (p[last] + p[last - 1])
    else:                                             # :85, :231-238
        lower[last] = upper[last] = 0.0
        d[last] = np.dot(explicit6, p.window(range(last - 2, last + 4)))
    x = _solve(rows, lower, upper, d)
    return np.array([x[n] for n in range(-1, dim + 1)])


# ------------------------------------------------------------------------------------------------
# explicit 6th-order central derivative with the 2-2-4 wall ladder ('642e', src/derivative.F90:338-413)
# ------------------------------------------------------------------------------------------------
def _fd_weights(offsets):
    """Exact first-derivative weights on the integer stencil `offsets` (unit spacing): the unique weights that
    differentiate every polynomial of degree < len(offsets) exactly (Vandermonde system over the rationals).
    The reference hard-codes them (0.75, -0.15, 1/60; 2/3, -1/12; ...); here they are DERIVED."""
    n = len(offsets)
    # sum_j w_j * offsets_j**p = d/dx x**p at 0 = (1 if p == 1 else 0)
    A = [[Fr(o) ** p for o in offsets] + [Fr(1 if p == 1 else 0)] for p in range(n)]
    for col in range(n):                              # Gauss-Jordan over Fraction
        piv = next(r for r in range(col, n) if A[r][col] != 0)
        A[col], A[piv] = A[piv], A[col]
        A[col] = [v / A[col][col] for v in A[col]]
        for r in range(n):
            if r != col and A[r][col] != 0:
                A[r] = [vr - A[r][col] * vc for vr, vc in zip(A[r], A[col])]
    return [float(A[r][n]) for r in range(n)]


def diff6ec(f, ntype, dim):
    p = Pencil(f, dim)
    lo_wall, hi_wall = _wall(ntype)
    out = np.empty((dim + 1,) + p.f.shape[1:])
    for i in range(dim + 1):
        lo, hi = (i if lo_wall else HM), (dim - i if hi_wall else HM)      # nodes available towards each wall
        if lo == 0:
            offs = [0, 1, 2]                          # one-sided 2nd order at the wall node
        elif hi == 0:
            offs = [0, -1, -2]
        else:
            half = min(lo, hi, 3)                     # central: 2nd order next to the wall, 4th, then 6th
            offs = list(range(-half, half + 1))
        out[i] = np.dot(_fd_weights(offs), p.window([i + o for o in offs]))
    return out
