"""A SECOND restatement of the explicit reconstruction family `recons_exp` (src/flux.F90:269-1055) and of
`convrsduwd` (src/solver.F90:548-1201), the explicit-upwind convection path (`conschm='..e'`),
in vectorised NumPy.  Test infrastructure: cross-checks oracle/recons.hpp and oracle/upwind.hpp's convrsduwd.

Every constant the reference hard-codes as a 15-digit decimal is DERIVED here over the rationals:
  * the linear upwind reconstructions suw3 / suw5 / suw7 and the candidate reconstructions of WENO5 / WENO7 are the
    unique weights that reproduce the interface value of the polynomial with the given CELL AVERAGES (`recon_weights`);
  * the ideal weights C_k of the WENO schemes follow from matching the (2r-1)-point linear scheme;
  * the smoothness indicators are  beta_k = d1^2 + 13/12 d2^2 (+ 781/720 d3^2 for r = 4)  with d_l = h^l p_k^(l)(x_i),
    the derivatives at the cell centre of the polynomial INTERPOLATING the point values of stencil k (`deriv_weights`)
    -- that is what the reference's 12 + 6 hand-written difference stencils are;
  * the low-dissipation linear parts of MP5LD / MP7LD are (1 - w) * upwind + w * central of the next even order
    (the reference writes b_j = c_j + w/60 * binomial, resp. w/280).
What is transcribed because it is the reference's own choice: eps = 1e-6 and the exponent in the WENO weights, the
Z-weights  C (1 + tau / (beta + eps)^2)  with tau = abs(beta_last - beta_first), the MP limiter and its switches
(MP7 and MP5LD: `var1 >= 1e-10`; MP7LD: the shock flag; MP5 here: both), ROUND, and the near-boundary ladder.
`convrsduwd` reuses the eigenvectors, the Steger-Warming split and the MP limiter of tests/second_opinion_upwind.py.
"""
from fractions import Fraction as Fr
from functools import lru_cache

import numpy as np

import second_opinion_upwind as U

HM = 5
EPS = 1e-6


def _solve_fr(A, b):
    n = len(A)
    M = [list(map(Fr, row)) + [Fr(rhs)] for row, rhs in zip(A, b)]
    for col in range(n):
        piv = next(r for r in range(col, n) if M[r][col] != 0)
        M[col], M[piv] = M[piv], M[col]
        M[col] = [v / M[col][col] for v in M[col]]
        for r in range(n):
            if r != col and M[r][col] != 0:
                M[r] = [vr - M[r][col] * vc for vr, vc in zip(M[r], M[col])]
    return [M[r][n] for r in range(n)]


def recon_weights(offsets):
    return list(_recon_weights(tuple(offsets)))


@lru_cache(maxsize=None)
def _recon_weights(offsets):
    """w_j with  sum_j w_j ubar_j = p(1/2),  p the polynomial of degree len-1 whose averages over the unit cells
    centred at `offsets` are ubar_j.  Exact rationals."""
    n = len(offsets)
    # unknowns: monomial coefficients a_m; cell average of x^m over [j-1/2, j+1/2]
    avg = [[(Fr(2 * j + 1, 2) ** (m + 1) - Fr(2 * j - 1, 2) ** (m + 1)) / (m + 1) for m in range(n)] for j in offsets]
    # p(1/2) = sum_m a_m (1/2)^m = sum_j w_j ubar_j with a = avg^-1 ubar  ->  w = avg^-T e
    e = [Fr(1, 2) ** m for m in range(n)]
    return tuple(_solve_fr([[avg[j][m] for j in range(n)] for m in range(n)], e))


def deriv_weights(offsets, order):
    return list(_deriv_weights(tuple(offsets), order))


@lru_cache(maxsize=None)
def _deriv_weights(offsets, order):
    """w_j with  sum_j w_j u_j = h^order p^(order)(0),  p interpolating the point values u_j at `offsets`."""
    n = len(offsets)
    fact = 1
    for k in range(2, order + 1):
        fact *= k
    rhs = [Fr(fact if m == order else 0) for m in range(n)]
    return tuple(_solve_fr([[Fr(o) ** m for o in offsets] for m in range(n)], rhs))


def _lin(weights, u):
    return sum(float(w) * x for w, x in zip(weights, u))


def suw(u):
    """Linear upwind-biased reconstruction on an odd stencil centred on the upwind node."""
    n = len(u) // 2
    return _lin(recon_weights(range(-n, n + 1)), u)


def weno(u, z):
    """WENO of order 2r-1 on u(2r-1, ...): r = 3 (5 points) or r = 4 (7 points)."""
    r = (len(u) + 1) // 2
    c = r - 1
    big = recon_weights(range(-c, c + 1))
    cand, beta = [], []
    for k in range(r):
        offs = list(range(-c + k, k + 1))
        w = recon_weights(offs)
        cand.append((offs, w))
        d = [_lin(deriv_weights(offs, l), [u[c + o] for o in offs]) for l in range(1, r)]
        b = d[0] ** 2 + 13.0 / 12.0 * d[1] ** 2
        if r == 4:
            b = b + 781.0 / 720.0 * d[2] ** 2
        beta.append(b)
    C = _ideal_weights(r)
    if z:
        tau = np.abs(beta[-1] - beta[0])
        alpha = [float(C[k]) + float(C[k]) * (tau / (beta[k] + EPS) ** 2) for k in range(r)]
    else:
        alpha = [float(C[k]) / (beta[k] + EPS) ** 2 for k in range(r)]
    tot = sum(alpha)
    return sum(alpha[k] / tot * _lin(cand[k][1], [u[c + o] for o in cand[k][0]]) for k in range(r))


@lru_cache(maxsize=None)
def _ideal_weights(r):
    """C_k with  sum_k C_k * (candidate k) = the (2r-1)-point linear scheme: the outermost point fixes C_0, the next
    one C_1, ...; the remaining r-1 equations are then satisfied identically (checked)."""
    c = r - 1
    big = recon_weights(range(-c, c + 1))
    cand = [(list(range(-c + k, k + 1)), recon_weights(range(-c + k, k + 1))) for k in range(r)]
    A = [[(cand[k][1][cand[k][0].index(j)] if j in cand[k][0] else Fr(0)) for k in range(r)] for j in range(-c, c + 1)]
    C = _solve_fr([A[i] for i in range(r)], [big[i] for i in range(r)])
    assert all(sum(A[i][k] * C[k] for k in range(r)) == big[i] for i in range(2 * r - 1))
    return tuple(C)


def mp_limit(u5, ul, active):
    """The MP limiter around the upwind node u5[2] (src/flux.F90:459-492); `active`: where it may act."""
    return U.mp5(np.stack(u5), ul, active)


def mpld_linear(u, w):
    """(1 - w) * upwind(2n-1 points) + w * central(2n points) on an even stencil u(2n, ...)."""
    n = len(u) // 2
    up = recon_weights(range(-n + 1, n))
    ce = recon_weights(range(-n + 1, n + 1))
    return sum((float(up[j]) * (1.0 - w) if j < 2 * n - 1 else 0.0) * u[j] + float(ce[j]) * w * u[j] for j in range(2 * n))


def round_scheme(u):
    eps = 1e-16
    z = (u[1] - u[0] + eps) / (u[2] - u[0] + eps)
    a1 = 1.0 + 12.0 * z * z
    a2 = 1.0 + 5.0 * (z - 1.0) ** 2
    pl = 1100.0 * (z - 0.05) ** 3 * (0.47 - z) ** 3
    pr = 18000.0 * (z - 0.55) ** 3 * (0.97 - z) ** 5
    p1 = 0.833333333333333 * z + 0.333333333333333 + np.maximum(pl, 0.0) + np.maximum(pr, 0.0)
    p2, p3 = 1.5 * z, 0.5 * z + 0.5
    w1, w2 = 1.0 / a1 ** 4, 1.0 / a2 ** 8
    g = (p1 * (1.0 - w1) + p2 * w1) * (1.0 - w2) + p3 * w2
    return g * (u[2] - u[0]) + u[0]


def recons_exp(f, inode, dim, ntype, reschem, shock, bfacmpld):
    """f(8, ...): the upwind-ordered 8-node stencil (f[3] is the upwind node of the interface)."""
    true = np.ones(np.shape(f[0]), dtype=bool)
    if reschem == -1:
        return f[3] + 0.0
    if (ntype == 1 and inode == 0) or (ntype == 2 and inode == dim - 1):
        return 0.5 * (f[3] + f[4])
    if (ntype == 1 and inode == 1) or (ntype == 2 and inode == dim - 2):
        return suw(f[2:5])
    near = (ntype == 1 and inode == 2) or (ntype == 2 and inode == dim - 3)
    if reschem == 6:
        return round_scheme(f[2:5])
    if near:
        s5 = f[1:6]
        if reschem == 0:
            return suw(s5)
        if reschem in (1, 2):
            return weno(s5, z=reschem == 2)
        if reschem == 3:
            return mp_limit(s5, suw(s5), true)
        if reschem == 5:
            return mp_limit(s5, mpld_linear(f[1:7], bfacmpld), true)
    else:
        s7 = f[0:7]
        if reschem == 0:
            return suw(s7)
        if reschem in (1, 2):
            return weno(s7, z=reschem == 2)
        if reschem == 3:
            return mp_limit(f[1:6], suw(s7), true)
        if reschem == 5:
            # MP7LD: the limiter acts wherever the shock flag is set -- no (ul - u)(ul - uMP) test (as written)
            ul = mpld_linear(f[0:8], bfacmpld)
            lim = _mp_unconditional(f[1:6], ul)
            return np.where(shock, lim, ul)
    raise NotImplementedError(reschem)


def _mp_unconditional(u, ul):
    u = np.stack(u)
    dm1, d0, d1 = u[0] - 2 * u[1] + u[2], u[1] - 2 * u[2] + u[3], u[2] - 2 * u[3] + u[4]
    dhm1 = U.minmod(4 * dm1 - d0, 4 * d0 - dm1, dm1, d0)
    dh0 = U.minmod(4 * d0 - d1, 4 * d1 - d0, d0, d1)
    uul = u[2] + 4.0 * (u[2] - u[1])
    umd = 0.5 * (u[2] + u[3]) - 0.5 * dh0
    ulc = u[2] + 0.5 * (u[2] - u[1]) + 1.333333333333333 * dhm1
    umin = np.maximum(np.minimum(np.minimum(u[2], u[3]), umd), np.minimum(np.minimum(u[2], uul), ulc))
    umax = np.minimum(np.maximum(np.maximum(u[2], u[3]), umd), np.maximum(np.maximum(u[2], uul), ulc))
    return ul + U.minmod(umin - ul, umax - ul)


def direction(F, ax, gamma, mach, lshock, reschem, lchardecomp, bfacmpld):
    """convrsduwd along one direction: increments of qrhs(5) on nodes 0..N (zero outside is:ie)."""
    npdc = F.npdc[ax]
    rho, shp = U._pencils(F.rho, ax, F)
    prs, tmp = U._pencils(F.prs, ax, F)[0], U._pencils(F.tmp, ax, F)[0]
    vel = np.stack([U._pencils(v, ax, F)[0] for v in F.vel])
    q = np.stack([U._pencils(v, ax, F)[0] for v in F.q])
    ddi = np.stack([U._pencils(F.dxi[ax][n], ax, F)[0] for n in range(3)])
    jac = U._pencils(F.jacob, ax, F)[0]
    lsh_all = U._pencils(lshock, ax, F)[0] > 0.5
    Lh, M = rho.shape
    dim = Lh - 1 - 2 * HM
    iss = 0 if npdc in (1, 4) else -HM
    iee = dim if npdc in (2, 4) else dim + HM
    fp, fm = np.zeros((5, Lh, M)), np.zeros((5, Lh, M))
    sl = slice(iss + HM, iee + HM + 1)
    nn = iee - iss + 1
    a, b = U.steger_warming(rho[sl].ravel(), vel[:, sl].reshape(3, -1), prs[sl].ravel(), tmp[sl].ravel(),
                            q[:, sl].reshape(5, -1), ddi[:, sl].reshape(3, -1), jac[sl].ravel(), gamma,
                            mach if callable(mach) else (lambda T: np.sqrt(T) / mach))
    fp[:, sl], fm[:, sl] = a.reshape(5, nn, M), b.reshape(5, nn, M)
    lo, hi = F.lo[ax], F.hi[ax]
    Fh = np.zeros((5, dim + 2, M))
    for i in range(lo - 1, hi + 1):
        nl, nr = i + HM, i + 1 + HM
        if i < 0:
            sh = lsh_all[nr]
        elif i + 1 > dim:
            sh = lsh_all[nl]
        else:
            sh = lsh_all[nl] | lsh_all[nr]
        stp = [min(max(i + n - 4, iss), iee) + HM for n in range(1, 9)]           # iwind8 '+'
        stm = [min(max(i + 5 - n, iss), iee) + HM for n in range(1, 9)]           # iwind8 '-'

        def total(proj):
            cp = np.stack([proj(fp[:, s]) for s in stp], axis=1)                  # (5 comps, 8 nodes, M)
            cm = np.stack([proj(fm[:, s]) for s in stm], axis=1)
            return np.stack([recons_exp(cp[m], i, dim, npdc, reschem, sh, bfacmpld) +
                             recons_exp(cm[m], i, dim, npdc, reschem, sh, bfacmpld) for m in range(5)])
        plain = total(lambda v: v)
        if lchardecomp:
            wl = np.sqrt(rho[nl]) / (np.sqrt(rho[nl]) + np.sqrt(rho[nr]))
            u = wl * vel[:, nl] + (1.0 - wl) * vel[:, nr]
            H = wl * (q[4, nl] + prs[nl]) / rho[nl] + (1.0 - wl) * (q[4, nr] + prs[nr]) / rho[nr]
            c = np.sqrt((gamma - 1.0) * (H - 0.5 * (u ** 2).sum(axis=0)))
            R, _, _ = U.right_eigenvectors(u, H, c, 0.5 * (ddi[:, nl] + ddi[:, nr]))
            L = U._inv(R)
            char = U._apply(R, total(lambda v: U._apply(L, v)))
            Fh[:, i + 1] = np.where(sh, char, plain)            # characteristic only on flagged interfaces
        else:
            Fh[:, i + 1] = plain
    inc = np.zeros((5, dim + 1, M))
    inc[:, lo:hi + 1] = Fh[:, lo + 1:hi + 2] - Fh[:, lo:hi + 1]
    return [np.moveaxis(inc[m].reshape((dim + 1,) + shp[1:]), 0, ax) for m in range(5)]


def convrsduwd(F, gamma, mach, lshock, reschem, lchardecomp, bfacmpld=0.3):
    shape = tuple(s - 2 * HM for s in F.prs.shape)
    qrhs = [np.zeros(shape) for _ in range(5)]
    for ax in range(2 if F.prs.shape[2] == 1 + 2 * HM else 3):
        inc = direction(F, ax, gamma, mach, lshock, reschem, lchardecomp, bfacmpld)
        tgt = [slice(F.lo[a], F.hi[a] + 1) for a in range(3)]
        tgt[ax] = slice(None)
        for m in range(5):
            qrhs[m][tuple(tgt)] += inc[m]
    return qrhs
