"""A SECOND, independent restatement of `convrsdcmp` (the 543c upwind compact convection path: Steger-Warming split,
compact 5th-order interface flux, Roe-averaged characteristic projection, MP5 limiter) for one or more blocks, in
vectorised NumPy on top of tests/second_opinion.py.  Test infrastructure: it cross-checks oracle/upwind.hpp, which no
stored number of the reference pins ("the whole upwind/shock path").

What is restated, and how it differs from the oracle's line-by-line form:
  convrsdcmp            src/solver.F90:1271-1937   one generic direction routine applied along each axis
  flux_steger_warming   src/riemann.F90:23-161     NOT the closed form: F+- = J R diag(lambda+-) R^-1 U from the
                                                   eigen-decomposition of the flux Jacobian at the node (perfect gas:
                                                   F = A U), with the reference's smoothed lambda+- (eps = 0.04), its
                                                   full-flux switch at local Mach >= 1 and its mixed use of the
                                                   primitives and q(5) (see steger_warming)
  chardecomp            src/solver.F90:1958-2162   right eigenvectors R transcribed (3 pivot branches on the raw
                                                   metric); the LEFT eigenvectors are NOT transcribed: L = R^-1 by LU
  iwind6                src/solver.F90:1236-1257   clamped 5-node stencils, mirrored for the '-' flux
  mplimiter, MP5        src/flux.F90:363-379, :434-501 (minmod2 / minmod4 src/commfunc.F90:699-738)
  flux_compact          tests/second_opinion.py (dense solves)
`lshock` (the Ducros flags) is an input of `convrsdcmp`; `ducros_ssf` / `ducros_flags` below restate the sensor itself
(src/commcal.F90:196-357).  `crinod` is all false.
"""
import numpy as np

import second_opinion as so

HM = 5
EPS_SW = 0.04
RERO = 1e-12


def right_eigenvectors(u, H, c, ddi):
    """R(5,5,M) for velocity u(3,M), total enthalpy H, sound speed c and the (un-normalised) metric row ddi(3,M)."""
    M = H.size
    nrm = 1.0 / np.sqrt((ddi ** 2).sum(axis=0))
    g = ddi * nrm
    ug = (u * g).sum(axis=0)
    K = 0.5 * (u ** 2).sum(axis=0)
    R = np.zeros((5, 5, M))
    R[0, 0], R[1:4, 0], R[4, 0] = 1.0, u - c * g, H - ug * c           # u - c
    R[0, 1], R[1:4, 1], R[4, 1] = 1.0, u, K                             # entropy wave: H - c^2/(gamma-1) = K
    R[0, 4], R[1:4, 4], R[4, 4] = 1.0, u + c * g, H + ug * c           # u + c
    # two shear waves: the tangent pair depends on which RAW metric component serves as pivot
    piv = np.where(np.abs(ddi[0]) > RERO, 0, np.where(np.abs(ddi[1]) > RERO, 1, 2))
    if np.any((piv == 2) & ~(np.abs(ddi[2]) > RERO)):
        raise ValueError("degenerate metric normal (the reference stops: ' !! ERROR 1 @ chardecomp')")
    # (column, [(row, +-1, component of g)], energy entry as (velocity a, g b) - (velocity b, g a))
    tang = {0: ((1, 0), (2, 0)), 1: ((0, 1), (2, 1)), 2: ((0, 2), (1, 2))}
    for p in range(3):
        sel = piv == p
        if not sel.any():
            continue
        for col, (a, b) in zip((2, 3), tang[p]):
            # tangent vector t = g_b e_a - g_a e_b  (so t . g = 0); a is the non-pivot axis, b = p the pivot
            t = np.zeros((3, M))
            t[a], t[b] = g[b], -g[a]
            R[1:4, col][:, sel] = t[:, sel]
            R[4, col][sel] = (u * t).sum(axis=0)[sel]
    return R, g, ug


def _inv(R):
    return np.moveaxis(np.linalg.inv(np.moveaxis(R, 2, 0)), 0, 2)


def _apply(Mx, v):
    """(5,5,M) x (5,M) -> (5,M)"""
    return np.einsum("abm,bm->am", Mx, v)


def steger_warming(rho, vel, prs, tmp, q, ddi, jac, gamma, sos):
    """F+ and F- (5,M) at nodes from the eigen-decomposition.

    Inside an RK stage the reference calls this with primitives and q that do NOT belong to the same state: the
    interior primitives are pre-filter, q is post-filter (quirk Q2).  Its closed form takes rho, u, p and the sound
    speed sqrt(T)/M from the PRIMITIVES and only the total energy from q(5), which enters the energy flux alone, as
    lambda_1 * q5.  The eigen-form equivalent: decompose the state the primitives define (energy E_c = p/(gamma-1) +
    rho |u|^2 / 2), then add J lambda_1 (q5 - E_c) to the energy flux.  The full-flux (supersonic) branches use q."""
    K = 0.5 * (vel ** 2).sum(axis=0)
    Ec = prs / (gamma - 1.0) + rho * K
    c = sos(tmp)
    H = (Ec + prs) / rho
    R, g, ug = right_eigenvectors(vel, H, c, ddi)
    L = _inv(R)
    mag = np.sqrt((ddi ** 2).sum(axis=0))
    uu, csa = ug * mag, c * mag
    lam = np.stack([uu - csa, uu, uu, uu, uu + csa])
    lamp = 0.5 * (lam + np.sqrt(lam ** 2 + EPS_SW ** 2))
    lamm = lam - lamp
    w = _apply(L, np.stack([rho, rho * vel[0], rho * vel[1], rho * vel[2], Ec]))
    fp, fm = jac * _apply(R, lamp * w), jac * _apply(R, lamm * w)
    fp[4] += jac * lamp[1] * (q[4] - Ec)
    fm[4] += jac * lamm[1] * (q[4] - Ec)
    full = jac * np.stack([q[0] * uu, q[1] * uu + ddi[0] * prs, q[2] * uu + ddi[1] * prs, q[3] * uu + ddi[2] * prs,
                           (q[4] + prs) * uu])
    lmach = uu / csa
    fp = np.where(lmach >= 1.0, full, np.where(lmach <= -1.0, 0.0, fp))
    fm = np.where(lmach >= 1.0, 0.0, np.where(lmach <= -1.0, full, fm))
    return fp, fm


def minmod(*v):
    v = np.stack(v)
    pos, neg = np.all(v > 0, axis=0), np.all(v < 0, axis=0)
    m = np.abs(v).min(axis=0)
    return np.where(pos, m, np.where(neg, -m, 0.0))


def mp5(u, ul, shock):
    """u(5,...) upwind-ordered stencil, ul the compact (unlimited) value, shock a boolean array."""
    ump = u[2] + minmod(u[3] - u[2], 4.0 * (u[2] - u[1]))
    active = shock & ((ul - u[2]) * (ul - ump) >= 1e-10)
    dm1, d0, d1 = u[0] - 2 * u[1] + u[2], u[1] - 2 * u[2] + u[3], u[2] - 2 * u[3] + u[4]
    dhm1 = minmod(4 * dm1 - d0, 4 * d0 - dm1, dm1, d0)
    dh0 = minmod(4 * d0 - d1, 4 * d1 - d0, d0, d1)
    uul = u[2] + 4.0 * (u[2] - u[1])
    umd = 0.5 * (u[2] + u[3]) - 0.5 * dh0
    ulc = u[2] + 0.5 * (u[2] - u[1]) + 1.333333333333333 * dhm1
    umin = np.maximum(np.minimum(np.minimum(u[2], u[3]), umd), np.minimum(np.minimum(u[2], uul), ulc))
    umax = np.minimum(np.maximum(np.maximum(u[2], u[3]), umd), np.maximum(np.maximum(u[2], uul), ulc))
    return np.where(active, ul + minmod(umin - ul, umax - ul), ul)


def _pencils(a, ax, F):
    """halo'd 3-D array -> (L, M): `ax` first, the other two directions restricted to the node ranges the reference
    loops over (js:je, ks:ke ...); also returns the shape to undo it."""
    idx = [slice(F.lo[d] + HM, F.hi[d] + HM + 1) for d in range(3)]
    idx[ax] = slice(None)
    m = np.moveaxis(a[tuple(idx)], ax, 0)
    return m.reshape(m.shape[0], -1), m.shape


def direction(F, ax, gamma, mach, lshock, lchardecomp=True, bfacmpld=0.3):
    """Fh(i+1/2) differences of one direction: returns the increment of qrhs(5) on nodes 0..N (zero outside is:ie).
    mach: the Mach number of the nondimensional gas, or a callable T -> sound speed (dimensional gas)."""
    npdc = F.npdc[ax]
    rho, shp = _pencils(F.rho, ax, F)
    prs, tmp = _pencils(F.prs, ax, F)[0], _pencils(F.tmp, ax, F)[0]
    vel = np.stack([_pencils(v, ax, F)[0] for v in F.vel])               # (3, L, M)
    q = np.stack([_pencils(v, ax, F)[0] for v in F.q])                   # (5, L, M)
    ddi = np.stack([_pencils(F.dxi[ax][n], ax, F)[0] for n in range(3)])
    jac = _pencils(F.jacob, ax, F)[0]
    lsh = _pencils(lshock, ax, F)[0] > 0.5
    Lh, M = rho.shape
    dim = Lh - 1 - 2 * HM
    iss = 0 if npdc in (1, 4) else -HM                                  # solver.F90:1306-1321
    iee = dim if npdc in (2, 4) else dim + HM
    # split fluxes on iss..iee (elsewhere: never read by the stencils below, kept finite for the dense solves)
    fp, fm = np.zeros((5, Lh, M)), np.zeros((5, Lh, M))
    sl = slice(iss + HM, iee + HM + 1)
    n_nodes = iee - iss + 1
    a, b = steger_warming(rho[sl].ravel(), vel[:, sl].reshape(3, -1), prs[sl].ravel(), tmp[sl].ravel(),
                          q[:, sl].reshape(5, -1), ddi[:, sl].reshape(3, -1), jac[sl].ravel(), gamma,
                          mach if callable(mach) else (lambda T: np.sqrt(T) / mach))
    fp[:, sl], fm[:, sl] = a.reshape(5, n_nodes, M), b.reshape(5, n_nodes, M)
    fhp = np.stack([so.flux_compact(fp[m], npdc, dim, True, bfacmpld) for m in range(5)])     # (5, dim+2, M): i = -1..dim
    fhm = np.stack([so.flux_compact(fm[m], npdc, dim, False, bfacmpld) for m in range(5)])
    lo, hi = F.lo[ax], F.hi[ax]
    Fh = np.zeros((5, dim + 2, M))                                       # index i + 1
    for i in range(lo - 1, hi + 1):
        nl, nr = i + HM, i + 1 + HM
        if lchardecomp:
            wl = np.sqrt(rho[nl]) / (np.sqrt(rho[nl]) + np.sqrt(rho[nr]))
            u = wl * vel[:, nl] + (1.0 - wl) * vel[:, nr]
            H = wl * (q[4, nl] + prs[nl]) / rho[nl] + (1.0 - wl) * (q[4, nr] + prs[nr]) / rho[nr]
            c = np.sqrt((gamma - 1.0) * (H - 0.5 * (u ** 2).sum(axis=0)))
            R, _, _ = right_eigenvectors(u, H, c, 0.5 * (ddi[:, nl] + ddi[:, nr]))
            L = _inv(R)
        proj = (lambda v: _apply(L, v)) if lchardecomp else (lambda v: v)
        stp = [min(max(i + n - 3, iss), iee) + HM for n in range(1, 6)]          # iwind6 '+'
        stm = [min(max(i + 4 - n, iss), iee) + HM for n in range(1, 6)]          # iwind6 '-'
        cp = np.stack([proj(fp[:, s]) for s in stp], axis=1)                    # (5 comps, 5 nodes, M)
        cm = np.stack([proj(fm[:, s]) for s in stm], axis=1)
        hp, hm_ = proj(fhp[:, i + 1]), proj(fhm[:, i + 1])
        if i < 0:
            sh = lsh[nr]
        elif i + 1 > dim:
            sh = lsh[nl]
        else:
            sh = lsh[nl] | lsh[nr]
        skip = (npdc == 1 and i in (0, 1)) or (npdc == 2 and i in (dim - 1, dim - 2))       # mplimiter, flux.F90:370-375
        out = np.empty((5, M))
        for m in range(5):
            if skip:
                out[m] = hp[m] + hm_[m]
            else:
                out[m] = mp5(cp[m], hp[m], sh) + mp5(cm[m], hm_[m], sh)
        Fh[:, i + 1] = _apply(R, out) if lchardecomp else out
    inc = np.zeros((5, dim + 1, M))
    inc[:, lo:hi + 1] = Fh[:, lo + 1:hi + 2] - Fh[:, lo:hi + 1]
    res = []
    for m in range(5):
        full = np.moveaxis(inc[m].reshape((dim + 1,) + shp[1:]), 0, ax)
        res.append(full)
    return res


def convrsdcmp(F, gamma, mach, lshock, lchardecomp=True, bfacmpld=0.3):
    """qrhs(5) on nodes 0..N (before rhscal's sign flip)."""
    shape = tuple(s - 2 * HM for s in F.prs.shape)
    qrhs = [np.zeros(shape) for _ in range(5)]
    for ax in range(2 if F.prs.shape[2] == 1 + 2 * HM else 3):       # 2-D block: i and j only (solver.F90:1719)
        inc = direction(F, ax, gamma, mach, lshock, lchardecomp, bfacmpld)
        tgt = [slice(F.lo[a], F.hi[a] + 1) for a in range(3)]       # the other two directions: js:je | ks:ke ... only
        tgt[ax] = slice(None)
        for m in range(5):
            qrhs[m][tuple(tgt)] += inc[m]
    return qrhs


# ------------------------------------------------------------------------------------------------
# ducrossensor (src/commcal.F90:196-357)
# ------------------------------------------------------------------------------------------------
def _shift(a, ax, off, npdc, core_only=True):
    """a(node + off) along `ax` on nodes 0..N of every direction; the index is clamped at a physical boundary of an
    end block (npdc 1: below 0, npdc 2: above N), exactly as the reference does -- and NOT for npdc 4."""
    n = a.shape[ax] - 1 - 2 * HM
    idx = np.arange(n + 1) + off
    if npdc == 1:
        idx = np.maximum(idx, 0)
    if npdc == 2:
        idx = np.minimum(idx, n)
    sl = [slice(HM, -HM)] * 3
    out = np.take(a, idx + HM, axis=ax)
    sl[ax] = slice(None)
    return out[tuple(sl)]


def ducros_ssf(F, dvel):
    """ssf on nodes 0..N from the velocity gradient dvel[m][n] (nodes 0..N) and the halo'd pressure."""
    div2 = (dvel[0][0] + dvel[1][1] + dvel[2][2]) ** 2
    vort = (dvel[2][1] - dvel[1][2]) ** 2 + (dvel[0][2] - dvel[2][0]) ** 2 + (dvel[1][0] - dvel[0][1]) ** 2
    p0 = F.prs[HM:-HM, HM:-HM, HM:-HM]
    dp = []
    for ax in range(3):
        pp, pm = _shift(F.prs, ax, 1, F.npdc[ax]), _shift(F.prs, ax, -1, F.npdc[ax])
        dp.append(np.abs(pp - 2.0 * p0 + pm) / (pp + 2.0 * p0 + pm))
    return div2 / (div2 + vort + 1e-30) * np.maximum(np.maximum(dp[0], dp[1]), dp[2])


def ducros_flags(ssf_halo, npdc, shkcrt):
    """lshock (0/1) on nodes 0..N from the halo'd (exchanged) sensor: maximum over nodes -4..+5 of each direction."""
    best = np.zeros(tuple(s - 2 * HM for s in ssf_halo.shape))
    for ax in range(3):
        for off in range(-HM + 1, HM + 1):
            best = np.maximum(best, _shift(ssf_halo, ax, off, npdc[ax]))
    return (best > shkcrt).astype(float)
