"""A SECOND, independent restatement of `gradcal` and `rhscal` (central compact path: `convrsdcal6`,
`qrhs=-qrhs`, `diffrsdcal6` with `miucal`) for ONE block in main-solver mode, in vectorised NumPy on top of
tests/second_opinion.py's dense-solve derivative.  Test infrastructure (see second_opinion.py): it cross-checks
oracle/solver.cpp where the reference ships no stored numbers (curvilinear metrics, main-solver
assembly, wall closures in the RHS).

Written from the Fortran alone:
  gradcal       src/comsolver.F90:244-497     dvel(m,n) = sum_d d(u_m)/d(xi_d) * dxi(d,n),  dtmp likewise
  rhscal        src/solver.F90:185-282        convection, qrhs=-qrhs, diffusion
  convrsdcal6   src/solver.F90:2173-2341      J*(rho U, rho u U + xi_x p, ..., (E+p) U), compact derivative
  diffrsdcal6   src/solver.F90:2354-2873      sigma, qflux pointwise on 0:N; dataswap; metric contraction * J; derivative
  miucal        src/fludyna.F90:791-812       nondimensional Sutherland law
  dataswap      src/parallel.F90:4169-4174    single-block periodic wrap: halo -k <- node N-k, halo N+k <- node k

Arrays are the oracle's own (halo'd, Fortran order, index = node + 5); what is compared is everything computed
from them.  Unlike the oracle this code works on whole arrays (einsum / broadcasting), solves each direction's
tridiagonal systems for all pencils at once with a dense LU, and keeps no per-pencil state.
"""
import numpy as np

import second_opinion as so

HM = 5


def deriv(F, axis, ntype):
    """Compact 6th-order d/dxi along `axis` of a halo'd array: nodes 0..N along that axis, every index elsewhere."""
    Fm = np.moveaxis(F, axis, 0)
    shp = Fm.shape
    dim = shp[0] - 1 - 2 * HM
    out = so.df_compact(Fm.reshape(shp[0], -1), ntype, dim).reshape((dim + 1,) + shp[1:])
    return np.moveaxis(out, 0, axis)


def core(a):
    return a[HM:-HM, HM:-HM, HM:-HM]


def _core_other_axes(a, axis):
    """Restrict the two axes other than `axis` to nodes 0..N (the derivative already did so along `axis`)."""
    sl = [slice(HM, -HM)] * 3
    sl[axis] = slice(None)
    return a[tuple(sl)]


def wrap_halos(a, homo):
    """dataswap of a single block: periodic directions wrap (shared end node), the others keep what they have."""
    for ax in range(3):
        if not homo[ax]:
            continue
        n = a.shape[ax] - 1 - 2 * HM
        idx = [slice(HM, -HM)] * 3            # the reference copies 0:jm, 0:km of the other two directions only
        lo, src_lo, hi, src_hi = list(idx), list(idx), list(idx), list(idx)
        lo[ax], src_lo[ax] = slice(0, HM), slice(n, n + HM)                       # -5..-1  <-  N-5..N-1
        hi[ax], src_hi[ax] = slice(n + HM + 1, n + 2 * HM + 1), slice(HM + 1, 2 * HM + 1)   # N+1..N+5 <- 1..5
        a[tuple(lo)] = a[tuple(src_lo)]
        a[tuple(hi)] = a[tuple(src_hi)]
    return a


class Fields:
    """The inputs, pulled from an oracle case by name (halo'd arrays)."""

    def __init__(self, c, ib=0):
        g = lambda nm: c.get(nm, ib)
        self.q = [g(f"q{m + 1}") for m in range(5)]
        self.vel = [g("u"), g("v"), g("w")]
        self.prs, self.tmp, self.rho = g("prs"), g("tmp"), g("rho")
        self.jacob = g("jacob")
        self.dxi = [[g(f"dxi{d + 1}{n + 1}") for n in range(3)] for d in range(3)]     # dxi[d][n] = d xi_d / d x_n
        info = c.block_info(ib)
        self.npdc = info["npdc"]
        r = info["is_ie"]
        self.lo, self.hi = (r[0], r[2], r[4]), (r[1], r[3], r[5])                 # is, js, ks / ie, je, ke


def ntype_of(npdc):
    """npdc (1: boundary at 0, 2: boundary at N, 3: interfaces / periodic, 4: both boundaries) is the closure type."""
    return npdc


def gradcal(F):
    """dvel[m][n], dtmp[n] on nodes 0..N."""
    scal = F.vel + [F.tmp]
    raw = [[core_d(deriv(f, d, ntype_of(F.npdc[d])), d) for d in range(3)] for f in scal]      # raw[f][d]
    grad = [[sum(raw[f][d] * core(F.dxi[d][n]) for d in range(3)) for n in range(3)] for f in range(4)]
    return grad[:3], grad[3]


def core_d(a, axis):
    return _core_other_axes(a, axis)


def miucal(T, tempconst):
    return T * np.sqrt(T) * (1.0 + tempconst) / (T + tempconst)


def stress_and_heat_flux(F, dvel, dtmp, th):
    """sigma(6) [11 12 13 22 23 33] and qflux(3) on nodes 0..N (diffrsdcal6's pointwise block, nondimensional)."""
    T, v = core(F.tmp), [core(u) for u in F.vel]
    miu = miucal(T, th["tempconst"]) / th["reynolds"]
    S = [[0.5 * (dvel[a][b] + dvel[b][a]) for b in range(3)] for a in range(3)]
    skk = (S[0][0] + S[1][1] + S[2][2]) / 3.0
    two_mu = 2.0 * miu
    tau = [[two_mu * (S[a][b] - (skk if a == b else 0.0)) for b in range(3)] for a in range(3)]
    hcc = (miu / th["prandtl"]) / th["const5"]
    qf = [hcc * dtmp[n] + sum(tau[n][b] * v[b] for b in range(3)) for n in range(3)]
    return tau, qf


def rhscal(F, th, homo, diffterm=True):
    """qrhs(5) on nodes 0..N, zero outside the ranges the reference updates."""
    shape = core(F.prs).shape
    qrhs = [np.zeros(shape) for _ in range(5)]
    lo, hi = F.lo, F.hi

    def box(ranges):
        return tuple(slice(a, b + 1) for a, b in ranges)

    # ---- convection: loops over js:je, ks:ke (resp.) and adds on is:ie -> the box is:ie x js:je x ks:ke --------------
    conv_box = box(zip(lo, hi))
    for d in range(3):
        U = sum(F.dxi[d][n] * F.vel[n] for n in range(3))
        flux = [F.jacob * F.q[0] * U] + \
               [F.jacob * (F.q[1 + n] * U + F.dxi[d][n] * F.prs) for n in range(3)] + \
               [F.jacob * (F.q[4] + F.prs) * U]
        for m in range(5):
            qrhs[m][conv_box] += core_d(deriv(flux[m], d, ntype_of(F.npdc[d])), d)[conv_box]
    qrhs = [-r for r in qrhs]
    if not diffterm:
        return qrhs
    # ---- diffusion -----------------------------------------------------------------------------------------------
    dvel, dtmp = gradcal(F)
    tau, qf = stress_and_heat_flux(F, dvel, dtmp, th)

    def halo_field(a):
        full = np.zeros(F.prs.shape)
        core(full)[...] = a
        return wrap_halos(full, homo)

    tauh = [[halo_field(tau[a][b]) for b in range(3)] for a in range(3)]
    qfh = [halo_field(q) for q in qf]
    n_nodes = [s - 1 for s in shape]
    for d in range(3):
        # loops over 0:N of the two other directions, adds on (is:ie | js:je | ks:ke) of direction d only
        ranges = [(0, n_nodes[a]) for a in range(3)]
        ranges[d] = (lo[d], hi[d])
        dbox = box(ranges)
        cols = [sum(tauh[m][n] * F.dxi[d][n] for n in range(3)) * F.jacob for m in range(3)] + \
               [sum(qfh[n] * F.dxi[d][n] for n in range(3)) * F.jacob]
        for m in range(4):
            qrhs[1 + m][dbox] += core_d(deriv(cols[m], d, ntype_of(F.npdc[d])), d)[dbox]
    return qrhs
