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
  dataswap      src/parallel.F90:4132-4370    halo -k <- low neighbour's node N-k, halo N+k <- high neighbour's node k
                                              (a periodic direction owned by one block wraps onto itself, :4169-4174)

Arrays are the oracle's own (halo'd, Fortran order, index = node + 5); what is compared is everything computed
from them.  Unlike the oracle this code works on whole arrays (einsum / broadcasting), solves each direction's
tridiagonal systems for all pencils at once with a dense LU, and keeps no per-pencil state.
"""
import numpy as np

import second_opinion as so

HM = 5


def deriv(F, axis, ntype, explicit=False):
    """d/dxi along `axis` of a halo'd array (`fds%central`): compact 6th order ('643c') or, with explicit=True, the
    explicit 6th-order ladder ('642e'); nodes 0..N along that axis, every index elsewhere."""
    Fm = np.moveaxis(F, axis, 0)
    shp = Fm.shape
    dim = shp[0] - 1 - 2 * HM
    op = so.diff6ec if explicit else so.df_compact
    out = op(Fm.reshape(shp[0], -1), ntype, dim).reshape((dim + 1,) + shp[1:])
    return np.moveaxis(out, 0, axis)


class Gas:
    """The gas law and its constants in both modes of the reference (src/fludyna.F90:136-179 thermal, :832-859 sos,
    :312-376 fvar2q's cotem): nondimensional (const1, const2, Mach) or SI (rgas = 287.1, cv)."""

    def __init__(self, th):
        self.dim = bool(th.get("dimensional", False))
        self.gamma = th.get("gamma", 1.4)
        if self.dim:
            self.rgas = th.get("rgas", 287.1)
            self.cotem = self.rgas / (self.gamma - 1.0)
        else:
            self.mach = th.get("mach")          # only what the caller's dictionary allows; the rest stays None
            m2 = None if self.mach is None else self.mach ** 2
            self.const2 = th.get("const2", None if m2 is None else self.gamma * m2)
            self.cotem = th.get("const1", None if m2 is None else 1.0 / (self.gamma * (self.gamma - 1.0) * m2))
        self.const6 = 1.0 / (self.gamma - 1.0)

    def T_of(self, p, rho):
        return p / rho / self.rgas if self.dim else p / rho * self.const2

    def rho_of(self, p, T):
        return p / T / self.rgas if self.dim else p / T * self.const2

    def sos(self, T):
        return np.sqrt(self.gamma * self.rgas * T) if self.dim else np.sqrt(T) / self.mach


def core(a):
    return a[HM:-HM, HM:-HM, HM:-HM]


def ndims_of(a):
    """2 for a 2-D block (ka = 0: one node plane and its ten halo planes), else 3."""
    return 2 if a.shape[2] == 1 + 2 * HM else 3


def _core_other_axes(a, axis):
    """Restrict the two axes other than `axis` to nodes 0..N (the derivative already did so along `axis`)."""
    sl = [slice(HM, -HM)] * 3
    sl[axis] = slice(None)
    return a[tuple(sl)]


def exchange_halos(arrs, blocks, homo):
    """dataswap of one field over all blocks (src/parallel.F90:4132-4370): halo planes -5..-1 take the low
    neighbour's nodes N-5..N-1, halo planes N+1..N+5 the high neighbour's nodes 1..5 (blocks share their end
    nodes); a periodic direction owned by one block wraps onto itself (:4169-4174); faces without a neighbour keep
    what they have.  Only nodes 0..N of the two other directions travel."""
    out = [a.copy() for a in arrs]
    for b, F in enumerate(blocks):
        for ax in range(3):
            n = arrs[b].shape[ax] - 1 - 2 * HM
            if ax == 2 and n == 0:
                # 2-D block (ka = 0): every plane -5..5 is a copy of plane 0 (src/parallel.F90:4296-4299)
                if homo[2]:
                    out[b][HM:-HM, HM:-HM, :] = arrs[b][HM:-HM, HM:-HM, HM:HM + 1]
                continue
            lo_nb, hi_nb = F.nb[2 * ax], F.nb[2 * ax + 1]
            if lo_nb < 0 and hi_nb < 0 and homo[ax]:
                lo_nb = hi_nb = b
            idx = [slice(HM, -HM)] * 3
            if lo_nb >= 0:
                nn = arrs[lo_nb].shape[ax] - 1 - 2 * HM
                dst, src = list(idx), list(idx)
                dst[ax], src[ax] = slice(0, HM), slice(nn, nn + HM)                            # -5..-1 <- N-5..N-1
                out[b][tuple(dst)] = arrs[lo_nb][tuple(src)]
            if hi_nb >= 0:
                dst, src = list(idx), list(idx)
                dst[ax], src[ax] = slice(n + HM + 1, n + 2 * HM + 1), slice(HM + 1, 2 * HM + 1)    # N+1..N+5 <- 1..5
                out[b][tuple(dst)] = arrs[hi_nb][tuple(src)]
    return out


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
        self.nb = info["nb"]                     # neighbour block per face (ilo, ihi, jlo, jhi, klo, khi) or -1
        r = info["is_ie"]
        self.lo, self.hi = (r[0], r[2], r[4]), (r[1], r[3], r[5])                 # is, js, ks / ie, je, ke
        self.explicit = False                    # difschm '642e': fds%central is the explicit ladder


def ntype_of(npdc):
    """npdc (1: boundary at 0, 2: boundary at N, 3: interfaces / periodic, 4: both boundaries) is the closure type."""
    return npdc


def gradcal(F):
    """dvel[m][n], dtmp[n] on nodes 0..N."""
    scal = F.vel + [F.tmp]
    nd = ndims_of(F.prs)                       # the k sweep exists for ndims == 3 only (comsolver.F90:418)
    raw = [[core_d(deriv(f, d, ntype_of(F.npdc[d]), F.explicit), d) for d in range(nd)] for f in scal]      # raw[f][d]
    grad = [[sum(raw[f][d] * core(F.dxi[d][n]) for d in range(nd)) for n in range(3)] for f in range(4)]
    return grad[:3], grad[3]


def core_d(a, axis):
    return _core_other_axes(a, axis)


def miucal(T, tempconst):
    return T * np.sqrt(T) * (1.0 + tempconst) / (T + tempconst)


def miucal_dimensional(T):
    """src/fludyna.F90:806-809: Sutherland's law in SI units (nondimen = .false.)."""
    tn = T / 273.15
    return 1.716e-5 * tn * np.sqrt(tn) * (273.15 + 110.4) / (T + 110.4)


def stress_and_heat_flux(F, dvel, dtmp, th):
    """sigma(6) [11 12 13 22 23 33] and qflux(3) on nodes 0..N (diffrsdcal6's pointwise block, nondimensional)."""
    T, v = core(F.tmp), [core(u) for u in F.vel]
    dimensional = th.get("dimensional", False)          # nondimen = .false.: SI viscosity, hcc = cp miu / Pr
    miu = miucal_dimensional(T) if dimensional else miucal(T, th["tempconst"]) / th["reynolds"]
    S = [[0.5 * (dvel[a][b] + dvel[b][a]) for b in range(3)] for a in range(3)]
    skk = (S[0][0] + S[1][1] + S[2][2]) / 3.0
    two_mu = 2.0 * miu
    tau = [[two_mu * (S[a][b] - (skk if a == b else 0.0)) for b in range(3)] for a in range(3)]
    hcc = th["cp"] * miu / th["prandtl"] if dimensional else (miu / th["prandtl"]) / th["const5"]
    qf = [hcc * dtmp[n] + sum(tau[n][b] * v[b] for b in range(3)) for n in range(3)]
    return tau, qf


def rhscal_blocks(blocks, th, homo, diffterm=True, upwind=None):
    """qrhs(5) on nodes 0..N of every block, zero outside the ranges the reference updates.
    upwind: None for the central path (convrsdcal6), or dict(shkcrt, lchardecomp, bfacmpld) for conschm 543c:
    ducrossensor (when lchardecomp) + convrsdcmp (src/solver.F90:224-226)."""

    def box(ranges):
        return tuple(slice(a, b + 1) for a, b in ranges)

    out = []
    if upwind is not None:
        import second_opinion_upwind as U
        lchar = upwind.get("lchardecomp", True)
        flags = [np.ones(F.prs.shape) for F in blocks]
        if lchar:
            ssf = []
            for F in blocks:
                full = np.zeros(F.prs.shape)
                core(full)[...] = U.ducros_ssf(F, gradcal(F)[0])
                ssf.append(full)
            ssf = exchange_halos(ssf, blocks, homo)
            flags = []
            for F, s in zip(blocks, ssf):
                full = np.zeros(F.prs.shape)
                core(full)[...] = U.ducros_flags(s, F.npdc, upwind["shkcrt"])
                flags.append(full)
        for F, fl in zip(blocks, flags):
            conv = U.convrsdcmp(F, th["gamma"], Gas(th).sos if th.get("dimensional") else th["mach"], fl, lchar, upwind.get("bfacmpld", 0.3))
            out.append([-r for r in conv])
    for F in ([] if upwind is not None else blocks):
        qrhs = [np.zeros(core(F.prs).shape) for _ in range(5)]
        # convection: loops over js:je, ks:ke (resp.) and adds on is:ie -> the box is:ie x js:je x ks:ke
        conv_box = box(zip(F.lo, F.hi))
        for d in range(ndims_of(F.prs)):          # solver.F90:2284, :2312: j for ndims >= 2, k for ndims == 3
            U = sum(F.dxi[d][n] * F.vel[n] for n in range(3))
            flux = [F.jacob * F.q[0] * U] + \
                   [F.jacob * (F.q[1 + n] * U + F.dxi[d][n] * F.prs) for n in range(3)] + \
                   [F.jacob * (F.q[4] + F.prs) * U]
            for m in range(5):
                qrhs[m][conv_box] += core_d(deriv(flux[m], d, ntype_of(F.npdc[d]), F.explicit), d)[conv_box]
        out.append([-r for r in qrhs])
    if not diffterm:
        return out
    # diffusion: pointwise stresses per block, one exchange of the 9 fields, then the contracted fluxes
    fields = [[] for _ in range(9)]
    for F in blocks:
        dvel, dtmp = gradcal(F)
        tau, qf = stress_and_heat_flux(F, dvel, dtmp, th)
        for slot, a in enumerate([tau[0][0], tau[0][1], tau[0][2], tau[1][1], tau[1][2], tau[2][2]] + qf):
            full = np.zeros(F.prs.shape)
            core(full)[...] = a
            fields[slot].append(full)
    fields = [exchange_halos(f, blocks, homo) for f in fields]
    sym = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
    for b, F in enumerate(blocks):
        tauh = [[fields[sym[(min(a, c), max(a, c))]][b] for c in range(3)] for a in range(3)]
        qfh = [fields[6 + n][b] for n in range(3)]
        n_nodes = [s - 1 for s in core(F.prs).shape]
        for d in range(ndims_of(F.prs)):
            # loops over 0:N of the two other directions, adds on (is:ie | js:je | ks:ke) of direction d only
            ranges = [(0, n_nodes[a]) for a in range(3)]
            ranges[d] = (F.lo[d], F.hi[d])
            dbox = box(ranges)
            cols = [sum(tauh[m][n] * F.dxi[d][n] for n in range(3)) * F.jacob for m in range(3)] + \
                   [sum(qfh[n] * F.dxi[d][n] for n in range(3)) * F.jacob]
            for m in range(4):
                out[b][1 + m][dbox] += core_d(deriv(cols[m], d, ntype_of(F.npdc[d]), F.explicit), d)[dbox]
    return out


def rhscal(F, th, homo, diffterm=True):
    return rhscal_blocks([F], th, homo, diffterm)[0]


def src_chan(blocks, ys, force):
    """src_chan (src/solver.F90:295-353): trapezoidal bulk integrals of q1..q4 over y, summed over cells i = 1..N,
    k = 1..N of every block (psum), then force * J on the momentum and force . u_bulk * J on the energy equation.
    ys[b]: halo'd y coordinate of block b.  Returns the increments of qrhs(5) per block (nodes 0..N)."""
    tot = np.zeros(4)
    for F, y in zip(blocks, ys):
        yc = core(y)
        ks = slice(0, 1) if ndims_of(F.prs) == 2 else slice(1, None)       # ndims == 2: k1 = k2 = 0
        dy = (yc[1:, 1:, ks] - yc[1:, :-1, ks])
        for m in range(4):
            qm = core(F.q[m])
            tot[m] += np.sum(0.5 * (qm[1:, :-1, ks] + qm[1:, 1:, ks]) * dy)
    ubulk = tot[1:] / tot[0]
    out = []
    for F in blocks:
        J = core(F.jacob)
        out.append([np.zeros(J.shape)] + [force[n] * J for n in range(3)] +
                   [(force[0] * ubulk[0] + force[1] * ubulk[1] + force[2] * ubulk[2]) * J])
    return out


# ------------------------------------------------------------------------------------------------
# per-step diagnostics (raw block sums / maxima before the reference's normalisation)
# ------------------------------------------------------------------------------------------------
def tgv_sums(blocks, th):
    """kenergycal, enstophycal, diss_rate_cal (src/statistic.F90:938-990, :871-936, :994-1046): sums over cells
    1..N of every block of rho |u|^2, rho |curl u|^2 and 2 mu (S:S - div^2/3)."""
    out = np.zeros(3)
    inner = (slice(1, None),) * 3
    for F in blocks:
        dvel, _ = gradcal(F)
        g = np.array([[dvel[a][b][inner] for b in range(3)] for a in range(3)])       # g[a][b] = d u_a / d x_b
        rho, T = core(F.rho)[inner], core(F.tmp)[inner]
        u2 = sum(core(v)[inner] ** 2 for v in F.vel)
        curl2 = (g[2, 1] - g[1, 2]) ** 2 + (g[0, 2] - g[2, 0]) ** 2 + (g[1, 0] - g[0, 1]) ** 2
        S = 0.5 * (g + np.swapaxes(g, 0, 1))
        div = S[0, 0] + S[1, 1] + S[2, 2]
        miu = miucal_dimensional(T) if th.get("dimensional") else miucal(T, th["tempconst"]) / th["reynolds"]
        out += [np.sum(rho * u2), np.sum(rho * curl2), np.sum(2.0 * miu * ((S * S).sum(axis=(0, 1)) - div ** 2 / 3.0))]
    return out


def cfl_maxima(blocks, th):
    """cflcal (src/commcal.F90:27-74): max over nodes 0..N and blocks of the contravariant speeds U, U -+ c |grad xi|
    (and 0) per direction; nondimensional sound speed sqrt(T)/M (src/fludyna.F90 sos)."""
    out = np.zeros(3)
    for F in blocks:
        c = np.sqrt(core(F.tmp)) / th["mach"]
        for d in range(3):
            U = sum(core(F.dxi[d][n]) * core(F.vel[n]) for n in range(3))
            cs = c * np.sqrt(sum(core(F.dxi[d][n]) ** 2 for n in range(3)))
            out[d] = max(out[d], 0.0, U.max(), (U + cs).max(), (U - cs).max())
    return out


def channel_sums(blocks, ys, th, jsize):
    """massfluxchan and fbcxchan (src/statistic.F90:1437-1476, :1303-1367) before the division by ia*ka: the
    trapezoidal integral of rho u over y on cells 1..N, and mu du/dy at the two walls (lower wall +, upper wall -)
    summed over i = 1..N, k = 1..N of the blocks that own a wall."""
    mf = fb = 0.0
    for F, y in zip(blocks, ys):
        yc, q2 = core(y), core(F.q[1])
        mf += np.sum(0.5 * (q2[1:, 1:, 1:] + q2[1:, :-1, 1:]) * (yc[1:, 1:, 1:] - yc[1:, :-1, 1:]))
        dvel, _ = gradcal(F)
        mu = miucal(core(F.tmp), th["tempconst"]) / th["reynolds"]
        if F.nb[2] < 0:                                   # jrk == 0: the lower wall
            fb += np.sum((mu * dvel[0][1])[1:, 0, 1:])
        if F.nb[3] < 0:                                   # jrk == jrkm
            fb -= np.sum((mu * dvel[0][1])[1:, -1, 1:])
    return mf, fb
