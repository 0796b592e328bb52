"""A SECOND, independent restatement of one Runge-Kutta stage of the main solver's `time_integration_rk`
(src/mainloop.F90:396-482) for periodic block grids, in whole-array NumPy on top of tests/second_opinion.py and
tests/second_opinion_rhs.py.  Test infrastructure: it cross-checks the STAGE ORDER and the exchange semantics of
oracle/solver.cpp, which no stored number of the reference pins.

  filterq   src/comsolver.F90:514-632   per direction: dataswap(q, direction) -- a plain copy -- then the filter
  qswap     src/parallel.F90:4848-5318  per direction i, j, k: 5 halo planes + the shared end node AVERAGED, then
                                        q2fvar on the slabs -5..0 and N..N+5 only (interior primitives stay pre-filter)
  qsave = q J at stage 1; rhscal; RK3 update (coefficients :350-366); spongefilter (src/sponge_layer.F90:55-364);
  updatefvar = q2fvar on 0..N
  q2fvar    src/fludyna.F90:545-634 with thermal_3d :136-179 (nondimensional)
Physical boundaries: the isothermal no-slip wall (bctype 41) is restated (`noslip`), the other types are not.
"""
import numpy as np

import second_opinion as so
import second_opinion_rhs as R

HM = 5
RKCOE = [(1.0, 0.0, 1.0), (0.75, 0.25, 0.25), (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0)]


def _sl(ax, s, n_other=slice(HM, -HM)):
    idx = [n_other] * 3
    idx[ax] = s
    return tuple(idx)


def _n(a, ax):
    return a.shape[ax] - 1 - 2 * HM


def _neighbours(b, F, ax, homo):
    lo, hi = F.nb[2 * ax], F.nb[2 * ax + 1]
    if lo < 0 and hi < 0 and homo[ax]:
        lo = hi = b
    return lo, hi


def filter_axis(a, ax, ntype, alfa):
    """compact_filter along `ax` for every pencil of nodes 0..N of the other two directions; in place on 0..N."""
    Fm = np.moveaxis(a, ax, 0)
    shp = Fm.shape
    dim = shp[0] - 1 - 2 * HM
    out = so.compact_filter(Fm.reshape(shp[0], -1), ntype, dim, alfa).reshape((dim + 1,) + shp[1:])
    out = np.moveaxis(out, 0, ax)
    a[_sl(ax, slice(HM, -HM))] = out[_sl(ax, slice(None))]


def filterq(blocks, homo, alfa=0.49):
    for ax in range(R.ndims_of(blocks[0].prs)):       # comsolver.F90:559, :588: j for ndims >= 2, k for ndims == 3
        for m in range(5):
            new = R.exchange_halos([F.q[m] for F in blocks], blocks, [h if a == ax else False for a, h in enumerate(homo)])
            for F, a in zip(blocks, new):
                # exchange_halos treats every direction: keep only this direction's halos
                F.q[m][_sl(ax, slice(0, HM))] = a[_sl(ax, slice(0, HM))]
                F.q[m][_sl(ax, slice(-HM, None))] = a[_sl(ax, slice(-HM, None))]
        for F in blocks:
            for m in range(5):
                filter_axis(F.q[m], ax, R.ntype_of(F.npdc[ax]), alfa)


def q2fvar(F, idx, th):
    rho = F.q[0][idx]
    v = [F.q[1 + n][idx] / rho for n in range(3)]
    prs = (F.q[4][idx] - 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)) / R.Gas(th).const6
    F.rho[idx] = rho
    for n in range(3):
        F.vel[n][idx] = v[n]
    F.prs[idx] = prs
    F.tmp[idx] = R.Gas(th).T_of(prs, rho)


def qswap(blocks, homo, th):
    for ax in range(3):
        pre = [[q.copy() for q in F.q] for F in blocks]          # what the send buffers hold
        for b, F in enumerate(blocks):
            lo, hi = _neighbours(b, F, ax, homo)
            n = _n(F.q[0], ax)
            if ax == 2 and n == 0:
                # 2-D block (ka = 0, src/parallel.F90:5161-5164): planes -5..5 copy plane 0, no averaging, then q2fvar
                # on the two slabs (which are all eleven planes)
                if homo[2]:
                    for m in range(5):
                        F.q[m][HM:-HM, HM:-HM, :] = pre[b][m][HM:-HM, HM:-HM, HM:HM + 1]
                    q2fvar(F, (slice(HM, -HM), slice(HM, -HM), slice(None)), th)
                continue
            for m in range(5):
                if hi >= 0:
                    F.q[m][_sl(ax, slice(n + HM + 1, n + 2 * HM + 1))] = pre[hi][m][_sl(ax, slice(HM + 1, 2 * HM + 1))]
                    F.q[m][_sl(ax, slice(n + HM, n + HM + 1))] = 0.5 * (pre[b][m][_sl(ax, slice(n + HM, n + HM + 1))] +
                                                                      pre[hi][m][_sl(ax, slice(HM, HM + 1))])
                if lo >= 0:
                    nl = _n(pre[lo][m], ax)
                    F.q[m][_sl(ax, slice(0, HM))] = pre[lo][m][_sl(ax, slice(nl, nl + HM))]
                    F.q[m][_sl(ax, slice(HM, HM + 1))] = 0.5 * (pre[b][m][_sl(ax, slice(HM, HM + 1))] +
                                                               pre[lo][m][_sl(ax, slice(nl + HM, nl + HM + 1))])
            if hi >= 0:
                q2fvar(F, _sl(ax, slice(n + HM, n + 2 * HM + 1)), th)
            if lo >= 0:
                q2fvar(F, _sl(ax, slice(0, HM + 1)), th)


def noslip(blocks, homo, bctype, twall, th):
    """boucon -> noslip (bctype 41) on the faces a block owns (src/bc.F90:327-407, :6306-6723): wall pressure
    extrapolated from the two interior nodes (pe = (4 p1 - p2)/3, from the primitives as they stand: quirk Q10), zero
    velocity, wall temperature; density from the gas law, q from fvar2q with pressure (src/fludyna.F90:312-376)."""
    for face in range(6):                               # imin, imax, jmin, jmax, kmin, kmax: the reference's order
        if bctype[face] != 41:
            continue
        ax, side = face // 2, face % 2
        for F in blocks:
            if homo[ax] or F.nb[face] >= 0:
                continue
            n = _n(F.prs, ax)
            w, s = (n, -1) if side else (0, 1)
            at = lambda i: _sl(ax, slice(i + HM, i + HM + 1))
            pe = (4.0 * F.prs[at(w + s)] - F.prs[at(w + 2 * s)]) / 3.0
            for v in F.vel:
                v[at(w)] = 0.0
            F.prs[at(w)] = pe
            F.tmp[at(w)] = twall[face]
            F.rho[at(w)] = R.Gas(th).rho_of(pe, twall[face])
            F.q[0][at(w)] = F.rho[at(w)]
            for m in (1, 2, 3):
                F.q[m][at(w)] = 0.0
            F.q[4][at(w)] = pe * R.Gas(th).const6


def _swap_q(blocks, homo, axes):
    """dataswap(q, direction): plain copies of the 5 halo planes of the given directions."""
    for ax in axes:
        for m in range(5):
            new = R.exchange_halos([F.q[m] for F in blocks], blocks, [h if a == ax else False for a, h in enumerate(homo)])
            for F, a in zip(blocks, new):
                F.q[m][_sl(ax, slice(0, HM))] = a[_sl(ax, slice(0, HM))]
                F.q[m][_sl(ax, slice(-HM, None))] = a[_sl(ax, slice(-HM, None))]


def _damp(F, box, coef):
    """q <- (1 - c) q + c/6 * (sum of the six neighbours) on `box` (node ranges per direction), all five fields from
    the values before the update (the reference's qtemp)."""
    idx = tuple(slice(lo + HM, hi + HM + 1) for lo, hi in box)
    new = []
    for m in range(5):
        q = F.q[m]
        nb = 0.0
        for ax in range(3):
            for off in (1, -1):
                sh = list(idx)
                sh[ax] = slice(idx[ax].start + off, idx[ax].stop + off)
                nb = nb + q[tuple(sh)]
        new.append((1.0 - coef) * q[idx] + coef / 6.0 * nb)
    for m in range(5):
        F.q[m][idx] = new[m]


def spongefilter_layer(blocks, homo, layers):
    """spongefilter_layer (src/sponge_layer.F90:67-319): faces in the order i0, im, jm, k0, km; for each one
    dataswap(q, direction) on every block, then the damped average on the blocks the layer reaches.
    layers: {face: [per block (beg, end, coef) or None]} with face 0 i0, 1 im, 3 jm, 4 k0, 5 km and coef on
    (beg:end along the face's direction) x (is:ie ... of the other two)."""
    for face in (0, 1, 3, 4, 5):
        if face not in layers:
            continue
        ax = face // 2
        _swap_q(blocks, homo, [ax])
        for F, lay in zip(blocks, layers[face]):
            if lay is None or lay[0] < 0:
                continue
            box = [(F.lo[a], F.hi[a]) for a in range(3)]
            box[ax] = (lay[0], lay[1])
            _damp(F, box, lay[2])


def sponge_circle_coefficients(blocks, xs, centre, range_spange, dampfac):
    """spongelayer_define_circle (src/sponge_layer.F90:369-440): squared excess distance beyond range_spange from the
    centre on is:ie x js:je x ks:ke, normalised by its maximum over all blocks (pmax), times dampfac; None for a block
    without a damped node (lsponge_loc)."""
    raw = []
    for F, x in zip(blocks, xs):
        idx = tuple(slice(F.lo[a] + HM, F.hi[a] + HM + 1) for a in range(3))
        dist = np.sqrt(sum((x[m][idx] - centre[m]) ** 2 for m in range(3)))
        raw.append(np.where(dist >= range_spange, (dist - range_spange) ** 2, 0.0))
    big = max(r.max() for r in raw)
    return [r / big * dampfac if (r > 0).any() else None for r in raw]


def spongefilter_global(blocks, homo, coefs):
    """spongefilter_global (src/sponge_layer.F90:321-364): dataswap(q) in every direction when any block has a damped
    node, then the damped average over is:ie x js:je x ks:ke of the blocks that have one."""
    if all(c is None for c in coefs):
        return
    _swap_q(blocks, homo, [0, 1, 2])
    for F, c in zip(blocks, coefs):
        if c is not None:
            _damp(F, [(F.lo[a], F.hi[a]) for a in range(3)], c)


def rk_stage(blocks, rkstep, th, homo, deltat, qsave, alfa=0.49, bctype=None, twall=None, force=None, ys=None,
             upwind=None, bc_extra=None, lfilter=True, rk4=None, sponge=None):
    """One stage; `qsave` is a list (per block) of 5 arrays on 0..N, filled at stage 1.  bctype / twall: boundary
    types per face; only no-slip walls (41) unless bc_extra = dict(free=..., inflow_data=...) brings the data of the
    open types (tests/second_opinion_bc.py); force, ys: the channel's body force and the halo'd y coordinate per
    block (src_chan); upwind: see second_opinion_rhs.rhscal_blocks."""
    if lfilter:
        filterq(blocks, homo, alfa)
    if bctype is not None and bc_extra is not None:
        import second_opinion_bc as B
        B.boucon(blocks, homo, bctype, twall, th, bc_extra["free"], deltat, bc_extra.get("inflow_data"))
    elif bctype is not None:
        noslip(blocks, homo, bctype, twall, th)
    qswap(blocks, homo, th)
    c = (slice(HM, -HM),) * 3
    if rkstep == 1:
        for b, F in enumerate(blocks):
            qsave[b] = [F.q[m][c] * F.jacob[c] for m in range(5)]
    qrhs = R.rhscal_blocks(blocks, th, homo, upwind=upwind)
    if force is not None:
        src = R.src_chan(blocks, ys, force)
        qrhs = [[qrhs[b][m] + src[b][m] for m in range(5)] for b in range(len(blocks))]
    if rk4 is not None:
        # rkscheme = 'rk4' (src/mainloop.F90:368-381, :452-476): rk4 is the caller's list of accumulated right-hand
        # sides per block (rhsav, zeroed at stage 1); stages 1..3 step from qsave by c1 dt qrhs and accumulate
        # c2 qrhs, stage 4 steps by dt/6 (qrhs + rhsav)
        c1, c2 = ((0.5, 1.0), (0.5, 2.0), (1.0, 2.0), (1.0 / 6.0, 1.0))[rkstep - 1]
        for b, F in enumerate(blocks):
            J = F.jacob[c]
            if rkstep == 1:
                rk4[b] = [np.zeros_like(J) for _ in range(5)]
            for m in range(5):
                if rkstep <= 3:
                    F.q[m][c] = (qsave[b][m] + c1 * deltat * qrhs[b][m]) / J
                    rk4[b][m] = rk4[b][m] + c2 * qrhs[b][m]
                else:
                    F.q[m][c] = (qsave[b][m] + c1 * deltat * (qrhs[b][m] + rk4[b][m])) / J
            q2fvar(F, c, th)
        return
    a1, a2, a3 = RKCOE[rkstep - 1]
    for b, F in enumerate(blocks):
        J = F.jacob[c]
        for m in range(5):
            F.q[m][c] = (a1 * qsave[b][m] + a2 * F.q[m][c] * J + a3 * qrhs[b][m] * deltat) / J
    if sponge is not None:                  # call spongefilter (mainloop.F90:478), before updatefvar
        sponge(blocks)
    for F in blocks:
        q2fvar(F, c, th)
