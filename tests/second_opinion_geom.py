"""A SECOND, independent restatement of `gridgeom` for 3-D blocks (src/geom.F90:99-700) in whole-array NumPy on top
of tests/second_opinion.py.  Test infrastructure: it cross-checks oracle/solver.cpp's metrics, which no stored number
of the reference pins ("curvilinear metrics").

Where the reference writes the 18 conservative-form terms one by one (geom.F90:399-519), this restatement evaluates
the identity they spell out,

    J d(xi_a)/d(x_n) = 1/2 sum_{b,c,m,p} eps(a,b,c) eps(n,m,p) d/d(xi_b) [ x_m d(x_p)/d(xi_c) ]

with Levi-Civita loops, so a slip in any one of the 18 lines (of the reference's restatement in the oracle, or of
this reading) breaks the agreement.  The rest follows the reference's order: dx/dxi on 0..N (:130-164),
dataswap(dx), J = det (:352-381), dataswap + datasync of J, the terms above, `/J` (:662-666), dataswap + datasync
of dxi (:668-680).  Coordinate halos (gridsendrecv) are taken from the oracle; geombc's boundary extrapolation is
not restated (halos beyond physical boundaries are not compared).
datasync (src/parallel.F90:3725-3921): per direction i, j, k the shared end nodes take the mean of the two owners.
"""
import numpy as np

import second_opinion_rhs as R

HM = 5


def _eps(i, j, k):
    return (i - j) * (j - k) * (k - i) // 2


def _embed(core_arr, shape):
    full = np.zeros(shape)
    R.core(full)[...] = core_arr
    return full


def datasync(arrs, blocks, homo):
    """Shared-node average, direction by direction (the send buffers of a direction are packed before it changes)."""
    for ax in range(3):
        pre = [a.copy() for a in arrs]
        for b, F in enumerate(blocks):
            lo, hi = F.nb[2 * ax], F.nb[2 * ax + 1]
            if lo < 0 and hi < 0 and homo[ax]:
                lo = hi = b
            n = arrs[b].shape[ax] - 1 - 2 * HM

            def node(a, i):
                idx = [slice(HM, -HM)] * 3
                idx[ax] = slice(i + HM, i + HM + 1)
                return a[tuple(idx)]
            if hi >= 0:
                node(arrs[b], n)[...] = 0.5 * (node(pre[b], n) + node(pre[hi], 0))
            if lo >= 0:
                nl = pre[lo].shape[ax] - 1 - 2 * HM
                node(arrs[b], 0)[...] = 0.5 * (node(pre[b], 0) + node(pre[lo], nl))
    return arrs


def gridgeom(xs, blocks, homo):
    """xs[b][m]: halo'd coordinate x_m of block b.  Returns (jacob[b], dxi[b][a][n]) as halo'd arrays."""
    nb = len(blocks)
    shape = [xs[b][0].shape for b in range(nb)]
    if R.ndims_of(xs[0][0]) == 2:
        return _gridgeom_2d(xs, blocks, homo)
    # dx[b][m][c] = d x_m / d xi_c on 0..N, then halos by exchange
    dx = [[[None] * 3 for _ in range(3)] for _ in range(nb)]
    for m in range(3):
        for c in range(3):
            cores = [R.core_d(R.deriv(xs[b][m], c, R.ntype_of(blocks[b].npdc[c])), c) for b in range(nb)]
            full = R.exchange_halos([_embed(cores[b], shape[b]) for b in range(nb)], blocks, homo)
            for b in range(nb):
                dx[b][m][c] = full[b]
    # Jacobian: the determinant, as a Levi-Civita sum
    jac = []
    for b in range(nb):
        det = sum(_eps(i, j, k) * dx[b][0][i] * dx[b][1][j] * dx[b][2][k]
                  for i in range(3) for j in range(3) for k in range(3) if _eps(i, j, k))
        jac.append(_embed(R.core(det), shape[b]))
    jac = datasync(R.exchange_halos(jac, blocks, homo), blocks, homo)
    # conservative-form metrics
    dxi = [[[None] * 3 for _ in range(3)] for _ in range(nb)]
    for a in range(3):
        for n in range(3):
            cores = []
            for b in range(nb):
                acc = 0.0
                for bb in range(3):
                    for c in range(3):
                        if not _eps(a, bb, c):
                            continue
                        phi = 0.0
                        for m in range(3):
                            for p in range(3):
                                s = _eps(a, bb, c) * _eps(n, m, p)
                                if s:
                                    phi = phi + 0.5 * s * xs[b][m] * dx[b][p][c]
                        acc = acc + R.core_d(R.deriv(phi, bb, R.ntype_of(blocks[b].npdc[bb])), bb)
                cores.append(acc / R.core(jac[b]))
            full = datasync(R.exchange_halos([_embed(cores[b], shape[b]) for b in range(nb)], blocks, homo), blocks, homo)
            for b in range(nb):
                dxi[b][a][n] = full[b]
    return jac, dxi


def _gridgeom_2d(xs, blocks, homo):
    """ndims == 2 (ka = 0; src/geom.F90:130-164, :383-389, :521-528): J = x_xi y_eta - x_eta y_xi and the four metrics
    are the cofactors of the 2 x 2 Jacobian matrix (no conservative form needed in 2-D), the third row and column zero."""
    nb = len(blocks)
    shape = [xs[b][0].shape for b in range(nb)]
    d = [[[R.core_d(R.deriv(xs[b][m], c, R.ntype_of(blocks[b].npdc[c])), c) for c in range(2)] for m in range(2)]
         for b in range(nb)]                                        # d[b][m][c] = d x_m / d xi_c on nodes 0..N
    jac = [_embed(d[b][0][0] * d[b][1][1] - d[b][0][1] * d[b][1][0], shape[b]) for b in range(nb)]
    jac = datasync(R.exchange_halos(jac, blocks, homo), blocks, homo)
    dxi = [[[None] * 3 for _ in range(3)] for _ in range(nb)]
    cof = {(0, 0): lambda D: D[1][1], (0, 1): lambda D: -D[0][1], (1, 0): lambda D: -D[1][0], (1, 1): lambda D: D[0][0]}
    for a in range(3):
        for n in range(3):
            cores = [cof[(a, n)](d[b]) / R.core(jac[b]) if (a, n) in cof else np.zeros(R.core(jac[b]).shape)
                     for b in range(nb)]
            full = datasync(R.exchange_halos([_embed(cores[b], shape[b]) for b in range(nb)], blocks, homo), blocks, homo)
            for b in range(nb):
                dxi[b][a][n] = full[b]
    return jac, dxi
