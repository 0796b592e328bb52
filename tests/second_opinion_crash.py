"""A SECOND restatement of the crash control's node operations (src/mainloop.F90:709-1198, `lcracon`): `crashcheck`
(detection), `crashfix` (repair in storage order) and `crinod_expansion` (dilation + exchange), in NumPy with an
explicit Python loop where the reference's loop order matters.  Test infrastructure: cross-checks oracle/solver.cpp.
No immersed boundary: `nodestat` is "fluid" everywhere.  `databakup` is bookkeeping of two copies and is covered by
tests/test_oracle_solver.py::test_crash_control_restatement."""
import numpy as np

import second_opinion_rhs as R

HM = 5


def crashcheck(blocks, crinod):
    """Nodes 0..N whose density q1 is not >= 0 (NaN included) become critical; returns their number per block."""
    counts = []
    for F, cn in zip(blocks, crinod):
        bad = ~(R.core(F.q[0]) >= 0.0)
        R.core(cn)[bad] = 1.0
        counts.append(int(bad.sum()))
    return counts


def crashfix(blocks, crinod, th, dims, g0s):
    """Sick nodes (rho, p or T below 1e-5-based thresholds, NaN included) are visited in storage order (i fastest);
    each becomes critical and takes the mean q of its 26 neighbours that lie inside the global domain in i and j
    (k is not tested) and have non-negative rho, p, T -- AS THEY ARE AT THAT MOMENT, earlier repairs included -- and
    its primitives are rebuilt; a node without an admissible neighbour stays as it is.  Returns the repaired counts."""
    gas = R.Gas(th)
    eps_rho = eps_tmp = 1.0e-5
    eps_prs = eps_rho * eps_tmp / gas.const2 if not gas.dim else eps_rho * eps_tmp * gas.rgas     # thermal(rho, T)
    ia, ja = dims[0], dims[1]
    counts = []
    for F, cn, g0 in zip(blocks, crinod, g0s):
        rho, prs, tmp = F.rho, F.prs, F.tmp
        ok = (R.core(rho) >= eps_rho) & (R.core(prs) >= eps_prs) & (R.core(tmp) >= eps_tmp)
        sick = np.argwhere(~ok)
        order = np.lexsort((sick[:, 0], sick[:, 1], sick[:, 2]))          # k slowest, then j, then i
        n = 0
        for i, j, k in sick[order]:
            cn[i + HM, j + HM, k + HM] = 1.0
            acc, norm = np.zeros(5), 0
            for kk in (-1, 0, 1):
                for jj in (-1, 0, 1):
                    for ii in (-1, 0, 1):
                        if ii == jj == kk == 0:
                            continue
                        if not (0 <= g0[0] + i + ii <= ia and 0 <= g0[1] + j + jj <= ja):
                            continue
                        at = (i + ii + HM, j + jj + HM, k + kk + HM)
                        if rho[at] >= 0.0 and prs[at] >= 0.0 and tmp[at] >= 0.0:
                            acc += [F.q[m][at] for m in range(5)]
                            norm += 1
            if norm >= 1:
                at = (i + HM, j + HM, k + HM)
                for m in range(5):
                    F.q[m][at] = acc[m] / float(norm)
                r = F.q[0][at]
                v = [F.q[1 + a][at] / r for a in range(3)]
                p = (F.q[4][at] - 0.5 * r * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)) / gas.const6
                F.rho[at], F.prs[at], F.tmp[at] = r, p, gas.T_of(p, r)
                for a in range(3):
                    F.vel[a][at] = v[a]
                n += 1
        counts.append(n)
    return counts


def crinod_expansion(blocks, crinod, homo):
    """Every critical node of -1..N+1 marks its 27-neighbourhood; the result replaces crinod on -2..N+2, then
    dataswap(crinod).  Returns the reference's counter per block (27 per critical node)."""
    counts, out = [], []
    for cn in crinod:
        src = cn[HM - 1:-(HM - 1), HM - 1:-(HM - 1), HM - 1:-(HM - 1)] > 0.5       # nodes -1..N+1
        counts.append(27 * int(src.sum()))
        grown = np.zeros(tuple(s + 2 for s in src.shape), dtype=bool)             # nodes -2..N+2
        for di in range(3):
            for dj in range(3):
                for dk in range(3):
                    grown[di:di + src.shape[0], dj:dj + src.shape[1], dk:dk + src.shape[2]] |= src
        new = cn.copy()
        new[HM - 2:-(HM - 2), HM - 2:-(HM - 2), HM - 2:-(HM - 2)] = grown.astype(float)
        out.append(new)
    out = R.exchange_halos(out, blocks, homo)
    for cn, new in zip(crinod, out):
        cn[...] = new
    return counts
