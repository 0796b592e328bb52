"""oracle/solver.cpp's crash control (crashcheck, crashfix, crinod_expansion; src/mainloop.F90:709-1198) against
tests/second_opinion_crash.py on a 2 x 2 x 1 block grid with physical boundaries in i and j: sick nodes in the
interior, next to each other (a repair feeds the next one), at a block interface, on a physical boundary (neighbours
outside the global domain are skipped in i and j) and at the periodic k ends (k is not tested: halos are read)."""
import numpy as np

import second_opinion_crash as C
import second_opinion_rhs as R
from gpu_common import stretched_x

HM = 5
NAMES = [f"q{m + 1}" for m in range(5)] + ["rho", "u", "v", "w", "prs", "tmp"]


def test_crash_control(oracle):
    n, blocks, homo = (20, 16, 10), (2, 2, 1), (False, False, True)
    gamma, mach = 1.4, 0.1
    c = oracle.Case(*n, blocks=blocks, homo=homo, mach=mach)
    x = stretched_x(n, homo)
    g0s = []
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        g0s.append(g0)
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); c.tgvini()
    c.updatefvar(); c.qswap()                    # halos hold neighbour data, as in a running simulation
    # (block, i, j, k, field, value)
    poison = [(0, 3, 3, 3, "q1", -1.0), (0, 4, 3, 3, "q1", -2.0), (0, 5, 3, 3, "q5", float("nan")),
              (0, 10, 4, 5, "q1", 1e-7),                 # node on the interface to block 1 (im = 10): small density
              (1, 0, 4, 5, "q1", 1e-7),                  # the same global node, owned by block 1 too
              (2, 0, 8, 0, "q1", -0.5),                  # on the physical boundary imin of block 2, periodic k end
              (3, 10, 8, 10, "q5", -3.0),                # corner imax / jmax, other k end: negative pressure
              (1, 6, 0, 2, "q1", -1.0)]                  # on the physical boundary jmin
    for ib, i, j, k, nm, val in poison:
        a = c.get(nm, ib)
        a[i + HM, j + HM, k + HM] = val
        c.set(nm, a, ib)
    c.updatefvar()
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    crinod = [c.get("crinod", ib) for ib in range(c.nblocks)]
    th = dict(mach=mach, gamma=gamma)
    # detection
    want = c.crashcheck()
    got = C.crashcheck(state, crinod)
    assert sum(got) == want == 4                 # the four nodes with a negative density
    for ib in range(c.nblocks):
        np.testing.assert_array_equal(R.core(crinod[ib]), R.core(c.get("crinod", ib)))
    # repair
    want = c.crashfix()
    got = C.crashfix(state, crinod, th, n, g0s)
    assert sum(got) == want == len(poison)
    for ib, F in enumerate(state):
        vals = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
        for nm in NAMES:
            ref = R.core(c.get(nm, ib))
            assert np.all(np.isfinite(ref))
            np.testing.assert_allclose(R.core(vals[nm]), ref, rtol=1e-13, atol=1e-15, err_msg=f"{nm} block {ib}")
        np.testing.assert_array_equal(R.core(crinod[ib]), R.core(c.get("crinod", ib)))
    # dilation + exchange
    want = c.crinod_expansion()
    got = C.crinod_expansion(state, crinod, homo)
    assert sum(got) == want
    inner = (slice(HM - 2, -(HM - 2)),) * 3
    for ib in range(c.nblocks):
        np.testing.assert_array_equal(crinod[ib][inner], c.get("crinod", ib)[inner])
    c.close()
