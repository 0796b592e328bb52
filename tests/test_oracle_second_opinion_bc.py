"""oracle/solver.cpp's `boucon` (inflow 11, outflow 21, noslip 41, farfield 51, slipadibwall 421 on every face the
reference implements them for) against the second, face-generic transcription tests/second_opinion_bc.py.  The state
is chosen so that every branch runs: far-field faces with flow entering and leaving, the jmax outflow with super- and
subsonic nodes, an inflow profile from Mach 0.2 to 2.2 (both ends of the tanh blend)."""
import numpy as np
import pytest

import second_opinion_bc as B
import second_opinion_rhs as R
from gpu_common import stretched_x

GAMMA, MACH, DT = 1.4, 0.5, 1e-3
NAMES = [f"q{m + 1}" for m in range(5)] + ["rho", "u", "v", "w", "prs", "tmp"]


@pytest.mark.parametrize("n,homo,bctype,twall,blocks", [
    ((14, 12, 10), (False, False, False), (11, 21, 421, 51, 51, 51), (0,) * 6, (1, 1, 1)),
    ((14, 12, 10), (False, False, False), (11, 21, 51, 421, 51, 51), (0,) * 6, (1, 1, 1)),
    ((14, 12, 10), (False, False, True), (11, 21, 41, 21, 1, 1), (0, 0, 1.05, 0, 0, 0), (1, 1, 1)),
    ((20, 16, 10), (False, False, True), (11, 21, 41, 51, 1, 1), (0, 0, 1.05, 0, 0, 0), (2, 2, 1))])
def test_boucon(oracle, n, homo, bctype, twall, blocks):
    c = oracle.Case(*n, homo=homo, blocks=blocks, mach=MACH, deltat=DT)
    c.set_bc(bctype, twall)
    x = stretched_x(n, homo)
    for ib in range(c.nblocks):
        info = c.block_info(ib)
        g0, dims = info["g0"], (info["im"], info["jm"], info["km"])
        c.set_x(np.asfortranarray(x[tuple(slice(g, g + d + 1) for g, d in zip(g0, dims))]), ib)
    c.gridgeom(); c.tgvini()
    th = dict(mach=MACH, gamma=GAMMA, const1=1.0 / (GAMMA * (GAMMA - 1.0) * MACH ** 2), const2=GAMMA * MACH ** 2,
              const6=1.0 / (GAMMA - 1.0))
    inflow = []
    for ib in range(c.nblocks):
        X = [c.get(f"x{d + 1}", ib) for d in range(3)]
        fields = dict(rho=1.0 + 0.1 * np.sin(X[0]) * np.cos(X[1]), tmp=1.0 + 0.05 * np.cos(X[1]) * np.sin(X[2] + 0.3),
                      u=1.0 + 0.6 * np.sin(X[1]) * np.cos(X[2]), v=2.6 * np.sin(X[0] + 0.4) * np.cos(0.5 * X[2]),
                      w=0.8 * np.sin(X[0]) * np.cos(X[1] + 0.2))
        for nm, a in fields.items():
            c.set(nm, np.asfortranarray(a), ib)
        info = c.block_info(ib)
        jm, km = info["jm"], info["km"]
        yy = (np.arange(jm + 1) + info["g0"][1]) / n[1]
        vel_in = np.zeros((jm + 1, km + 1, 3), order="F")
        vel_in[:, :, 0] = (0.4 + 4.0 * yy)[:, None] * (1.0 + 0.02 * np.cos(np.arange(km + 1)))[None, :]   # css = 2
        vel_in[:, :, 1] = 0.05 * np.sin(3 * yy)[:, None]
        vel_in[:, :, 2] = 0.02
        tmp_in = np.asfortranarray(1.0 + 0.05 * yy[:, None] * np.ones((1, km + 1)))
        tmp_prof = 1.0 + 0.05 * yy
        c.set_inflow(vel_in, tmp_in, tmp_prof, ib)
        inflow.append((vel_in, tmp_in, tmp_prof))
    c.updateq(); c.updatefvar()
    state = [R.Fields(c, ib) for ib in range(c.nblocks)]
    # branch coverage of this state (whole domain, before the conditions are applied)
    v_top = np.concatenate([R.core(F.vel[1])[:, -1, :].ravel() for F in state if F.nb[3] < 0])
    T_top = np.concatenate([R.core(F.tmp)[:, -1, :].ravel() for F in state if F.nb[3] < 0])
    if bctype[3] == 21:
        sup = v_top >= np.sqrt(T_top) / MACH
        assert sup.any() and (~sup).any()
    for face in (2, 4, 5):
        if bctype[face] == 51:
            ax, side = face // 2, face % 2
            vn = np.concatenate([np.take(R.core(F.vel[ax]), -1 if side else 0, axis=ax).ravel() for F in state
                                 if F.nb[face] < 0])
            assert (vn > 0).any() and (vn < 0).any(), face
    malo = np.concatenate([d[0][..., 0].ravel() for d in inflow]) / 2.0
    assert malo.min() < 0.5 and malo.max() > 1.5
    c.boucon()
    B.boucon(state, homo, bctype, twall, th, (1.0, 0.0, 0.0, 1.0, c.pinf), DT, inflow)
    for ib, F in enumerate(state):
        got = dict(zip(NAMES, F.q + [F.rho] + F.vel + [F.prs, F.tmp]))
        for nm in NAMES:
            ref = R.core(c.get(nm, ib))
            # measured <= 3e-15
            assert np.abs(R.core(got[nm]) - ref).max() <= 1e-13 * np.abs(ref).max(), (ib, nm)
    c.close()
