"""A SECOND reading of the open boundary conditions of `boucon` (src/bc.F90:327-407) in whole-face NumPy: `inflow`
(11, imin), `outflow` (21, imax and jmax), `farfield` (51, jmin / jmax / kmin / kmax), `slipadibwall` (421, jmin / jmax),
next to `noslip` (41) in tests/second_opinion_stage.py.  Test infrastructure: cross-checks oracle/solver.cpp's `boucon`
(SURVEY.md 8f-1), which no stored number of the reference pins.

These routines are pointwise algebra on one plane, so unlike the other second opinions there is no independent
formulation to be had -- what this file adds is a second, separately made transcription, written face-generically:
one function per boundary TYPE with the face's normal direction and side (s = +1 for a low face, -1 for a high face)
as parameters, where the reference (and the oracle) spell every face out.  A sign or index slip in any one face of
either transcription breaks the agreement.

  inflow        src/bc.F90:1366-1562   subsonic / supersonic blend on the Mach number of the inflow profile
  outflow       src/bc.F90:3404-3617   imax: copy of the neighbour plane; jmax: supersonic copy or relaxed pressure wave
  farfield      src/bc.F90:3008-3392   jmin, kmin, kmax: characteristic in / outflow against the free stream; jmax: extrapolation
  slipadibwall  src/bc.F90:7231-7430   jmin: tangential u extrapolated, v = w = 0; jmax: u AND v extrapolated, w = 0 (as written)
  extrapolate   src/commfunc.F90:277-285   (4 v1 - v2 - 2 dv)/3
  fvar2q        src/fludyna.F90:312-376    energy from temperature (const1) or from pressure (const6)
  thermal       src/fludyna.F90:136-179    nondimensional gas law with const2
Both gas modes (tests/second_opinion_rhs.py::Gas).  Every routine works on nodes 0..N of the two in-plane directions of a face the block owns.
"""
import numpy as np

HM = 5


def _plane(a, ax, node):
    """View of plane `node` (node index along `ax`), nodes 0..N of the other two directions."""
    idx = [slice(HM, -HM)] * 3
    idx[ax] = node + HM
    return a[tuple(idx)]


def _ext(a, ax, w, s):
    """extrapolate(v1, v2, dv = 0) from the two interior neighbours of wall node w."""
    return (4.0 * _plane(a, ax, w + s) - _plane(a, ax, w + 2 * s)) / 3.0


def _sos(T, th):
    return _gas(th).sos(T)


def _gas(th):
    import second_opinion_rhs as R
    return R.Gas(th)


def _store(F, ax, w, rho, vel, prs, tmp, th, energy_from):
    _plane(F.rho, ax, w)[...] = rho
    for n in range(3):
        _plane(F.vel[n], ax, w)[...] = vel[n]
    _plane(F.prs, ax, w)[...] = prs
    _plane(F.tmp, ax, w)[...] = tmp
    k = 0.5 * (vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2)
    _plane(F.q[0], ax, w)[...] = rho
    for n in range(3):
        _plane(F.q[1 + n], ax, w)[...] = rho * vel[n]
    g = _gas(th)
    _plane(F.q[4], ax, w)[...] = rho * (tmp * g.cotem + k) if energy_from == "temperature" else prs * g.const6 + rho * k


def _cur(F, ax, w):
    return (_plane(F.rho, ax, w).copy(), [_plane(v, ax, w).copy() for v in F.vel], _plane(F.prs, ax, w).copy(),
            _plane(F.tmp, ax, w).copy())


def inflow_imin(F, th, vel_in, tmp_in, tmp_prof, pinf):
    """vel_in(j,k,3), tmp_in(j,k), tmp_prof(j): the inflow data of alloinflow."""
    ax, w, s = 0, 0, 1
    rho_ref = _plane(F.rho, ax, 1)
    css = _sos(tmp_prof, th)[:, None]
    pe, ue = _ext(F.prs, ax, w, s), _ext(F.vel[0], ax, w, s)
    blend = 0.5 * (np.tanh((vel_in[..., 0] / css - 1.0) * 6.0) + 1.0)
    prs = (0.5 * (pinf + pe) + 0.5 * rho_ref * css * (vel_in[..., 0] - ue)) * (1.0 - blend) + pinf * blend
    u = vel_in[..., 0] + (pinf - prs) / rho_ref / css
    rho = _gas(th).rho_of(prs, tmp_in)
    _store(F, ax, w, rho, [u, vel_in[..., 1], vel_in[..., 2]], prs, tmp_in, th, "temperature")


def outflow(F, th, ax, pinf, deltat):
    """High face of direction `ax` (0: imax, 1: jmax)."""
    w, s = F.prs.shape[ax] - 1 - 2 * HM, -1
    rho0, vel0, prs0, tmp0 = _cur(F, ax, w)
    if ax == 0:        # every primitive copied from the neighbour plane, density from the gas law
        vel = [_plane(v, ax, w + s).copy() for v in F.vel]
        prs, tmp = _plane(F.prs, ax, w + s).copy(), _plane(F.tmp, ax, w + s).copy()
        _store(F, ax, w, _gas(th).rho_of(prs, tmp), vel, prs, tmp, th, "temperature")
        return
    css = _sos(tmp0, th)
    ve = [_ext(v, ax, w, s) for v in F.vel]
    pe, te, roe = _ext(F.prs, ax, w, s), _ext(F.tmp, ax, w, s), _ext(F.rho, ax, w, s)
    alpha = 0.25
    sup = vel0[ax] >= css
    pwave = (prs0 + alpha * deltat * pinf + rho0 * css * (ve[ax] - vel0[ax])) / (1.0 + alpha * deltat)
    prs = np.where(sup, pe, pwave)
    rho = np.where(sup, roe, _gas(th).rho_of(pwave, te))
    _store(F, ax, w, rho, ve, prs, _gas(th).T_of(prs, rho), th, "pressure")


def farfield(F, th, ax, side, free):
    """free = (uinf, vinf, winf, roinf, pinf).  side 0: low face (s = +1), 1: high face (s = -1)."""
    n = F.prs.shape[ax] - 1 - 2 * HM
    w, s = (n, -1) if side else (0, 1)
    vinf, roinf, pinf = free[:3], free[3], free[4]
    rho0, vel0, prs0, tmp0 = _cur(F, ax, w)
    ve = [_ext(v, ax, w, s) for v in F.vel]
    pe, roe = _ext(F.prs, ax, w, s), _ext(F.rho, ax, w, s)
    if ax == 1 and side == 1:          # jmax: plain extrapolation, energy from the temperature
        _store(F, ax, w, roe, ve, pe, _gas(th).T_of(pe, roe), th, "temperature")
        return
    css = _sos(tmp0, th)
    csse = (4.0 * _sos(_plane(F.tmp, ax, w + s), th) - _sos(_plane(F.tmp, ax, w + 2 * s), th)) / 3.0
    entering = s * vel0[ax] >= 0.0
    # flow entering the domain: the incoming acoustic wave carries the free stream
    p_in = 0.5 * (pinf + pe) + s * 0.5 * rho0 * css * (vinf[ax] - ve[ax])
    vn_in = s * 0.5 * (pinf - pe) / (rho0 * css) + 0.5 * (vinf[ax] + ve[ax])
    rho_in = roinf * (p_in / pinf) ** (1.0 / th["gamma"])
    # flow leaving: free-stream pressure, entropy and tangential velocity from inside
    rho_out = roe + (pinf - pe) / csse / csse
    vn_out = ve[ax] - s * (pe - pinf) / roe / csse
    vel = []
    for m in range(3):
        if m == ax:
            vel.append(np.where(entering, vn_in, vn_out))
        else:
            vel.append(np.where(entering, vinf[m], ve[m]))
    prs = np.where(entering, p_in, pinf)
    rho = np.where(entering, rho_in, rho_out)
    _store(F, ax, w, rho, vel, prs, _gas(th).T_of(prs, rho), th, "pressure")


def slipadibwall(F, th, side):
    """jmin (side 0) / jmax (side 1)."""
    ax = 1
    n = F.prs.shape[ax] - 1 - 2 * HM
    w, s = (n, -1) if side else (0, 1)
    pe, te, ue = _ext(F.prs, ax, w, s), _ext(F.tmp, ax, w, s), _ext(F.vel[0], ax, w, s)
    v = _ext(F.vel[1], ax, w, s) if side else np.zeros_like(pe)        # the reference keeps v at jmax
    _store(F, ax, w, _gas(th).rho_of(pe, te), [ue, v, np.zeros_like(pe)], pe, te, th, "pressure")


def boucon(blocks, homo, bctype, twall, th, free, deltat, inflow_data=None):
    """The reference's loop over the six faces (n = 1..6), each type on the blocks that own the face."""
    import second_opinion_stage as S
    for face in range(6):
        ax, side = face // 2, face % 2
        bt = bctype[face]
        if bt == 41:
            one = [0] * 6
            one[face] = 41
            S.noslip(blocks, homo, one, twall, th)
            continue
        for b, F in enumerate(blocks):
            if homo[ax] or F.nb[face] >= 0:
                continue
            if bt == 11:
                assert face == 0, "inflow exists for imin only"
                inflow_imin(F, th, *inflow_data[b], free[4])
            elif bt == 21:
                assert face in (1, 3), "outflow exists for imax and jmax only"
                outflow(F, th, ax, free[4], deltat)
            elif bt == 51:
                assert face >= 2, "farfield exists for the j and k faces only"
                farfield(F, th, ax, side, free)
            elif bt == 421:
                assert face in (2, 3)
                slipadibwall(F, th, side)
            elif bt != 1:
                raise NotImplementedError(bt)
