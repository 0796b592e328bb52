"""Shared helpers of the GPU parity tests: build an oracle case and a CUDA engine on
IDENTICAL inputs (the oracle's own arrays are uploaded through the C ABI)."""
import numpy as np

import astr_b200
from astr_b200 import RhsEngine, decompose, refcal, refcal_dimensional

HM = 5
PRIMS = ["rho", "u", "v", "w", "prs", "tmp"]
QS = [f"q{n + 1}" for n in range(5)]


def stretched_x(n, homo):
    """Node coordinates of a smoothly distorted grid: tanh stretching in non-periodic
    directions (grichan-like, src/gridgeneration.F90:272-303) and a periodic 3-D
    distortion elsewhere, so that all 9 metric terms are non-trivial."""
    ia, ja, ka = n
    L = 2 * np.pi
    s = [np.arange(m + 1) / max(m, 1) for m in n]      # ka == 0: 2-D block, one plane
    S = np.meshgrid(*s, indexing="ij")
    X = []
    for d in range(3):
        if homo[d]:
            X.append(L * S[d])
        else:
            X.append(L * 0.5 * (1.0 + np.tanh(1.07 * (2 * S[d] - 1)) / np.tanh(1.07)))
    amp = 0.08
    o = [np.sin(2 * np.pi * S[(d + 1) % 3]) * np.sin(2 * np.pi * S[(d + 2) % 3]) if homo[(d + 1) % 3] and homo[(d + 2) % 3]
         else 0.0 * S[0] for d in range(3)]
    x = np.stack([X[d] + amp * o[d] for d in range(3)], axis=-1)
    return np.asfortranarray(x)


def skewed_x(n, homo):
    """Sheared lattice with smooth 1-D stretching: y += 0.3 x, z += -0.2 x + 0.15 y.  Every metric
    component is either exactly zero or O(1) everywhere, which keeps chardecomp's pivot division
    (rgp = 1/gpd, src/solver.F90:2052) well conditioned; a periodic direction stays periodic up to
    a constant coordinate jump, which gridsendrecv carries as relative offsets."""
    L = 2 * np.pi
    s = [np.arange(m + 1) / max(m, 1) for m in n]
    S = np.meshgrid(*s, indexing="ij")
    X = [L * (S[d] + 0.04 * np.sin(2 * np.pi * S[d])) if homo[d]
         else L * 0.5 * (1.0 + np.tanh(1.07 * (2 * S[d] - 1)) / np.tanh(1.07)) for d in range(3)]
    x = np.stack([X[0], X[1] + 0.3 * X[0], X[2] - 0.2 * X[0] + 0.15 * X[1]], axis=-1)
    return np.asfortranarray(x)


def channel_x(n, lengths):
    """grichan (src/gridgeneration.F90:272-303): uniform x,z, tanh-stretched y, varc=1.07."""
    lx, ly, lz = lengths
    varc = 1.07
    var1 = np.arctanh(1.0 / varc)
    s = [np.arange(m + 1) / m for m in n]
    S = np.meshgrid(*s, indexing="ij")
    x = np.stack([lx * S[0], 0.5 * ly * (1.0 + varc * np.tanh(var1 * (2.0 * S[1] - 1.0))), lz * S[2]], axis=-1)
    return np.asfortranarray(x)


def channel_state(c, th, amp=0.05, ib=0):
    """Laminar Poiseuille profile of chanini (src/initialisation.F90:791-799) plus a
    deterministic sinusoidal perturbation (the reference's synthetic-eddy seeding is random);
    q from fvar2q with temperature (src/fludyna.F90:312-376)."""
    X, Y, Z = (core(c.get(f"x{d + 1}", ib)) for d in range(3))
    eta = Y - 1.0
    rho = np.ones_like(X)
    env = 1.0 - eta ** 2
    u = 1.5 * env + amp * env * np.sin(2 * X) * np.cos(4 * Z) * np.cos(np.pi * eta)
    v = amp * env * np.cos(2 * X) * np.sin(4 * Z)
    w = amp * env * np.sin(X + 1.0) * np.sin(2 * Z)
    tmp = 1.0 + (th["gamma"] - 1.0) * th["prandtl"] * th["mach"] ** 2 / 3.0 * 1.5 * (1.0 - eta ** 4)
    prs = rho * tmp / th["const2"]
    q5 = rho * (tmp * th["const1"] + 0.5 * (u * u + v * v + w * w))
    vals = dict(rho=rho, u=u, v=v, w=w, prs=prs, tmp=tmp, q1=rho, q2=rho * u, q3=rho * v, q4=rho * w, q5=q5)
    for name, a in vals.items():
        full = c.get(name, ib)
        core(full)[...] = a
        c.set(name, full, ib)


def clean_metrics(c):
    """chardecomp picks its pivot with `abs(var1)>1.d-12` on the raw metric (src/solver.F90:2051) and then
    divides by it: metric components that are analytically zero but come out of gridgeom as 1e-12-ish
    rounding residue make the reference itself amplify rounding by 1e12 (both sides then compute noise).
    The upwind parity grids therefore carry exact zeros where the metric vanishes."""
    names = [f"dxi{a + 1}{b + 1}" for a in range(3) for b in range(3)]
    big = max(np.abs(c.get(nm, ib)).max() for nm in names for ib in range(c.nblocks))
    for ib in range(c.nblocks):
        for nm in names:
            v = c.get(nm, ib)
            v[np.abs(v) < 1e-8 * big] = 0.0
            c.set(nm, v, ib)


def auto_shkcrt(c, bfacmpld=0.3, quantile=0.5):
    """Ducros threshold in the middle of the widest gap of the oracle's sensor values around the requested
    quantile, so that no lshock flag sits within rounding distance of the threshold."""
    c.set_upwind(543, True, bfacmpld, 1.0)
    c.qswap(); c.gradcal(); c.ducrossensor()
    # lshock compares the maximum of ssf over the -4..+5 neighbours along each direction with the threshold
    # (src/commcal.F90:286-309): take the quantile of that local-maximum field
    vals = []
    for ib in range(c.nblocks):
        f = c.get("ssf", ib)
        m = np.zeros_like(f)
        for ax in range(3):
            for o in range(-HM + 1, HM + 1):
                m = np.maximum(m, np.roll(f, -o, axis=ax))
        vals.append(core(m).ravel())
    v = np.sort(np.concatenate(vals))
    k0 = min(max(int(quantile * v.size), 200), v.size - 200)
    win = v[k0 - 200:k0 + 200]
    g = int(np.argmax(np.diff(win)))
    return 0.5 * (win[g] + win[g + 1])


def dimensional_state(c, th, ib=0):
    """Smooth SI-unit field around the HBL reference state (Mach 3): q from fvar2q with temperature and cv."""
    X, Y, Z = (core(c.get(f"x{d + 1}", ib)) for d in range(3))
    rho = 0.0180119 * (1.0 + 0.1 * np.sin(X) * np.cos(Y))
    tmp = 226.65 * (1.0 + 0.05 * np.cos(X + 0.3) * np.sin(Y) * np.cos(Z))
    u = 900.0 * (0.6 + 0.2 * np.sin(Y) * np.cos(Z))
    v = 900.0 * 0.1 * np.cos(X) * np.sin(Z + 0.2)
    w = 900.0 * 0.05 * np.sin(X) * np.sin(Y)
    prs = rho * tmp * th["rgas"]
    q5 = rho * (tmp * th["cv"] + 0.5 * (u * u + v * v + w * w))
    vals = dict(rho=rho, u=u, v=v, w=w, prs=prs, tmp=tmp, q1=rho, q2=rho * u, q3=rho * v, q4=rho * w, q5=q5)
    for name, a in vals.items():
        full = c.get(name, ib)
        core(full)[...] = a
        c.set(name, full, ib)


def make_pair(oracle, n=(32, 32, 32), homo=(True, True, True), perturb=1e-3, stretch=False, seed=1234,
              lfilter=True, diffterm=True, sutherland_s=110.3, device_metrics=False, channel=False, explicit=False, upwind=None, open_faces=False,
              dimensional=False, sponge=None, inflow_from_state=False, engine_kw=None, bc=None):
    reynolds, mach = (3000.0, 0.3) if channel else (1600.0, 0.1)    # input.chl / input.tgv
    lengths = (2 * np.pi, 2.0, np.pi) if channel else None
    th = refcal(reynolds, mach, sutherland_s=sutherland_s)
    deltat = 1e-5 if dimensional else 1e-3
    c = oracle.Case(*n, homo=homo, sutherland_s=sutherland_s, reynolds=reynolds, mach=mach, lengths=lengths,
                    deltat=deltat)
    if dimensional:
        # nondimen=f with the reference state of examples/Hypersonic_Boundary_Layer/datin/input.M3
        ref = (226.65, 900.0, 1.0, 0.0180119)
        c.set_dimensional(*ref)
        th = refcal_dimensional(*ref)
    c.set_flags(lfilter=lfilter, diffterm=diffterm)
    c.set_scheme(explicit)
    bctype, twall, force = (1,) * 6, (0.0,) * 6, (0.0, 0.0, 0.0)
    if channel:
        # examples/Channel/datin/input.chl: walls 41 with T_w=1 at jmin/jmax, body force along x
        assert tuple(homo) == (True, False, True)
        bctype, twall, force = (1, 1, 41, 41, 1, 1), (0, 0, 1.0, 1.0, 0, 0), (2.5e-3, 0.0, 1e-4)
        c.set_bc(bctype, twall)
        c.set_flow(1, force)
        c.set_x(channel_x(n, lengths))
    if open_faces:
        # the boundary set of the HBL / SWLBI inputs: inflow (imin), outflow (imax), isothermal wall (jmin),
        # farfield (jmax) -- examples/Hypersonic_Boundary_Layer/datin/input.M3
        assert tuple(homo) == (False, False, True)
        bctype = (11, 21, 41, 21 if open_faces == "outflow_top" else 51, 1, 1)
        twall = (0, 0, 568.89 if dimensional else 1.05, 0, 0, 0)       # input.M3: `41, 568.89d0`
        if open_faces == "swbli":      # examples/SWLBI/datin/input.2d: slip adiabatic wall at jmin
            bctype = (11, 21, 421, 51, 1, 1)
        c.set_bc(bctype, twall)
    if bc is not None:          # explicit (bctype(1:6), twall(1:6)) for faces no named option set covers
        bctype, twall = tuple(bc[0]), tuple(bc[1])
        c.set_bc(bctype, twall)
    if channel:
        pass
    elif stretch == "skew":
        c.set_x(skewed_x(n, homo))
    elif stretch:
        c.set_x(stretched_x(n, homo))
    c.gridgeom()
    if upwind is not None:
        clean_metrics(c)
    if channel:
        channel_state(c, th)
    elif dimensional:
        dimensional_state(c, th)
    else:
        c.tgvini()
    if perturb:
        rng = np.random.default_rng(seed)
        for name in QS:
            a = c.get(name)
            a *= 1.0 + perturb * rng.standard_normal(a.shape)
            c.set(name, a)
        c.updatefvar()
    up_kw = {}
    if upwind is not None:
        # conschm='543c' (convrsdcmp); shkcrt='auto': see auto_shkcrt
        up_kw = dict(conschm=543, lchardecomp=upwind.get("lchardecomp", True), bfacmpld=upwind.get("bfacmpld", 0.3),
                     shkcrt=upwind.get("shkcrt", 0.01))
        if up_kw["shkcrt"] == "auto":
            up_kw["shkcrt"] = auto_shkcrt(c, up_kw["bfacmpld"], upwind.get("quantile", 0.5))
        if "recon_schem" in upwind:
            # conschm='753e': the explicit upwind family (convrsduwd + recons_exp, src/solver.F90:548)
            up_kw.update(conschm=753, conschm_explicit=True, recon_schem=upwind["recon_schem"])
            c.set_upwind_explicit(upwind["recon_schem"], up_kw["lchardecomp"], up_kw["bfacmpld"], up_kw["shkcrt"])
        else:
            c.set_upwind(543, up_kw["lchardecomp"], up_kw["bfacmpld"], up_kw["shkcrt"])
    block = decompose(n, (1, 1, 1), homo)[0]
    eng = RhsEngine(block, n, homo, th, deltat=deltat, lfilter=lfilter, diffterm=diffterm, device=0,
                    flowtype=int(channel), bctype=bctype, twall=twall, explicit=explicit, **up_kw, **(engine_kw or {}))
    eng.set_force(force)
    if open_faces:
        # synthetic inflow data: half of the profile supersonic (blend -> 1), half subsonic (blend -> 0)
        jm, km = n[1], n[2]
        yy = np.arange(jm + 1) / jm
        uref, tref = (900.0, 226.65) if dimensional else (1.0, 1.0)
        prof = (0.2 + 0.8 * (yy > 0.5)) if dimensional else (1.0 + 14.0 * (yy > 0.5))   # sound speed ~302 m/s | 10
        vel_in = np.zeros((jm + 1, km + 1, 3), order="F")
        vel_in[:, :, 0] = uref * prof[:, None] * (1.0 + 0.01 * np.cos(np.arange(km + 1)))[None, :]
        vel_in[:, :, 1] = uref * 0.02 * np.sin(3 * yy)[:, None]
        vel_in[:, :, 2] = uref * 0.01
        tmp_in = np.asfortranarray(tref * (1.0 + 0.05 * yy[:, None] * np.ones((1, km + 1))))
        tmp_prof = tref * (1.0 + 0.05 * yy)
        if inflow_from_state:
            # what profileinflow would hold for a flow that is already established: the state at i=0
            for m, nm in enumerate(("u", "v", "w")):
                vel_in[:, :, m] = core(c.get(nm))[0, :, :]
            tmp_in = np.asfortranarray(core(c.get("tmp"))[0, :, :])
            tmp_prof = tmp_in[:, 0].copy()
        c.set_inflow(vel_in, tmp_in, tmp_prof)
        eng.set_inflow(vel_in, tmp_in, tmp_prof)
        if not inflow_from_state:
            # make part of the jmax face supersonic outwards so that outflow(4)/farfield see both branches
            v = c.get("v")
            v[HM + n[0] // 2:, HM + n[1], :] = 400.0 if dimensional else 12.0
            c.set("v", v)
    if sponge:
        # sponge layers as `spg_imax = 20` of the SWLBI input would define them (spongelayer_define_ijk): the
        # damping coefficient is a smooth ramp towards the face
        for face, width in sponge.items():
            d = face // 2
            dm = n[d]
            beg, end = (block.s[d], block.s[d] + width - 1) if face % 2 == 0 else (block.e[d] - width + 1, block.e[d])
            shape = [block.e[o] - block.s[o] + 1 for o in range(3)]
            shape[d] = width
            ramp = (np.arange(width) + 1.0) / width
            if face % 2 == 0:
                ramp = ramp[::-1]
            idx = [None, None, None]
            idx[d] = slice(None)
            coef = 0.5 * ramp[tuple(idx)] ** 2 * np.ones(shape)
            coef *= 1.0 + 0.1 * np.cos(np.arange(shape[(d + 1) % 3]))[tuple(slice(None) if o == (d + 1) % 3 else None for o in range(3))]
            coef = np.asfortranarray(coef)
            c.set_sponge(face, beg, end, coef)
            eng.set_sponge(face, beg, end, coef)
    x = eng.empty(3)
    for d in range(3):
        x[..., d] = c.get(f"x{d + 1}")
    if device_metrics:
        eng.gridgeom(x)
    else:
        eng.set_grid(x)
        dxi = eng.empty(9).reshape(eng.shape + (3, 3), order="F")
        for a in range(3):
            for b in range(3):
                dxi[..., a, b] = c.get(f"dxi{a + 1}{b + 1}")
        eng.set_metrics(dxi, c.get("jacob"))
    sync_state(c, eng)
    return c, eng


def sync_state(c, eng, names=QS + PRIMS):
    for name in names:
        eng.set(name, c.get(name))


def core(a):
    return a[HM:-HM, HM:-HM, HM:-HM]


# components of one vector field share one scale (w is ~0 in a Taylor-Green vortex)
GROUPS = [["q2", "q3", "q4"], ["u", "v", "w"], ["qrhs2", "qrhs3", "qrhs4"],
          [f"dvel{m + 1}{n + 1}" for m in range(3) for n in range(3)], [f"dtmp{n + 1}" for n in range(3)],
          [f"sigma{n + 1}" for n in range(6)], [f"qflux{n + 1}" for n in range(3)],
          [f"dxi{a + 1}{b + 1}" for a in range(3) for b in range(3)]]


def rel_err(got, ref, region=core, scale=None):
    g, r = region(got), region(ref)
    if scale is None:
        scale = max(np.abs(r).max(), 1e-300)
    return np.abs(g - r).max() / scale


def assert_fields_close(c, eng, names, tol, region=core, what=""):
    """max|gpu-oracle| / max|oracle| per field (vector components: per vector field) <= tol."""
    refs = {name: c.get(name) for name in names}
    scales = {name: max(np.abs(region(refs[name])).max(), 1e-300) for name in names}
    for grp in GROUPS:
        present = [n for n in grp if n in scales]
        if present:
            s = max(scales[n] for n in present)
            for n in present:
                scales[n] = s
    worst = {name: rel_err(eng.get(name), refs[name], region, scales[name]) for name in names}
    bad = {k: v for k, v in worst.items() if not (v <= tol)}
    assert not bad, f"{what}: relative max-norm error above {tol:g}: {bad} (all: {worst})"
    return worst
