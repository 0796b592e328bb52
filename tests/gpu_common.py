"""Shared helpers of the GPU parity tests: build an oracle case and a CUDA engine on
IDENTICAL inputs (the oracle's own arrays are uploaded through the C ABI)."""
import numpy as np

import astr_b200
from astr_b200 import RhsEngine, decompose, refcal

HM = 5
PRIMS = ["rho", "u", "v", "w", "prs", "tmp"]
QS = [f"q{n + 1}" for n in range(5)]


def stretched_x(n, homo):
    """Node coordinates of a smoothly distorted grid: tanh stretching in non-periodic
    directions (grichan-like, src/gridgeneration.F90:272-303) and a periodic 3-D
    distortion elsewhere, so that all 9 metric terms are non-trivial."""
    ia, ja, ka = n
    L = 2 * np.pi
    s = [np.arange(m + 1) / m for m in n]
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


def make_pair(oracle, n=(32, 32, 32), homo=(True, True, True), perturb=1e-3, stretch=False, seed=1234,
              lfilter=True, diffterm=True, sutherland_s=110.3, device_metrics=False):
    c = oracle.Case(*n, homo=homo, sutherland_s=sutherland_s)
    c.set_flags(lfilter=lfilter, diffterm=diffterm)
    if stretch:
        c.set_x(stretched_x(n, homo))
    c.gridgeom()
    c.tgvini()
    if perturb:
        rng = np.random.default_rng(seed)
        for name in QS:
            a = c.get(name)
            a *= 1.0 + perturb * rng.standard_normal(a.shape)
            c.set(name, a)
        c.updatefvar()
    block = decompose(n, (1, 1, 1), homo)[0]
    th = refcal(1600.0, 0.1, sutherland_s=sutherland_s)
    eng = RhsEngine(block, n, homo, th, deltat=1e-3, lfilter=lfilter, diffterm=diffterm, device=0)
    if device_metrics:
        x = eng.empty(3)
        for d in range(3):
            x[..., d] = c.get(f"x{d + 1}")
        eng.gridgeom(x)
    else:
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
