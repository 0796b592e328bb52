"""Synthetic inputs of the benchmark configurations (host side, numpy): the grid and
initial fields the reference's setup phase would hand to the time loop.

* ``gridcube``  -- src/gridgeneration.F90:233-265
* ``tgvini``    -- src/initialisation.F90:621-702 + updateq (src/fludyna.F90:254-300)
Arrays are Fortran ordered with hm=5 halos: (-hm:im+hm, -hm:jm+hm, -hm:km+hm[, n]).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

from .parallel import Block

HM = 5


def _shape(block: Block):
    im, jm, km = block.dims
    return (im + 1 + 2 * HM, jm + 1 + 2 * HM, km + 1 + 2 * HM)


def gridcube(block: Block, global_dims: Sequence[int], lengths=(2 * np.pi, 2 * np.pi, 2 * np.pi)) -> np.ndarray:
    """x(i,j,k,1:3) = L/ia * (i+ig0) on nodes 0..im etc.; halos are left to gridsendrecv."""
    shp = _shape(block)
    x = np.zeros(shp + (3,), order="F")
    for d in range(3):
        n = block.dims[d]
        coord = lengths[d] / float(global_dims[d]) * (np.arange(n + 1, dtype=np.float64) + block.g0[d])
        idx = [None, None, None]
        idx[d] = slice(None)
        sl = [slice(HM, HM + block.dims[0] + 1), slice(HM, HM + block.dims[1] + 1), slice(HM, HM + block.dims[2] + 1)]
        x[sl[0], sl[1], sl[2], d] = coord[tuple(idx)]
    return x


def tgvini(x: np.ndarray, thermo: Dict[str, float]):
    """Taylor-Green vortex on nodes 0..n (halo entries stay 0): returns q, rho, vel, prs, tmp."""
    shp = x.shape[:3]
    core = (slice(HM, shp[0] - HM), slice(HM, shp[1] - HM), slice(HM, shp[2] - HM))
    X, Y, Z = (x[core + (d,)] for d in range(3))
    const1, const2 = thermo["const1"], thermo["const2"]
    roinf = uinf = 1.0
    pinf = roinf * 1.0 / const2                     # src/solver.F90:120
    rho = np.zeros(shp, order="F"); prs = np.zeros(shp, order="F"); tmp = np.zeros(shp, order="F")
    vel = np.zeros(shp + (3,), order="F"); q = np.zeros(shp + (5,), order="F")
    r = np.full(X.shape, roinf)
    u = uinf * np.sin(X) * np.cos(Y) * np.cos(Z)
    v = -uinf * np.cos(X) * np.sin(Y) * np.cos(Z)
    w = np.zeros_like(u)
    p = pinf + r / 16.0 * (uinf * uinf) * (np.cos(2.0 * X) + np.cos(2.0 * Y)) * (np.cos(2.0 * Z) + 2.0)
    t = p / r * const2
    rho[core] = r; prs[core] = p; tmp[core] = t
    vel[core + (0,)] = u; vel[core + (1,)] = v; vel[core + (2,)] = w
    q[core + (0,)] = r; q[core + (1,)] = r * u; q[core + (2,)] = r * v; q[core + (3,)] = r * w
    q[core + (4,)] = r * (t * const1 + 0.5 * (u * u + v * v + w * w))   # fvar2q, fludyna.F90:501-505
    return q, rho, vel, prs, tmp


def grichan(block: Block, global_dims: Sequence[int], lengths=(2 * np.pi, 2.0, np.pi), varc: float = 1.07) -> np.ndarray:
    """Channel grid of `grichan` (src/gridgeneration.F90:272-303): uniform x and z, y stretched towards both
    walls with y = Ly/2 (1 + varc tanh(atanh(1/varc) (2 j/ja - 1))), varc = 1.07 (examples/Channel)."""
    shp = _shape(block)
    x = np.zeros(shp + (3,), order="F")
    var1 = np.arctanh(1.0 / varc)
    for d in range(3):
        n = block.dims[d]
        s = (np.arange(n + 1, dtype=np.float64) + block.g0[d]) / float(global_dims[d])
        coord = lengths[d] * s if d != 1 else 0.5 * lengths[1] * (1.0 + varc * np.tanh(var1 * (2.0 * s - 1.0)))
        idx = [None, None, None]
        idx[d] = slice(None)
        sl = [slice(HM, HM + block.dims[0] + 1), slice(HM, HM + block.dims[1] + 1), slice(HM, HM + block.dims[2] + 1)]
        x[sl[0], sl[1], sl[2], d] = coord[tuple(idx)]
    return x


def chanini(x: np.ndarray, thermo: Dict[str, float], amp: float = 0.05):
    """Laminar Poiseuille profile of `chanini` (src/initialisation.F90:791-799) plus a deterministic sinusoidal
    perturbation that vanishes at the walls (the reference seeds synthetic eddies from a random generator,
    :764-767; a seed-free field keeps the benchmark input reproducible); q from fvar2q with temperature."""
    shp = x.shape[:3]
    core = (slice(HM, shp[0] - HM), slice(HM, shp[1] - HM), slice(HM, shp[2] - HM))
    X, Y, Z = (x[core + (d,)] for d in range(3))
    const1, const2 = thermo["const1"], thermo["const2"]
    eta = Y - 1.0
    env = 1.0 - eta ** 2
    r = np.ones_like(X)
    u = 1.5 * env + amp * env * np.sin(2 * X) * np.cos(4 * Z) * np.cos(np.pi * eta)
    v = amp * env * np.cos(2 * X) * np.sin(4 * Z)
    w = amp * env * np.sin(X + 1.0) * np.sin(2 * Z)
    t = 1.0 + (thermo["gamma"] - 1.0) * thermo["prandtl"] * thermo["mach"] ** 2 / 3.0 * 1.5 * (1.0 - eta ** 4)
    p = r * t / const2
    rho = np.zeros(shp, order="F"); prs = np.zeros(shp, order="F"); tmp = np.zeros(shp, order="F")
    vel = np.zeros(shp + (3,), order="F"); q = np.zeros(shp + (5,), order="F")
    rho[core] = r; prs[core] = p; tmp[core] = t
    vel[core + (0,)] = u; vel[core + (1,)] = v; vel[core + (2,)] = w
    q[core + (0,)] = r; q[core + (1,)] = r * u; q[core + (2,)] = r * v; q[core + (3,)] = r * w
    q[core + (4,)] = r * (t * const1 + 0.5 * (u * u + v * v + w * w))
    return q, rho, vel, prs, tmp
