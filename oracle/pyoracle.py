"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module (see oracle/astr_oracle.hpp).
The product package ``astr_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libastr_oracle.so")
HM = 5

# field ids of oracle_case_get / oracle_case_set (oracle/solver.cpp: field_by_id)
FIELD_IDS = {
    **{f"q{n + 1}": n for n in range(5)},
    "rho": 5, "u": 6, "v": 7, "w": 8, "prs": 9, "tmp": 10,
    **{f"qrhs{n + 1}": 11 + n for n in range(5)},
    "jacob": 16,
    **{f"dxi{a + 1}{b + 1}": 17 + 3 * a + b for a in range(3) for b in range(3)},
    **{f"dvel{m + 1}{n + 1}": 26 + 3 * m + n for m in range(3) for n in range(3)},
    **{f"dtmp{n + 1}": 35 + n for n in range(3)},
    **{f"sigma{n + 1}": 38 + n for n in range(6)},
    **{f"qflux{n + 1}": 44 + n for n in range(3)},
    **{f"x{n + 1}": 47 + n for n in range(3)},
    **{f"qsave{n + 1}": 50 + n for n in range(5)},
    **{f"vor{n + 1}": 55 + n for n in range(3)},
    "ssf": 58, "lshock": 59, "crinod": 60,
}


def build(force: bool = False) -> str:
    """Compile oracle/*.cpp into libastr_oracle.so (g++, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("lineops.cpp", "miniapp.cpp", "solver.cpp", "astr_oracle.hpp", "upwind.hpp",
                                             "recons.hpp")]
    if not force and os.path.exists(_LIB):
        if all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs):
            return _LIB
    r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        # some images ship a g++ without libgomp: retry single-threaded
        r = subprocess.run(
            ["make", "-C", _HERE, "-B",
             "CXXFLAGS=-O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -Wno-unknown-pragmas"],
            capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.oracle_miniapp_create.restype = vp
        L.oracle_miniapp_create.argtypes = [ci]
        L.oracle_miniapp_destroy.argtypes = [vp]
        L.oracle_miniapp_run.argtypes = [vp, ci]
        L.oracle_miniapp_run.restype = ci
        L.oracle_miniapp_history.argtypes = [vp, vp]
        L.oracle_miniapp_get.argtypes = [vp, ci, vp]
        L.oracle_df_compact.argtypes = [ci, ci, vp, vp]
        L.oracle_diff6ec.argtypes = [ci, ci, vp, vp]
        L.oracle_compact_filter.argtypes = [ci, ci, cd, cd, vp, vp]
        L.oracle_scheme_tables.argtypes = [ci, ci, ci, cd, vp, vp, vp, vp, vp, vp]
        L.oracle_scheme_tables.restype = ci
        L.oracle_num_threads.restype = ci
        L.oracle_set_num_threads.restype = ci
        L.oracle_set_num_threads.argtypes = [ci]
        L.oracle_case_create.restype = vp
        L.oracle_case_create.argtypes = [ci] * 9 + [cd] * 8
        L.oracle_case_destroy.argtypes = [vp]
        L.oracle_case_nblocks.argtypes = [vp]
        L.oracle_case_nblocks.restype = ci
        L.oracle_case_block_info.argtypes = [vp, ci, vp]
        L.oracle_case_set_x.argtypes = [vp, ci, vp]
        for name in ("gridgeom", "tgvini", "filterq", "qswap", "gradcal", "zero_qrhs", "convrsdcal6",
                     "rhscal", "save_q", "updatefvar"):
            getattr(L, "oracle_case_" + name).argtypes = [vp]
        L.oracle_case_get.argtypes = [vp, ci, ci, vp]
        L.oracle_case_set.argtypes = [vp, ci, ci, vp]
        L.oracle_case_rk_update.argtypes = [vp, ci]
        L.oracle_case_rk_stage.argtypes = [vp, ci]
        L.oracle_case_set_flags.argtypes = [vp, ci, ci]
        L.oracle_case_set_bc.argtypes = [vp, vp, vp]
        L.oracle_case_set_flow.argtypes = [vp, ci, vp]
        L.oracle_case_set_inflow.argtypes = [vp, ci, vp, vp, vp]
        L.oracle_case_set_sponge.argtypes = [vp, ci, ci, ci, ci, vp]
        L.oracle_case_spongefilter.argtypes = [vp]
        L.oracle_case_updateq.argtypes = [vp]
        for name in ("oracle_case_crashcheck", "oracle_case_crashfix", "oracle_case_crinod_expansion"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = ctypes.c_longlong
        L.oracle_case_databakup.argtypes = [vp, ci]
        L.oracle_case_databakup.restype = ci
        L.oracle_case_nstep.argtypes = [vp]
        L.oracle_case_nstep.restype = ci
        L.oracle_case_set_sponge_circle.argtypes = [vp, cd, cd, cd, cd, cd]
        L.oracle_case_sponge_circle_coef.argtypes = [vp, ci, vp]
        L.oracle_case_sponge_circle_coef.restype = ci
        L.oracle_case_set_dimensional.argtypes = [vp, cd, cd, cd, cd]
        L.oracle_case_thermo.argtypes = [vp, vp]
        L.oracle_case_pinf.argtypes = [vp]
        L.oracle_case_pinf.restype = cd
        L.oracle_case_set_scheme.argtypes = [vp, ci]
        L.oracle_case_set_rkscheme.argtypes = [vp, ci]
        L.oracle_case_set_rkscheme.restype = ci
        L.oracle_case_set_upwind.argtypes = [vp, ci, ci, cd, cd]
        L.oracle_case_set_upwind_explicit.argtypes = [vp, ci, ci, cd, cd]
        L.oracle_case_convrsduwd.argtypes = [vp]
        L.oracle_case_convrsduwd.restype = ci
        L.oracle_recons_exp.argtypes = [vp, ci, ci, ci, ci, ci, cd]
        L.oracle_recons_exp.restype = cd
        L.oracle_case_ducrossensor.argtypes = [vp]
        L.oracle_case_convrsdcmp.argtypes = [vp]
        L.oracle_case_convrsdcmp.restype = ci
        L.oracle_flux_compact.argtypes = [ci, ci, ci, cd, vp, vp]
        L.oracle_mp5.argtypes = [vp, cd, ci]
        L.oracle_mp5.restype = cd
        L.oracle_steger_warming.argtypes = [cd, cd, vp, vp, vp]
        L.oracle_chardecomp.argtypes = [cd, vp, vp, vp, vp]
        L.oracle_chardecomp.restype = ci
        L.oracle_case_boucon.argtypes = [vp]
        L.oracle_case_boucon.restype = ci
        L.oracle_case_run.argtypes = [vp, ci]
        L.oracle_case_run.restype = ci
        L.oracle_case_history.argtypes = [vp, vp]
        L.oracle_case_reduce.argtypes = [vp, ci, vp]
        _lib = L
    return _lib


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> int:
    """OpenMP threads of the following runs (overrides OMP_NUM_THREADS); returns the count in effect."""
    return lib().oracle_set_num_threads(int(n))


# --------------------------------------------------------------------------------------
# single-pencil operators
# --------------------------------------------------------------------------------------
def df_compact(f: np.ndarray, ntype: int) -> np.ndarray:
    """6th-order compact derivative of one pencil f(-hm:dim+hm) -> df(0:dim)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    dim = f.size - 1 - 2 * HM
    out = np.empty(dim + 1)
    lib().oracle_df_compact(ntype, dim, f.ctypes.data, out.ctypes.data)
    return out


def diff6ec(f: np.ndarray, ntype: int) -> np.ndarray:
    f = np.ascontiguousarray(f, dtype=np.float64)
    dim = f.size - 1 - 2 * HM
    out = np.empty(dim + 1)
    lib().oracle_diff6ec(ntype, dim, f.ctypes.data, out.ctypes.data)
    return out


def compact_filter(f: np.ndarray, ntype: int, alfa: float = 0.49, beter_bound: float = 0.98) -> np.ndarray:
    f = np.ascontiguousarray(f, dtype=np.float64)
    dim = f.size - 1 - 2 * HM
    out = np.empty(dim + 1)
    lib().oracle_compact_filter(ntype, dim, alfa, beter_bound, f.ctypes.data, out.ctypes.data)
    return out


def flux_compact(f: np.ndarray, ntype: int, plus: bool, bfacmpld: float = 0.3) -> np.ndarray:
    """Compact 5th-order upwind interface flux of one pencil f(-hm:dim+hm) -> fh(-1:dim)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    dim = f.size - 1 - 2 * HM
    out = np.empty(dim + 2)
    lib().oracle_flux_compact(ntype, dim, int(plus), bfacmpld, f.ctypes.data, out.ctypes.data)
    return out


def recons_exp(f8, inode: int, dim: int, ntype: int, reschem: int, shock: bool = True, bfacmpld: float = 0.3) -> float:
    """recons_exp(f(1:8), ...) of src/flux.F90:269-350 (interface between f(4) and f(5))."""
    f = np.ascontiguousarray(f8, dtype=np.float64)
    assert f.size == 8
    return lib().oracle_recons_exp(f.ctypes.data, inode, dim, ntype, reschem, int(shock), bfacmpld)


def mp5(u5, ul: float, discont: bool = True) -> float:
    u = np.ascontiguousarray(u5, dtype=np.float64)
    return lib().oracle_mp5(u.ctypes.data, float(ul), int(discont))


def steger_warming(rho, vel, prs, tmp, q, dxi, jacob, gamma=1.4, mach=0.1):
    a = np.ascontiguousarray([rho, *vel, prs, tmp, *q, *dxi, jacob], dtype=np.float64)
    fp, fm = np.zeros(5), np.zeros(5)
    lib().oracle_steger_warming(gamma, mach, a.ctypes.data, fp.ctypes.data, fm.ctypes.data)
    return fp, fm


def chardecomp(left, right, gamma=1.4):
    """left/right = (ro, p, E, vel(3), ddi(3)) -> (REV, LEV) 5x5 (src/solver.F90:1958)."""
    l = np.ascontiguousarray(left, dtype=np.float64); r = np.ascontiguousarray(right, dtype=np.float64)
    rev, lev = np.zeros((5, 5)), np.zeros((5, 5))
    if lib().oracle_chardecomp(gamma, l.ctypes.data, r.ctypes.data, rev.ctypes.data, lev.ctypes.data):
        raise ValueError("chardecomp: degenerate metric normal")
    return rev, lev


def scheme_tables(is_filter: bool, ntype: int, dim: int, alfa: float = 0.49):
    """(first_node, a, c, ac1, ac2, ac3) of the pre-factored tridiagonal operator."""
    first = ctypes.c_int(0)
    n = lib().oracle_scheme_tables(int(is_filter), ntype, dim, alfa, ctypes.byref(first), None, None, None, None, None)
    arrs = [np.zeros(n) for _ in range(5)]
    lib().oracle_scheme_tables(int(is_filter), ntype, dim, alfa, ctypes.byref(first), *[a.ctypes.data for a in arrs])
    return (first.value, *arrs)


# --------------------------------------------------------------------------------------
# mini-app mode (pinned against tests/golden/state.ref_128)
# --------------------------------------------------------------------------------------
class MiniApp:
    IDS = {**{f"q{n + 1}": n for n in range(5)}, "rho": 5, "u": 6, "v": 7, "w": 8, "prs": 9, "tmp": 10,
           **{f"qrhs{n + 1}": 11 + n for n in range(5)}}

    def __init__(self, n: int):
        self.n = n
        self._h = lib().oracle_miniapp_create(n)

    def run(self, nsteps: int) -> np.ndarray:
        rows = lib().oracle_miniapp_run(self._h, nsteps)
        out = np.zeros((rows, 4))
        lib().oracle_miniapp_history(self._h, out.ctypes.data)
        return out

    def get(self, name: str) -> np.ndarray:
        m = self.n + 1 + 2 * HM
        out = np.empty((m, m, m), order="F")
        lib().oracle_miniapp_get(self._h, self.IDS[name], out.ctypes.data)
        return out

    def close(self):
        if self._h:
            lib().oracle_miniapp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------
# main-solver mode on a virtual block grid
# --------------------------------------------------------------------------------------
class Case:
    def __init__(self, ia, ja, ka, blocks=(1, 1, 1), homo=(True, True, True), lengths=None, reynolds=1600.0,
                 mach=0.1, alfa_filter=0.49, deltat=1e-3, sutherland_s=110.3):
        two_pi = 2.0 * np.pi
        lengths = lengths or (two_pi, two_pi, two_pi)
        self.dims = (ia, ja, ka)
        self.blocks = tuple(blocks)
        self._h = lib().oracle_case_create(ia, ja, ka, *blocks, *[int(h) for h in homo], *lengths, reynolds, mach,
                                           alfa_filter, deltat, sutherland_s)
        self.nblocks = lib().oracle_case_nblocks(self._h)

    def block_info(self, ib: int) -> dict:
        info = (ctypes.c_int * 21)()
        lib().oracle_case_block_info(self._h, ib, info)
        v = list(info)
        return dict(im=v[0], jm=v[1], km=v[2], npdc=v[3:6], is_ie=v[6:12], g0=v[12:15], nb=v[15:21])

    def shape(self, ib: int):
        b = self.block_info(ib)
        return (b["im"] + 1 + 2 * HM, b["jm"] + 1 + 2 * HM, b["km"] + 1 + 2 * HM)

    def get(self, name: str, ib: int = 0) -> np.ndarray:
        out = np.empty(self.shape(ib), order="F")
        lib().oracle_case_get(self._h, ib, FIELD_IDS[name], out.ctypes.data)
        return out

    def set(self, name: str, arr: np.ndarray, ib: int = 0):
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.shape(ib)
        lib().oracle_case_set(self._h, ib, FIELD_IDS[name], a.ctypes.data)

    def set_x(self, x: np.ndarray, ib: int = 0):
        a = np.asfortranarray(x, dtype=np.float64)
        lib().oracle_case_set_x(self._h, ib, a.ctypes.data)

    def set_flags(self, lfilter=True, diffterm=True):
        lib().oracle_case_set_flags(self._h, int(lfilter), int(diffterm))

    def set_bc(self, bctype, twall=(0.0,) * 6):
        """bctype(1:6) / twall(1:6) of the input file (imin,imax,jmin,jmax,kmin,kmax)."""
        bt = (ctypes.c_int * 6)(*[int(b) for b in bctype])
        tw = (ctypes.c_double * 6)(*[float(t) for t in twall])
        lib().oracle_case_set_bc(self._h, bt, tw)

    def set_flow(self, flowtype: int, force=(0.0, 0.0, 0.0)):
        """flowtype 0 generic, 1 channel (src_chan with body force `force`)."""
        f = (ctypes.c_double * 3)(*[float(v) for v in force])
        lib().oracle_case_set_flow(self._h, int(flowtype), f)

    def set_upwind(self, conschm: int = 543, lchardecomp: bool = True, bfacmpld: float = 0.3, shkcrt: float = 0.01):
        """conschm='543c' (convrsdcmp) with the input file's `recon_schem, lchardecomp, bfacmpld, shkcrt`."""
        lib().oracle_case_set_upwind(self._h, int(conschm), int(lchardecomp), float(bfacmpld), float(shkcrt))

    def set_upwind_explicit(self, recon_schem: int = 3, lchardecomp: bool = True, bfacmpld: float = 0.3,
                            shkcrt: float = 0.01):
        """conschm='<odd>..e' (convrsduwd, src/solver.F90:548) with `recon_schem` of recons_exp."""
        lib().oracle_case_set_upwind_explicit(self._h, int(recon_schem), int(lchardecomp), float(bfacmpld), float(shkcrt))

    def convrsduwd(self) -> int:
        return lib().oracle_case_convrsduwd(self._h)

    def ducrossensor(self):
        lib().oracle_case_ducrossensor(self._h)

    def convrsdcmp(self) -> int:
        return lib().oracle_case_convrsdcmp(self._h)

    def set_inflow(self, vel_in, tmp_in, tmp_prof, ib: int = 0):
        """inflow(1) data (src/bc.F90:69-83): vel_in(0:jm,0:km,3), tmp_in(0:jm,0:km), tmp_prof(0:jm)."""
        a = [np.asfortranarray(v, dtype=np.float64) for v in (vel_in, tmp_in, tmp_prof)]
        lib().oracle_case_set_inflow(self._h, ib, *[v.ctypes.data for v in a])

    def set_sponge(self, face: int, beg: int, end: int, coef=None, ib: int = 0):
        """Sponge layer of face 0 i0 / 1 im / 3 jm / 4 k0 / 5 km (src/sponge_layer.F90:67-319)."""
        a = np.asfortranarray(coef, dtype=np.float64) if coef is not None else np.zeros(1)
        lib().oracle_case_set_sponge(self._h, ib, face, beg, end, a.ctypes.data)

    def spongefilter(self):
        lib().oracle_case_spongefilter(self._h)

    def updateq(self):
        """q from density, velocity and temperature (src/fludyna.F90:254-300), as readcheckpoint ends."""
        lib().oracle_case_updateq(self._h)

    # crash control (src/mainloop.F90:709-1198)
    def crashcheck(self) -> int:
        """Flags the nodes whose density is not >= 0 as critical nodes; returns how many."""
        return int(lib().oracle_case_crashcheck(self._h))

    def crashfix(self) -> int:
        """Wipes nodes with rho / prs / tmp under 1e-5 (mean of the admissible neighbours); returns how many."""
        return int(lib().oracle_case_crashfix(self._h))

    def crinod_expansion(self) -> int:
        return int(lib().oracle_case_crinod_expansion(self._h))

    def databakup(self, mode: str):
        """'backup' / 'recovery' (two alternating in-memory copies of q, src/mainloop.F90:826-972)."""
        if lib().oracle_case_databakup(self._h, {"backup": 0, "recovery": 1}[mode]) != 0:
            raise RuntimeError("no backup data available")

    @property
    def nstep(self) -> int:
        return int(lib().oracle_case_nstep(self._h))

    def set_sponge_circle(self, centre=(0.0, 0.0, 0.0), range_spange=0.06, dampfac=0.05):
        """spg_def='circl' (src/sponge_layer.F90:369-440; the reference hard-codes these defaults)."""
        lib().oracle_case_set_sponge_circle(self._h, *[float(v) for v in centre], float(range_spange), float(dampfac))

    def sponge_circle_coef(self, ib: int = 0):
        """sponge_damp_coef(is:ie,js:je,ks:ke) of a block, or None when the block has no damped node (lsponge_loc)."""
        b = self.block_info(ib)
        se = b["is_ie"]
        shape = tuple(se[2 * d + 1] - se[2 * d] + 1 for d in range(3))
        out = np.zeros(shape, order="F")
        return out if lib().oracle_case_sponge_circle_coef(self._h, ib, out.ctypes.data) else None

    def set_dimensional(self, ref_tem: float, ref_vel: float, ref_len: float, ref_den: float):
        """nondimen=f: SI units, rgas=287.1 (src/solver.F90:124-148)."""
        lib().oracle_case_set_dimensional(self._h, ref_tem, ref_vel, ref_len, ref_den)

    def thermo(self) -> dict:
        out = (ctypes.c_double * 14)()
        lib().oracle_case_thermo(self._h, out)
        keys = ["reynolds", "mach", "const1", "const2", "const3", "const4", "const5", "const6", "const7", "rgas", "cp",
                "cv", "pinf", "nondimen"]
        return dict(zip(keys, list(out)))

    @property
    def pinf(self) -> float:
        return lib().oracle_case_pinf(self._h)

    def set_scheme(self, explicit: bool):
        lib().oracle_case_set_scheme(self._h, int(explicit))

    def set_rkscheme(self, scheme: int):
        """3: 'rk3' (default), 4: 'rk4' (src/mainloop.F90:348-388)."""
        if lib().oracle_case_set_rkscheme(self._h, int(scheme)) != 0:
            raise ValueError("rkscheme must be 3 or 4")

    def boucon(self):
        if lib().oracle_case_boucon(self._h) != 0:
            raise NotImplementedError("oracle boucon: bctype not restated")

    def reduce(self, what: int) -> np.ndarray:
        """Raw block-summed diagnostics: 0 (KE, enstrophy, dissipation sums), 1 (CFL maxima), 2 (channel mass flux,
        wall friction sums) -- see oracle_case_reduce."""
        out = np.zeros(3)
        lib().oracle_case_reduce(self._h, int(what), out.ctypes.data)
        return out

    def history(self) -> np.ndarray:
        rows = lib().oracle_case_run(self._h, 0)
        out = np.zeros((rows, 4))
        if rows:
            lib().oracle_case_history(self._h, out.ctypes.data)
        return out

    def run(self, nsteps: int) -> np.ndarray:
        lib().oracle_case_run(self._h, nsteps)
        return self.history()

    def rk_update(self, rkstep: int):
        lib().oracle_case_rk_update(self._h, rkstep)

    def rk_stage(self, rkstep: int):
        lib().oracle_case_rk_stage(self._h, rkstep)

    def __getattr__(self, name):
        if name in ("gridgeom", "tgvini", "filterq", "qswap", "gradcal", "zero_qrhs", "convrsdcal6", "rhscal",
                    "save_q", "updatefvar"):
            fn = getattr(lib(), "oracle_case_" + name)
            return lambda: fn(self._h)
        raise AttributeError(name)

    def close(self):
        if getattr(self, "_h", None):
            lib().oracle_case_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
