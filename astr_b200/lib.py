"""ctypes binding of include/astr_gpu.h (the same symbols the Fortran bind(C) module binds)."""
from __future__ import annotations

import ctypes
import os
import subprocess

HM = 5
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libastr_gpu.so")

# field ids of astr_gpu_get_field / set_field (include/astr_gpu.h)
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

# every symbol include/astr_gpu.h declares (tests check the built library exports all)
SYMBOLS = [
    "astr_gpu_init", "astr_gpu_sizeof_cfg", "astr_gpu_finalize", "astr_gpu_last_error", "astr_gpu_synchronize",
    "astr_gpu_comm_unique_id", "astr_gpu_comm_init", "astr_gpu_set_metrics", "astr_gpu_gridgeom",
    "astr_gpu_upload_state", "astr_gpu_download_state", "astr_gpu_get_field", "astr_gpu_set_field",
    "astr_gpu_device_ptr", "astr_gpu_filterq", "astr_gpu_boucon", "astr_gpu_set_inflow", "astr_gpu_qswap", "astr_gpu_gradcal", "astr_gpu_rhscal",
    "astr_gpu_rk_update", "astr_gpu_spongefilter", "astr_gpu_updatefvar", "astr_gpu_rk_stage", "astr_gpu_rk_steps", "astr_gpu_rk_steps_timed",
    "astr_gpu_dataswap", "astr_gpu_set_force", "astr_gpu_set_sponge", "astr_gpu_set_sponge_global", "astr_gpu_crashcheck", "astr_gpu_databakup", "astr_gpu_crinod_expansion", "astr_gpu_crashfix", "astr_gpu_stage_checkpoint", "astr_gpu_restore_checkpoint", "astr_gpu_set_grid", "astr_gpu_reduce_tgv", "astr_gpu_reduce_cfl", "astr_gpu_reduce_channel", "astr_gpu_kernel_launches",
    "astr_gpu_set_profile", "astr_gpu_get_profile", "astr_gpu_bench_sweep",
]

PROFILE_CATEGORIES = ["filter_i", "filter_j", "filter_k", "halo", "grad_i", "grad_j", "grad_k", "visc", "flux",
                      "div_i", "div_j", "div_k", "rk", "fvar", "xpack", "xnccl", "xunpack"]


class AstrGpuError(RuntimeError):
    pass


class AstrCfg(ctypes.Structure):
    """struct astr_cfg of include/astr_gpu.h (field order must match)."""
    _fields_ = [
        ("abi_version", ctypes.c_int), ("device", ctypes.c_int),
        ("im", ctypes.c_int), ("jm", ctypes.c_int), ("km", ctypes.c_int),
        ("ia", ctypes.c_int), ("ja", ctypes.c_int), ("ka", ctypes.c_int),
        ("hm", ctypes.c_int), ("numq", ctypes.c_int), ("ndims", ctypes.c_int),
        ("npdc", ctypes.c_int * 3),
        ("is_", ctypes.c_int), ("ie", ctypes.c_int), ("js", ctypes.c_int), ("je", ctypes.c_int),
        ("ks", ctypes.c_int), ("ke", ctypes.c_int),
        ("lhomo", ctypes.c_int * 3), ("rank", ctypes.c_int * 3), ("size", ctypes.c_int * 3),
        ("nbr", ctypes.c_int * 6), ("my_rank", ctypes.c_int),
        ("conschm", ctypes.c_int), ("difschm", ctypes.c_int), ("scheme_compact", ctypes.c_int),
        ("rkscheme", ctypes.c_int), ("lfilter", ctypes.c_int), ("diffterm", ctypes.c_int),
        ("nondimen", ctypes.c_int), ("flowtype", ctypes.c_int),
        ("recon_schem", ctypes.c_int), ("conschm_explicit", ctypes.c_int), ("lchardecomp", ctypes.c_int), ("bctype", ctypes.c_int * 6),
        ("legacy_sweep", ctypes.c_int), ("overlap_visc", ctypes.c_int), ("xchg_nccl", ctypes.c_int),
        ("xchg_timeout_ms", ctypes.c_int), ("reserved0", ctypes.c_int),
        ("alfa_filter", ctypes.c_double),
        ("reynolds", ctypes.c_double), ("mach", ctypes.c_double), ("prandtl", ctypes.c_double),
        ("gamma", ctypes.c_double), ("ref_tem", ctypes.c_double),
        ("const1", ctypes.c_double), ("const2", ctypes.c_double), ("const3", ctypes.c_double),
        ("const4", ctypes.c_double), ("const5", ctypes.c_double), ("const6", ctypes.c_double),
        ("const7", ctypes.c_double),
        ("tempconst", ctypes.c_double), ("tempconst1", ctypes.c_double),
        ("deltat", ctypes.c_double), ("twall", ctypes.c_double * 6),
        ("pinf", ctypes.c_double), ("bfacmpld", ctypes.c_double), ("shkcrt", ctypes.c_double),
        ("uinf", ctypes.c_double), ("vinf", ctypes.c_double), ("winf", ctypes.c_double), ("roinf", ctypes.c_double),
    ]


def lib_path() -> str:
    return _SO


def use_debug_library(path: str):
    """Tests only: bind another build of the same sources (libastr_gpu_nofma.so, `make nofma`) instead of the
    production library.  Must be called before the first load()."""
    global _SO
    if _lib is not None:
        raise AstrGpuError("use_debug_library must be called before the library is loaded")
    _SO = path


def build_nofma() -> str:
    """Compile the -fmad=false debug library astr_b200/libastr_gpu_nofma.so (tests/test_gpu_nofma.py)."""
    csrc = os.path.join(_HERE, "csrc")
    r = subprocess.run(["make", "-C", csrc, "-j8", "nofma"], capture_output=True, text=True)
    if r.returncode != 0:
        raise AstrGpuError("building libastr_gpu_nofma.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return os.path.join(_HERE, "libastr_gpu_nofma.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into astr_b200/libastr_gpu.so (nvcc; no GPU needed)."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stdout[-4000:])
    if r.returncode != 0:
        raise AstrGpuError("building libastr_gpu.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return _SO


_lib = None


def load() -> ctypes.CDLL:
    """dlopen libastr_gpu.so and declare prototypes.  Fails loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise AstrGpuError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = ctypes.CDLL(_SO)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.astr_gpu_last_error.restype = ctypes.c_char_p
    L.astr_gpu_init.argtypes = [ctypes.POINTER(AstrCfg)]
    L.astr_gpu_comm_unique_id.argtypes = [vp]
    L.astr_gpu_comm_init.argtypes = [vp, ci, ci]
    L.astr_gpu_set_metrics.argtypes = [vp, vp]
    L.astr_gpu_gridgeom.argtypes = [vp]
    L.astr_gpu_upload_state.argtypes = [vp] * 5
    L.astr_gpu_download_state.argtypes = [vp] * 5
    L.astr_gpu_get_field.argtypes = [ci, vp]
    L.astr_gpu_set_field.argtypes = [ci, vp]
    L.astr_gpu_device_ptr.argtypes = [ci, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_longlong),
                                      ctypes.POINTER(ctypes.c_longlong)]
    L.astr_gpu_rk_update.argtypes = [ci, cd]
    L.astr_gpu_rk_stage.argtypes = [ci, cd]
    L.astr_gpu_rk_steps.argtypes = [ci, cd]
    L.astr_gpu_rk_steps_timed.argtypes = [ci, cd, ctypes.POINTER(ctypes.c_float)]
    L.astr_gpu_dataswap.argtypes = [ci, ci]
    L.astr_gpu_set_force.argtypes = [vp]
    L.astr_gpu_set_grid.argtypes = [vp]
    L.astr_gpu_set_sponge.argtypes = [ci, ci, ci, vp]
    L.astr_gpu_set_sponge_global.argtypes = [vp]
    L.astr_gpu_crashcheck.argtypes = [vp]
    L.astr_gpu_databakup.argtypes = [ci, vp, vp]
    L.astr_gpu_crinod_expansion.argtypes = [vp]
    L.astr_gpu_crashfix.argtypes = [ci, ci, vp]
    L.astr_gpu_stage_checkpoint.argtypes = [vp] * 6
    L.astr_gpu_restore_checkpoint.argtypes = [vp] * 6
    L.astr_gpu_set_inflow.argtypes = [vp, vp, vp]
    L.astr_gpu_reduce_tgv.argtypes = [vp]
    L.astr_gpu_reduce_cfl.argtypes = [vp]
    L.astr_gpu_reduce_channel.argtypes = [vp]
    L.astr_gpu_kernel_launches.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    L.astr_gpu_set_profile.argtypes = [ci]
    L.astr_gpu_get_profile.argtypes = [vp, vp, ci]
    L.astr_gpu_bench_sweep.argtypes = [ci, ci, ci, ci, ctypes.POINTER(ctypes.c_float)]
    if L.astr_gpu_sizeof_cfg() != ctypes.sizeof(AstrCfg):
        raise AstrGpuError("AstrCfg mirror out of sync with include/astr_gpu.h")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise AstrGpuError(load().astr_gpu_last_error().decode())
