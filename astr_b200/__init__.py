"""astr_b200 -- B200-native (sm_100a) right-hand-side / Runge-Kutta stage engine for ASTR.

The product is ``libastr_gpu.so`` (C ABI in ``include/astr_gpu.h``, CUDA in ``csrc/``).
This package is the host-side mirror of the reference's stage-operator interface
(``filterq``, ``qswap``, ``gradcal``, ``rhscal``, RK update, ``updatefvar`` --
src/mainloop.F90:396-482) over that ABI, plus the block decomposition of
src/parallel.F90.  There is no CPU path: every operator raises if the CUDA library or a
GPU is missing.
"""
from .lib import AstrCfg, AstrGpuError, build, load, lib_path, FIELD_IDS, HM  # noqa: F401
from .parallel import Block, HaloMessage, decompose, halo_plan, mpisizedis  # noqa: F401
from .solver import RhsEngine, refcal, refcal_dimensional  # noqa: F401
from . import cases  # noqa: F401

__all__ = ["AstrCfg", "AstrGpuError", "build", "load", "lib_path", "FIELD_IDS", "HM", "Block", "decompose",
           "mpisizedis", "RhsEngine", "refcal", "refcal_dimensional", "cases"]
