"""Host-side mirror of the reference's stage-operator interface over the C ABI.

Method names, argument meaning and call order are those of the reference's time loop
(src/mainloop.F90:301-482 ``time_integration_rk``): ``filterq`` (src/comsolver.F90:514),
``qswap`` (src/parallel.F90:4848), ``gradcal`` (src/comsolver.F90:244), ``rhscal``
(src/solver.F90:185), the RK update (src/mainloop.F90:427-476) and ``updatefvar``
(src/fludyna.F90:191).  Every method is a thin call into libastr_gpu.so; nothing is
computed in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np

from . import lib as _l
from .parallel import Block

HM = _l.HM


def refcal(reynolds: float, mach: float, ref_tem: float = 273.15, gamma: float = 1.4, prandtl: float = 0.72,
           sutherland_s: float = 110.3) -> Dict[str, float]:
    """Reference constants of src/solver.F90:104-126 (nondimen branch)."""
    m2 = mach * mach
    tempconst = sutherland_s / ref_tem
    return dict(
        reynolds=reynolds, mach=mach, prandtl=prandtl, gamma=gamma, ref_tem=ref_tem,
        const1=1.0 / (gamma * (gamma - 1.0) * m2),
        const2=gamma * m2,
        const3=(gamma - 1.0) / 3.0 * prandtl * m2,
        const4=(gamma - 1.0) * m2 * reynolds * prandtl,
        const5=(gamma - 1.0) * m2,
        const6=1.0 / (gamma - 1.0),
        const7=(gamma - 1.0) * m2 * reynolds * prandtl,
        tempconst=tempconst, tempconst1=1.0 + tempconst,
    )


def refcal_dimensional(ref_tem: float, ref_vel: float, ref_len: float, ref_den: float, gamma: float = 1.4,
                       prandtl: float = 0.72) -> Dict[str, float]:
    """nondimen=f branch of refcal (src/solver.F90:124-148): SI units, rgas=287.1; Mach and Reynolds follow
    from the reference state with sos (src/fludyna.F90:853) and the dimensional miucal (:809-811)."""
    rgas = 287.1
    cp = gamma / (gamma - 1.0) * rgas
    cv = rgas / (gamma - 1.0)
    pinf = ref_den * ref_tem * rgas
    tn = ref_tem / 273.15
    ref_miu = 1.716e-5 * tn * np.sqrt(tn) * (273.15 + 110.4) / (ref_tem + 110.4)
    mach = ref_vel / np.sqrt(gamma * rgas * ref_tem)
    reynolds = ref_den * ref_vel * ref_len / ref_miu
    th = refcal(float(reynolds), float(mach), ref_tem=ref_tem, gamma=gamma, prandtl=prandtl)
    th.update(nondimen=0, rgas=rgas, cp=cp, cv=cv, pinf=float(pinf))
    # free stream of the far-field faces: uinf=ref_vel, vinf=winf=0, roinf=ref_den (src/solver.F90:131-139)
    th.update(uinf=float(ref_vel), vinf=0.0, winf=0.0, roinf=float(ref_den))
    return th


class RhsEngine:
    """One block (= one MPI rank of the reference = one GPU) of the RHS / RK-stage engine."""

    def __init__(self, block: Block, global_dims: Sequence[int], homo: Sequence[bool], thermo: Dict[str, float],
                 deltat: float = 1e-3, alfa_filter: float = 0.49, lfilter: bool = True, diffterm: bool = True,
                 device: int = -1, flowtype: int = 0, bctype: Sequence[int] = (1,) * 6,
                 twall: Sequence[float] = (0.0,) * 6, explicit: bool = False, conschm: Optional[int] = None,
                 lchardecomp: bool = False, bfacmpld: float = 0.3, shkcrt: float = 0.01, recon_schem: int = 3,
                 conschm_explicit: bool = False, legacy_sweep: bool = False, overlap_visc: bool = False,
                 xchg_nccl: bool = False, xchg_timeout_ms: int = 0,
                 freestream: Optional[Sequence[float]] = None, rkscheme: int = 3):
        self.block = block
        self.global_dims = tuple(global_dims)
        self.deltat = deltat
        self.thermo = thermo
        self._lib = _l.load()
        c = _l.AstrCfg()
        c.abi_version = 3
        # engine switches of ABI v3 (0 = default: register line-solve engine, peer-memory halo exchange)
        c.legacy_sweep, c.overlap_visc, c.xchg_nccl = int(legacy_sweep), int(overlap_visc), int(xchg_nccl)
        c.xchg_timeout_ms = int(xchg_timeout_ms)
        c.device = device
        c.im, c.jm, c.km = block.dims
        c.ia, c.ja, c.ka = global_dims
        c.hm, c.numq, c.ndims = HM, 5, (2 if block.dims[2] == 0 else 3)   # ka==0: 2-D block
        c.npdc[:] = block.npdc
        c.is_, c.js, c.ks = block.s
        c.ie, c.je, c.ke = block.e
        c.lhomo[:] = [int(h) for h in homo]
        c.rank[:] = block.rk
        c.size[:] = block.size
        c.nbr[:] = block.nbr
        c.my_rank = block.rank
        # conschm/difschm '643c' (compact_central) or '642e' (explicit_central), comsolver.F90:76-84
        c.conschm, c.difschm, c.scheme_compact, c.rkscheme = (642, 642, 0, 3) if explicit else (643, 643, 1, 3)
        c.rkscheme = int(rkscheme)   # 3: 'rk3', 4: 'rk4' (src/mainloop.F90:348-388)
        if conschm is not None:      # e.g. 543: upwind compact convection (convrsdcmp) over difschm
            c.conschm = int(conschm)
        c.recon_schem, c.lchardecomp, c.bfacmpld, c.shkcrt = int(recon_schem), int(lchardecomp), bfacmpld, shkcrt
        c.conschm_explicit = int(conschm_explicit)      # conschm '<odd>..e': convrsduwd + recons_exp
        c.lfilter, c.diffterm, c.nondimen, c.flowtype = int(lfilter), int(diffterm), int(thermo.get("nondimen", 1)), flowtype
        c.bctype[:] = [int(b) for b in bctype]
        c.twall[:] = [float(t) for t in twall]
        c.alfa_filter = alfa_filter
        # roinf*tinf/const2 with roinf=tinf=1 (src/solver.F90:113-120), or thermal(tinf,roinf) for nondimen=f
        c.pinf = thermo.get("pinf", 1.0 / thermo["const2"])
        # free stream of the far-field faces (commvar uinf, vinf, winf, roinf): 1, 0, 0, 1 for the nondimensional
        # gas (src/solver.F90:113-120); refcal_dimensional brings ref_vel, 0, 0, ref_den (:131-139); an explicit
        # `freestream` wins over both
        c.uinf, c.vinf, c.winf, c.roinf = 1.0, 0.0, 0.0, 1.0
        for k, v in thermo.items():
            if k not in ("nondimen", "rgas", "cp", "cv", "pinf"):
                setattr(c, k, v)
        if freestream is not None:
            c.uinf, c.vinf, c.winf, c.roinf = freestream
        c.deltat = deltat
        self.cfg = c
        _l.check(self._lib.astr_gpu_init(ctypes.byref(c)))
        self._open = True

    # ---- shapes -------------------------------------------------------------------------
    @property
    def shape(self):
        im, jm, km = self.block.dims
        return (im + 1 + 2 * HM, jm + 1 + 2 * HM, km + 1 + 2 * HM)

    def empty(self, ncomp: Optional[int] = None) -> np.ndarray:
        shp = self.shape + ((ncomp,) if ncomp else ())
        return np.zeros(shp, order="F")

    @staticmethod
    def _ptr(a: Optional[np.ndarray]):
        if a is None:
            return None
        assert a.flags.f_contiguous and a.dtype == np.float64
        return a.ctypes.data

    # ---- communicator (replaces mpiinitial, src/parallel.F90:152) ---------------------------
    def comm_init(self, nranks: int, rank: int, bcast):
        """bcast(bytes_or_None) -> bytes broadcasts rank 0's 128-byte NCCL id (MPI_Bcast /
        torch.distributed on the host side)."""
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _l.check(self._lib.astr_gpu_comm_unique_id(buf))
        uid = bcast(bytes(buf.raw) if rank == 0 else None)
        buf2 = ctypes.create_string_buffer(uid, 128)
        _l.check(self._lib.astr_gpu_comm_init(buf2, nranks, rank))

    # ---- setup ----------------------------------------------------------------------------
    def set_metrics(self, dxi: np.ndarray, jacob: np.ndarray):
        """dxi(-hm:im+hm,-hm:jm+hm,-hm:km+hm,3,3), jacob(...) as left by geomcal (src/geom.F90:43)."""
        assert dxi.shape == self.shape + (3, 3) and jacob.shape == self.shape
        _l.check(self._lib.astr_gpu_set_metrics(self._ptr(np.asfortranarray(dxi)), self._ptr(np.asfortranarray(jacob))))

    def gridgeom(self, x: np.ndarray):
        """Device-side gridgeom (src/geom.F90:99) from x(-hm:im+hm,...,3)."""
        assert x.shape == self.shape + (3,)
        _l.check(self._lib.astr_gpu_gridgeom(self._ptr(np.asfortranarray(x))))

    def set_grid(self, x: np.ndarray):
        """Node coordinates x(-hm:im+hm,...,3) for src_chan's y integration (channel only)."""
        assert x.shape == self.shape + (3,)
        _l.check(self._lib.astr_gpu_set_grid(self._ptr(np.asfortranarray(x))))

    def set_inflow(self, vel_in: np.ndarray, tmp_in: np.ndarray, tmp_prof: np.ndarray):
        """Inflow data of bctype(1)=11 (src/bc.F90:69-83): vel_in(0:jm,0:km,3), tmp_in(0:jm,0:km), tmp_prof(0:jm)."""
        im, jm, km = self.block.dims
        a = [np.asfortranarray(v, dtype=np.float64) for v in (vel_in, tmp_in, tmp_prof)]
        assert a[0].shape == (jm + 1, km + 1, 3) and a[1].shape == (jm + 1, km + 1) and a[2].shape == (jm + 1,)
        _l.check(self._lib.astr_gpu_set_inflow(*[v.ctypes.data for v in a]))

    def set_force(self, force: Sequence[float]):
        f = (ctypes.c_double * 3)(*[float(v) for v in force])
        _l.check(self._lib.astr_gpu_set_force(f))

    def upload_state(self, q=None, rho=None, vel=None, prs=None, tmp=None):
        _l.check(self._lib.astr_gpu_upload_state(*[self._ptr(a) for a in (q, rho, vel, prs, tmp)]))

    def download_state(self, q=None, rho=None, vel=None, prs=None, tmp=None):
        _l.check(self._lib.astr_gpu_download_state(*[self._ptr(a) for a in (q, rho, vel, prs, tmp)]))

    def get(self, name: str) -> np.ndarray:
        out = self.empty()
        _l.check(self._lib.astr_gpu_get_field(_l.FIELD_IDS[name], out.ctypes.data))
        return out

    def set(self, name: str, arr: np.ndarray):
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.shape
        _l.check(self._lib.astr_gpu_set_field(_l.FIELD_IDS[name], a.ctypes.data))

    # ---- stage operators --------------------------------------------------------------------
    def filterq(self):
        _l.check(self._lib.astr_gpu_filterq())

    def boucon(self):
        _l.check(self._lib.astr_gpu_boucon())

    def qswap(self):
        _l.check(self._lib.astr_gpu_qswap())

    def gradcal(self):
        _l.check(self._lib.astr_gpu_gradcal())

    def rhscal(self):
        _l.check(self._lib.astr_gpu_rhscal())

    def rk_update(self, rkstep: int, deltat: Optional[float] = None):
        _l.check(self._lib.astr_gpu_rk_update(rkstep, self.deltat if deltat is None else deltat))

    def set_sponge(self, face: int, beg: int, end: int, coef: Optional[np.ndarray] = None):
        """Sponge layer of face 0 i0 / 1 im / 3 jm / 4 k0 / 5 km (src/sponge_layer.F90); coef over the layer box."""
        a = np.asfortranarray(coef, dtype=np.float64) if coef is not None else None
        _l.check(self._lib.astr_gpu_set_sponge(face, beg, end, a.ctypes.data if a is not None else None))

    # ---- crash control (src/mainloop.F90:709-1198) ------------------------------------------
    def crashcheck(self) -> int:
        """Flags the nodes whose density is not >= 0 as critical nodes (crinod); returns how many on this rank."""
        n = ctypes.c_longlong(0)
        _l.check(self._lib.astr_gpu_crashcheck(ctypes.byref(n)))
        return int(n.value)

    def crashfix(self) -> int:
        """Wipes nodes with rho / prs / tmp under 1e-5 (mean of the admissible neighbours); returns how many."""
        n = ctypes.c_longlong(0)
        g0 = getattr(self.block, "g0", (0, 0, 0))
        _l.check(self._lib.astr_gpu_crashfix(int(g0[0]), int(g0[1]), ctypes.byref(n)))
        return int(n.value)

    def crinod_expansion(self) -> int:
        n = ctypes.c_longlong(0)
        _l.check(self._lib.astr_gpu_crinod_expansion(ctypes.byref(n)))
        return int(n.value)

    def databakup(self, mode: str):
        """'backup' / 'recovery' over two alternating device copies of q; returns (copy used, its recover_counter)."""
        slot, cnt = ctypes.c_int(0), ctypes.c_int(0)
        _l.check(self._lib.astr_gpu_databakup({"backup": 0, "recovery": 1}[mode], ctypes.byref(slot), ctypes.byref(cnt)))
        return int(slot.value), int(cnt.value)

    # ---- checkpoint staging (src/readwrite.F90:1723-1984, :1381-1470) -------------------------
    def stage_checkpoint(self):
        """ro, u1, u2, u3, p, t as dense node arrays (0:im,0:jm,0:km), the datasets writeflfed writes."""
        dims = tuple(d + 1 for d in self.block.dims)
        out = [np.empty(dims, order="F") for _ in range(6)]
        _l.check(self._lib.astr_gpu_stage_checkpoint(*[a.ctypes.data for a in out]))
        return dict(zip(("ro", "u1", "u2", "u3", "p", "t"), out))

    def restore_checkpoint(self, data):
        """readcheckpoint + updateq: primitives from the six datasets, q from density, velocity and temperature."""
        dims = tuple(d + 1 for d in self.block.dims)
        arrs = [np.asfortranarray(data[k], dtype=np.float64) for k in ("ro", "u1", "u2", "u3", "p", "t")]
        assert all(a.shape == dims for a in arrs)
        _l.check(self._lib.astr_gpu_restore_checkpoint(*[a.ctypes.data for a in arrs]))

    def set_sponge_global(self, coef: Optional[np.ndarray]):
        """spg_def='circl': sponge_damp_coef(is:ie,js:je,ks:ke) (src/sponge_layer.F90:369-440), None if this rank has no damped node."""
        a = None if coef is None else np.asfortranarray(coef, dtype=np.float64)
        _l.check(self._lib.astr_gpu_set_sponge_global(a.ctypes.data if a is not None else None))

    def spongefilter(self):
        _l.check(self._lib.astr_gpu_spongefilter())

    def updatefvar(self):
        _l.check(self._lib.astr_gpu_updatefvar())

    def rk_stage(self, rkstep: int, deltat: Optional[float] = None):
        _l.check(self._lib.astr_gpu_rk_stage(rkstep, self.deltat if deltat is None else deltat))

    def steploop(self, nsteps: int, deltat: Optional[float] = None):
        """nsteps x RK3 (src/mainloop.F90:103-205 without I/O hooks)."""
        _l.check(self._lib.astr_gpu_rk_steps(nsteps, self.deltat if deltat is None else deltat))

    def steploop_timed(self, nsteps: int, deltat: Optional[float] = None) -> float:
        """steploop bracketed by CUDA events on the library stream; returns milliseconds."""
        ms = ctypes.c_float(0)
        _l.check(self._lib.astr_gpu_rk_steps_timed(nsteps, self.deltat if deltat is None else deltat,
                                                   ctypes.byref(ms)))
        return ms.value

    def dataswap(self, name: str, direction: int = 0):
        _l.check(self._lib.astr_gpu_dataswap(_l.FIELD_IDS[name], direction))

    def synchronize(self):
        _l.check(self._lib.astr_gpu_synchronize())

    # ---- statistics (src/statistic.F90:871-990) -----------------------------------------------
    def reduce_tgv(self):
        """Block sums of rho|u|^2 and rho|omega|^2 (kenergycal / enstophycal before psum)."""
        return self.reduce_tgv3()[:2]

    def reduce_tgv3(self):
        """Block sums (rho|u|^2, rho|omega|^2, 2 miu (S:S - div^2/3)): kenergycal, enstophycal, diss_rate_cal."""
        out = (ctypes.c_double * 3)()
        _l.check(self._lib.astr_gpu_reduce_tgv(out))
        return out[0], out[1], out[2]

    def reduce_cfl(self):
        """Block maxima (deltai, deltaj, deltak) of cflcal (src/commcal.F90:27-74)."""
        out = (ctypes.c_double * 3)()
        _l.check(self._lib.astr_gpu_reduce_cfl(out))
        return out[0], out[1], out[2]

    def cfl(self, pmax=lambda v: v) -> float:
        di, dj, dk = self.reduce_cfl()
        return self.deltat * (pmax(di) + pmax(dj) + pmax(dk))

    def reduce_channel(self):
        """Block sums of massfluxchan / fbcxchan (src/statistic.F90:1437, :1303) before psum and /norm."""
        out = (ctypes.c_double * 2)()
        _l.check(self._lib.astr_gpu_reduce_channel(out))
        return out[0], out[1]

    def channel_stats(self, psum=lambda v: v):
        """(massflux, fbcx) normalised as massfluxchan / fbcxchan: /ia (2-D) or /(ia*ka)."""
        mf, fb = self.reduce_channel()
        ia, ja, ka = self.global_dims
        norm = float(ia) if ka == 0 else float(ia * ka)
        return psum(mf) / norm, psum(fb) / norm

    @staticmethod
    def chanfoce(force: float, massflux: float, friction: float, massflux_target: float, nstep: int, deltat: float,
                 ly: float = 2.0, uinf: float = 1.0, ref_len: float = 1.0) -> float:
        """Body force that holds the target mass flux (chanfoce, src/statistic.F90:1494-1537): a scalar formula of the two
        device reductions; stays on the host."""
        if nstep == 0:
            return friction / ly
        gn = ly * force - friction
        qn1 = massflux + deltat * gn
        return force - (2.0 * (qn1 - massflux_target) - 0.2 * (massflux - massflux_target)) / ly * (uinf / ref_len)

    def tgv_stats(self, psum=lambda v: v, xmax: float = 2.0 * np.pi):
        """(kenergy, enstrophy) normalised as kenergycal / enstophycal; psum sums over ranks."""
        ke, en = self.reduce_tgv()
        ke, en = psum(ke), psum(en)
        ia, ja, ka = self.global_dims
        cnt = float(ia * ja * ka)
        l_0 = xmax / (2.0 * np.pi)
        return 0.5 * ke / cnt, 0.5 * en / cnt / ((1.0 / l_0) ** 2)

    # ---- introspection ----------------------------------------------------------------------
    def kernel_launches(self) -> int:
        n = ctypes.c_longlong(0)
        self._lib.astr_gpu_kernel_launches(ctypes.byref(n))
        return n.value

    def set_profile(self, on: bool = True):
        _l.check(self._lib.astr_gpu_set_profile(int(on)))

    def get_profile(self):
        n = len(_l.PROFILE_CATEGORIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        self._lib.astr_gpu_get_profile(ms, cnt, n)
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_l.PROFILE_CATEGORIES)}

    def bench_sweep(self, op: int, direction: int, nfields: int = 5, iters: int = 10) -> float:
        ms = ctypes.c_float(0)
        _l.check(self._lib.astr_gpu_bench_sweep(op, direction, nfields, iters, ctypes.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "_open", False):
            self._lib.astr_gpu_finalize()
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
