/* =====================================================================================
 * astr_gpu.h -- C ABI of libastr_gpu.so: the B200 (sm_100a) right-hand-side / Runge-Kutta
 * stage engine that sits behind ASTR's Fortran driver.
 *
 * Every entry point replaces one stage-operator call of the reference's time loop
 * (src/mainloop.F90:396-482, `time_integration_rk`).  The Fortran side binds them with
 * ISO_C_BINDING (`fortran/astr_gpu_mod.F90`); the Python mirror in
 * astr_b200/ binds them with ctypes.  All functions return 0 on success and a non-zero
 * status otherwise; `astr_gpu_last_error()` returns the message.  The reference's own
 * error style is print+stop (src/parallel.F90:1278 `mpistop`): the shim calls mpistop
 * with that message.
 *
 * Array layout at the boundary is exactly the reference's Fortran layout
 * (src/commarray.F90:63-106): column major, i fastest, halo'd arrays
 * (-hm:im+hm,-hm:jm+hm,-hm:km+hm[,n]) with hm=5, the variable index slowest.
 * The library owns device mirrors and all scratch; host pointers are only read in
 * upload / set_* calls and only written in download / get_* / reduce_* calls, and are
 * never retained.
 *
 * Threading model: one host process (MPI rank) per GPU and per block; all calls of one
 * context come from one thread.  Calls are asynchronous on the library's CUDA streams;
 * download / get / reduce / synchronize / finalize wait for completion.
 * ===================================================================================== */
#ifndef ASTR_GPU_H
#define ASTR_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASTR_GPU_HM 5          /* src/commvar.F90:184  parameter(hm=5) */
#define ASTR_GPU_NUMQ 5        /* src/solver.F90:50    numq=5+num_species+num_modequ */
#define ASTR_GPU_ABI_VERSION 3

/* Everything `solvrinit` (src/comsolver.F90:47-146), `refcal` (src/solver.F90:28-173)
 * and `parallelini` (src/parallel.F90:919-1248) have decided by the time the time loop
 * starts.  Plain ints/doubles only, so that a Fortran `type, bind(C)` mirrors it. */
typedef struct astr_cfg {
  int abi_version;            /* ASTR_GPU_ABI_VERSION                                   */
  int device;                 /* CUDA device ordinal; <0 = keep the current device       */
  int im, jm, km;             /* local block: nodes 0..im etc. (src/commvar.F90)         */
  int ia, ja, ka;             /* global grid intervals                                   */
  int hm;                     /* must be 5                                               */
  int numq;                   /* must be 5                                               */
  int ndims;                  /* 3, or 2 with km=0 (src/parallel.F90:208-214)            */
  int npdc[3];                /* npdci,npdcj,npdck: 1,2,3,4 (src/parallel.F90:1042-1228) */
  int is, ie, js, je, ks, ke; /* qrhs accumulation ranges (src/parallel.F90:1079-1095)   */
  int lhomo[3];               /* lihomo,ljhomo,lkhomo                                    */
  int rank[3];                /* irk,jrk,krk                                             */
  int size[3];                /* isize,jsize,ksize                                       */
  int nbr[6];                 /* neighbour ranks i-,i+,j-,j+,k-,k+ ; -1 = MPI_PROC_NULL  */
  int my_rank;                /* mpirank                                                 */
  int conschm;                /* 643 / 642: central convrsdcal6 (src/solver.F90:2173);
                                 543: upwind compact convrsdcmp (src/solver.F90:1271)   */
  int difschm;                /* 643 (compact) or 642 (explicit)                         */
  int scheme_compact;         /* difschm(4:4): 1 = 'c' compact_central, 0 = 'e' explicit */
  int rkscheme;               /* 3 = rk3 (TVD), 4 = rk4: src/mainloop.F90:348-388        */
  int lfilter;                /* filterq enabled                                         */
  int diffterm;               /* viscous terms enabled                                   */
  int nondimen;               /* 1 nondimensional; 0 SI units with rgas=287.1, cp, cv as
                                 src/solver.F90:124-128 (mach, reynolds, const1..7, pinf
                                 as refcal leaves them)                                  */
  int flowtype;               /* 0 = tgv / generic (no source), 1 = channel (src_chan)   */
  int recon_schem;            /* input-file `recon_schem`: scheme of recons_exp for the
                                 explicit family (-1, 0, 1, 2, 3, 5, 6; flux.F90:269-350) */
  int conschm_explicit;       /* conschm(4:4)=='e' with an odd first digit: convrsduwd
                                 (src/solver.F90:548); 0: '543c' / central                */
  int lchardecomp;            /* characteristic decomposition + Ducros sensor on/off     */
  int bctype[6];              /* bctype(1:6) of the input file: imin,imax,jmin,jmax,kmin,
                                 kmax (src/bc.F90:327-407 boucon).  On the device:
                                 1 periodic/none; 41 isothermal wall (any face);
                                 11 inflow (imin); 21 outflow (imax, jmax); 51 farfield
                                 (jmin, jmax, kmin, kmax); 421 slip adiabatic wall
                                 (jmin, jmax)                                            */
  /* engine switches (ABI v3; they were environment variables in v2).  0 = default.      */
  int legacy_sweep;           /* 1: every line solve on the shared-memory engine
                                 (sweep.cu) instead of the register engine (sweep2.cu)  */
  int overlap_visc;           /* 1: multi-block only: the sigma/qflux exchange runs on a
                                 side stream behind the interior stress+flux pass       */
  int xchg_nccl;              /* 1: halo exchange through ncclSend/ncclRecv instead of
                                 peer-memory stores (CUDA IPC)                          */
  int xchg_timeout_ms;        /* peer-memory exchange: a neighbour flag that does not
                                 arrive within this time poisons the exchange (sticky
                                 error at the next synchronising call).  0: 60 s;
                                 < 0: wait forever, like ncclRecv / MPI_Sendrecv         */
  int reserved0;              /* keeps the doubles 8-byte aligned; must be 0            */
  double alfa_filter;         /* 0.49 in every example                                   */
  double reynolds, mach, prandtl, gamma, ref_tem;
  double const1, const2, const3, const4, const5, const6, const7; /* solver.F90:104-126  */
  double tempconst, tempconst1; /* Sutherland: 110.3/ref_tem (src/solver.F90:122)        */
  double deltat;
  double twall[6];            /* wall temperature of the bctype-41 faces (twall(1:6))    */
  double pinf;                /* free-stream pressure roinf*tinf/const2 (solver.F90:120) */
  double bfacmpld;            /* blending factor of the compact upwind scheme (flux.F90) */
  double shkcrt;              /* shock-sensor threshold of ducrossensor (commcal.F90)    */
  double uinf, vinf, winf, roinf; /* free stream of the far-field faces (commvar; nondimensional
                                 defaults 1, 0, 0, 1: src/solver.F90:113-120)            */
} astr_cfg;

/* ---- field ids for astr_gpu_get_field / set_field / device_ptr ----------------------
 * Same numbering as the test oracle: 0-4 q | 5 rho | 6-8 vel | 9 prs | 10 tmp |
 * 11-15 qrhs | 16 jacob | 17-25 dxi(a,b) a-major (a = xi index, b = x index) |
 * 26-34 dvel(m,n) m-major | 35-37 dtmp | 38-43 sigma | 44-46 qflux | 47-49 x |
 * 50-54 qsave | 55-57 vor | 58 ssf | 59 lshock (0/1; both only with conschm 543 and
 * lchardecomp).  dvel/dtmp/vor are materialised on demand. */
enum {
  ASTR_F_Q = 0, ASTR_F_RHO = 5, ASTR_F_VEL = 6, ASTR_F_PRS = 9, ASTR_F_TMP = 10,
  ASTR_F_QRHS = 11, ASTR_F_JACOB = 16, ASTR_F_DXI = 17, ASTR_F_DVEL = 26,
  ASTR_F_DTMP = 35, ASTR_F_SIGMA = 38, ASTR_F_QFLUX = 44, ASTR_F_X = 47,
  ASTR_F_QSAVE = 50, ASTR_F_VOR = 55, ASTR_F_SSF = 58, ASTR_F_LSHOCK = 59, ASTR_F_CRINOD = 60, ASTR_F_COUNT = 61
};

/* replaces solvrinit: builds the line operators (fd_scheme_initiate
 * src/derivative.F90:63-158, compact_filter_initiate src/filter.F90:31-100) and
 * allocates the device mirrors of src/commarray.F90:56-110. */
int astr_gpu_init(const astr_cfg* cfg);
int astr_gpu_sizeof_cfg(void);   /* sizeof(astr_cfg): lets a binding verify its mirror */
int astr_gpu_finalize(void);
const char* astr_gpu_last_error(void);
int astr_gpu_synchronize(void);

/* multi-GPU: the NCCL communicator that replaces MPI_COMM_WORLD for halo traffic
 * (src/parallel.F90:152 mpiinitial).  Rank 0 creates the id, the host side broadcasts
 * the 128 bytes (MPI_Bcast / torch.distributed), every rank calls comm_init. */
int astr_gpu_comm_unique_id(char id[128]);
int astr_gpu_comm_init(const char id[128], int nranks, int rank);

/* end of geomcal (src/geom.F90:43): dxi(-hm:im+hm,..,3,3), jacob(-hm:im+hm,..) */
int astr_gpu_set_metrics(const double* dxi, const double* jacob);
/* device-side gridgeom (src/geom.F90:99-700) from node coordinates x(-hm:im+hm,..,3)
 * (only nodes 0..im,0..jm,0..km are read). */
int astr_gpu_gridgeom(const double* x);

/* end of flowinit / after boucon: q(..,numq), rho, vel(..,3), prs, tmp -- any pointer may
 * be NULL (skipped). */
int astr_gpu_upload_state(const double* q, const double* rho, const double* vel,
                          const double* prs, const double* tmp);
int astr_gpu_download_state(double* q, double* rho, double* vel, double* prs, double* tmp);
/* one halo'd 3-D field in or out (tests, checkpointing, device-side BC staging) */
int astr_gpu_get_field(int field_id, double* host);
int astr_gpu_set_field(int field_id, const double* host);
int astr_gpu_device_ptr(int field_id, void** dptr, long long strides[3], long long* origin);

/* ---- stage operators, one per reference subroutine ---------------------------------- */
int astr_gpu_filterq(void);      /* src/comsolver.F90:514  filterq                      */
int astr_gpu_boucon(void);       /* src/bc.F90:327         boucon: noslip :6306, inflow
                                    :1366, outflow :3404, farfield :3008                */
/* inflow data of alloinflow (src/bc.F90:69-83): vel_in(0:jm,0:km,3), tmp_in(0:jm,0:km),
 * tmp_prof(0:jm); call again whenever inflowintp / inflowintx refresh them (rkstep 1) */
int astr_gpu_set_inflow(const double* vel_in, const double* tmp_in, const double* tmp_prof);
int astr_gpu_qswap(void);        /* src/parallel.F90:4848  qswap                        */
int astr_gpu_gradcal(void);      /* src/comsolver.F90:244  gradcal                      */
int astr_gpu_rhscal(void);       /* src/solver.F90:185     rhscal (zeroes qrhs first,
                                    i.e. includes src/mainloop.F90:408 `qrhs=0`)        */
int astr_gpu_rk_update(int rkstep, double deltat); /* src/mainloop.F90:427-476          */
int astr_gpu_spongefilter(void); /* src/sponge_layer.F90:55 spongefilter (layer form)    */
int astr_gpu_updatefvar(void);   /* src/fludyna.F90:191    updatefvar                   */
/* all of the above in the order of src/mainloop.F90:396-482 */
int astr_gpu_rk_stage(int rkstep, double deltat);
/* nsteps x (rk stages 1..3) with nothing in between */
int astr_gpu_rk_steps(int nsteps, double deltat);
/* same, bracketed by CUDA events on the library stream; returns the device time in ms */
int astr_gpu_rk_steps_timed(int nsteps, double deltat, float* ms);

/* generic halo exchange of one device field (dataswap, src/parallel.F90:3499-4382) */
int astr_gpu_dataswap(int field_id, int direction /*0 = all, 1..3*/);

/* sponge layers (spongelayer_define_ijk / layer_setup stay on the host, src/sponge_layer.F90:442-1011):
 * face 0 i0, 1 im, 3 jm, 4 k0, 5 km (spongefilter_layer has no j0 block); beg..end = node range of the layer
 * along the face direction on this rank, beg<0 = the layer exists (lspg_* set) but not on this rank;
 * coef = sponge_damp_coef over [beg:end] x the is:ie / js:je / ks:ke ranges of the other two directions */
int astr_gpu_set_sponge(int face, int beg, int end, const double* coef);
/* spg_def='circl' (src/sponge_layer.F90:321-440): coef = sponge_damp_coef(is:ie,js:je,ks:ke) of this rank as
 * spongelayer_define_circle leaves it (host, start-up), or NULL on a rank whose lsponge_loc is false; switches
 * astr_gpu_spongefilter to spongefilter_global (dataswap(q) in every direction + damped 7-point average) */
int astr_gpu_set_sponge_global(const double* coef);

/* ---- crash control (lcracon), src/mainloop.F90:709-1198.  crinod lives on the device as a 0/1 field
 * (ASTR_F_CRINOD, all zero until one of these calls sets a node); convrsdcmp / convrsduwd read it (hdiss,
 * src/solver.F90:1456-1481, :660-759).  Counts are per rank: `por` / `psum`, the messages, the decision to stop and
 * the scalars saved with a backup (nstep, time, massflux, force ...) stay on the host.
 * crashcheck: *nbad = nodes with q(:,1) not >= 0 (flagged).  databakup: mode 0 'backup' / 1 'recovery' over two
 * alternating device copies of q(0:im,0:jm,0:km,:); *slot = copy used (0 dat_a, 1 dat_b), *recover_counter = its
 * recover_counter afterwards; a recovery runs updatefvar and, from the second recovery of a copy on,
 * crinod_expansion.  crashfix: nodes with rho / prs / tmp under 1e-5 are flagged and replaced by the mean of their
 * admissible neighbours in storage order; ig0, jg0 = global index of node 0 (module parallel). */
int astr_gpu_crashcheck(long long* nbad);
int astr_gpu_databakup(int mode, int* slot, int* recover_counter);
int astr_gpu_crinod_expansion(long long* counter);
int astr_gpu_crashfix(int ig0, int jg0, long long* nfixed);

/* ---- checkpoint staging: the datasets ro, u1, u2, u3, p, t of writeflfed (src/readwrite.F90:1723-1984) and
 * readcheckpoint (:1381-1470) as dense node arrays (0:im,0:jm,0:km), Fortran order.  stage: device -> host (packing
 * on the device overlaps the copies); restore: host -> device, then updateq (q from density, velocity and
 * temperature, src/fludyna.F90:254-300).  The HDF5 calls stay in Fortran. */
int astr_gpu_stage_checkpoint(double* ro, double* u1, double* u2, double* u3, double* p, double* t);
int astr_gpu_restore_checkpoint(const double* ro, const double* u1, const double* u2, const double* u3,
                                const double* p, const double* t);

/* body force of src_chan (src/solver.F90:295-353), added to qrhs in rhscal when
 * flowtype = 1; `force` is what massfluxchan/chanfoce (src/statistic.F90:1437-1520) keep
 * updating on the Fortran side.  src_chan integrates in y, so the channel case needs the
 * node coordinates x(-hm:im+hm,..,3): astr_gpu_set_grid (or astr_gpu_gridgeom, which keeps
 * them as a side effect). */
int astr_gpu_set_force(const double force[3]);
int astr_gpu_set_grid(const double* x);

/* per-step diagnostics of rkfirst / steploop as BLOCK-level partial results: the caller applies psum / pmax over
 * the ranks and the reference's normalisation, exactly where the Fortran routines call psum / pmax themselves.
 * reduce_tgv: out[0] = sum rho |u|^2 (kenergycal, src/statistic.F90:938), out[1] = sum rho |omega|^2 (enstophycal,
 *   :871), out[2] = sum 2 miu (S:S - div^2/3) (diss_rate_cal, :994) over nodes 1..im,1..jm,1..km; ndims=3.
 * reduce_cfl: out = max of (Ubar, Ubar -+ c|grad xi|) per direction over nodes 0..im,0..jm,0..km (cflcal,
 *   src/commcal.F90:27-74); CFL = deltat * (pmax(out[0]) + pmax(out[1]) + pmax(out[2])).
 * reduce_channel: out[0] = sum of 0.5 (q2(j)+q2(j-1)) dy (massfluxchan, src/statistic.F90:1437), out[1] = wall
 *   friction sum +miu du/dy at j=0 (jrk=0), -miu du/dy at j=jm (jrk=jrkm) (fbcxchan, :1303); chanfoce (:1494) is a
 *   scalar formula of these two and stays on the host.
 * reduce_tgv / reduce_channel require gradcal of the current stage. */
int astr_gpu_reduce_tgv(double out[3]);
int astr_gpu_reduce_cfl(double out[3]);
int astr_gpu_reduce_channel(double out[2]);

/* introspection for bench / tests */
int astr_gpu_kernel_launches(long long* count);       /* launches since init            */
/* CUDA-event profile of the stage, per category (filter i/j/k, halo, grad i/j/k, visc,
 * flux, div i/j/k, rk, fvar, exchange pack / NCCL / unpack): accumulated ms and span counts
 * since set_profile(1). */
int astr_gpu_set_profile(int on);
int astr_gpu_get_profile(double* ms, long long* n, int cap);
int astr_gpu_bench_sweep(int op /*0 deriv,1 filter*/, int dir /*0,1,2*/, int nfields,
                         int iters, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* ASTR_GPU_H */
