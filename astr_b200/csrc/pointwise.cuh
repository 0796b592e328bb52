// astr_b200/csrc/pointwise.cuh -- declarations of the pointwise / surface kernels.
#pragma once
#include "common.cuh"

// Internal pool slots (all share the Layout of common.cuh).
enum Slot {
  S_Q = 0,        // 5  conserved variables            src/commarray.F90:63
  S_RHO = 5,      // 1
  S_VEL = 6,      // 3
  S_PRS = 9,      // 1
  S_TMP = 10,     // 1
  S_QRHS = 11,    // 5
  S_JAC = 16,     // 1  jacob
  S_DXI = 17,     // 9  dxi(a,b): slot 17+3a+b, a = xi index, b = x index
  S_RAW = 26,     // 12 raw xi-derivatives: slot 26+4d+n, n=0..2 velocity, 3 temperature
  S_SIGMA = 38,   // 6
  S_QFLUX = 44,   // 3
  S_G = 47,       // 15 combined flux G_d,n = Fv - Fc : slot 47+5d+n
  S_QSAVE = 62,   // 5
  S_CORE = 67,    // number of always-resident slots
  S_SCR = 67,     // 15 lazily allocated scratch (dvel 9, dtmp 3, vor 3 | gridgeom temporaries)
  S_TOTAL = 82
};

struct Thermo {
  double reynolds, prandtl, const1, const2, const5, const6, tempconst, tempconst1;
  double gamma, mach;
  int nondimen;             // 0: SI units, rgas = 287.1 (src/solver.F90:124-148)
  double rgas, cp, cv;
#ifdef __CUDACC__
  // thermal_scar (src/fludyna.F90:45-88)
  __device__ __forceinline__ double T_of(double p, double rho) const { return nondimen ? p / rho * const2 : p / rho / rgas; }
  __device__ __forceinline__ double rho_of(double p, double t) const { return nondimen ? p / t * const2 : p / t / rgas; }
  __device__ __forceinline__ double cotem() const { return nondimen ? const1 : cv; }           // fvar2q, fludyna.F90:344-348
  __device__ __forceinline__ double sos(double t) const { return nondimen ? sqrt(t) / mach : sqrt(gamma * rgas * t); }  // :832-859
  // miucal (src/fludyna.F90:791-817) as diffrsdcal6 uses it (src/solver.F90:2456-2460)
  __device__ __forceinline__ double miu(double t) const {
    if (nondimen) return (t * sqrt(t) * tempconst1 / (t + tempconst)) / reynolds;
    const double tn = t / 273.15;
    return 1.716e-5 * tn * sqrt(tn) * (273.15 + 110.4) / (t + 110.4);
  }
  __device__ __forceinline__ double hcc(double m) const { return nondimen ? (m / prandtl) / const5 : cp * m / prandtl; }  // :2519-2523
#endif
};

struct Box { int lo[3], hi[3]; };   // inclusive node ranges

enum { XMODE_SWAP = 0, XMODE_QSWAP = 1, XMODE_SYNC = 2 };

struct FieldList { double* f[ASTR_MAXF]; int nf; };

int pw_halo_wrap(const Layout& L, const FieldList& fl, int dir, int mode, cudaStream_t st);
int pw_q2fvar(const Layout& L, double* pool, const Thermo& th, const Box& b, cudaStream_t st);
int pw_visc(const Layout& L, double* pool, const Thermo& th, const Box& b, cudaStream_t st);
int pw_materialise_grad(const Layout& L, double* pool, double* dvel9_dtmp3_vor3, cudaStream_t st);
// dmask bit d => build G_d ; cmask[d][2]: conv part only where the other two indices lie in
// [s,e] ranges (src/solver.F90:2197-2198) ; diffterm adds the viscous part.
struct FluxRanges { int s[3], e[3]; };
// store_shell: write sigma/qflux on the face shells (0: a separate shell pass already did)
int pw_visc_flux(const Layout& L, double* pool, const Thermo& th, const FluxRanges& fr, int ndims, int store_shell,
                 cudaStream_t st);
int pw_flux(const Layout& L, double* pool, const Box& b, int dmask, const FluxRanges& fr, int diffterm,
            cudaStream_t st);
// scheme 3: q = (c1 qsave + c2 q jac + c3 dt qrhs) / jac (src/mainloop.F90:441-450)
// scheme 4: stages 1-3 q = (qsave + c1 dt qrhs) / jac, rhsav += c2 qrhs; stage 4 (last) q = (qsave + c1 dt (qrhs + rhsav)) / jac
//           (src/mainloop.F90:452-476); rhsav: 5 fields of the common Layout
struct RkCoef { double c1, c2, c3, dt; int first; int with_fvar; int rhs_in_g; int scheme; int last; double* rhsav; };
// src: device (force(1:3), force.ubulk) of src_chan still to be added, or nullptr
int pw_sum_qrhs(const Layout& L, double* pool, const double* src, cudaStream_t st);
int pw_bulk(const Layout& L, const double* pool, const double* yc, double* partial, double* out4, cudaStream_t st);
int pw_src_coef(const double* bulk4, const double force[3], double* src4, cudaStream_t st);
// explicit_central (diff6ec, src/derivative.F90:350-413); a.op.n / a.op.ntype describe the line
int pw_diff6e(int dir, const SweepArgs& a, cudaStream_t st);
// inflow (11, imin) / outflow (21, imax or jmax) / farfield (51, jmax) on one face (k_bcface)
struct BcArgs {
  int kind, side;
  double pinf, deltat;
  double vinf[3], roinf;   // free stream of the far-field faces (commvar uinf, vinf, winf, roinf)
  const double* vel_in;    // device (0:jm,0:km,3)   inflow only
  const double* tmp_in;    // device (0:jm,0:km)
  const double* tmp_prof;  // device (0:jm)
};
int pw_repitch(const Layout& L, const double* stage, double* field, cudaStream_t st);
// crash control (src/mainloop.F90:709-1198); crinod: one field of the common Layout holding 0/1
int pw_crashcheck(const Layout& L, const double* pool, double* crinod, unsigned long long* count, cudaStream_t st);
int pw_crinod_dilate(const Layout& L, double* crinod, double* tmp, unsigned long long* count, cudaStream_t st);
int pw_crashfix_flag(const Layout& L, const double* pool, double* crinod, double eps_rho, double eps_prs, double eps_tmp,
                     unsigned long long* count, long long* list, long long cap, cudaStream_t st);
struct CrashFixArgs { int g0[3], ia, ja; long long n; };
int pw_crashfix_apply(const Layout& L, double* pool, const Thermo& th, const long long* list, const CrashFixArgs& a,
                      unsigned long long* fixed, cudaStream_t st);
int pw_dense(const Layout& L, double* field, double* dense, bool pack, cudaStream_t st);
int pw_updateq(const Layout& L, double* pool, const Thermo& th, cudaStream_t st);
int pw_copy_box(const Layout& L, double* dst, const double* src, int nf, const Box& b, cudaStream_t st);
int pw_sponge(const Layout& L, double* pool, const Box& b, const double* coef, cudaStream_t st);
int pw_bcface(const Layout& L, double* pool, const Thermo& th, int dir, const BcArgs& a, cudaStream_t st);
int pw_noslip(const Layout& L, double* pool, const Thermo& th, int dir, int side, double tw, cudaStream_t st);
int pw_rk_update(const Layout& L, double* pool, const Thermo& th, const RkCoef& rk, const double* src,
                 cudaStream_t st);
int pw_add_force(const Layout& L, double* pool, const double* src /*device (force, force.ubulk)*/, cudaStream_t st);
// block-level diagnostics: what = 0 KE / enstrophy / dissipation sums, 1 CFL maxima, 2 channel mass flux / wall friction
int pw_reduce(const Layout& L, double* pool, const double* yc, const Thermo& th, int what, int wall_lo, int wall_hi,
              double* partial, double* out3, cudaStream_t st);
// face pack / unpack for the multi-block exchange (src/parallel.F90:4180-4218)
int pw_pack(const Layout& L, const FieldList& fl, int dir, int side, int l0, int l1, double* buf,
            cudaStream_t st);
int pw_unpack(const Layout& L, const FieldList& fl, int dir, int side, int l0, int l1, const double* buf,
              cudaStream_t st);
// fused pack + peer-to-peer exchange (k_xface in pointwise.cu); one XSide per face of the direction
struct XSide {
  int active;
  double* remote;                    // SEND: the neighbour's receive window (peer memory)
  const double* local;               // RECV: this rank's receive window
  const unsigned long long* wait_flag;   // in this rank's memory, written by the neighbour
  unsigned long long wait_val;
  unsigned long long* signal_flag;   // in the neighbour's memory
  unsigned long long signal_val;
  unsigned int* counter;             // CTA completion counter of this side (this rank's memory)
};
struct XArgs {
  XSide s[2]; int l0, l1;
  unsigned int* err;                 // sticky error flag of this rank (XHeader::err)
  long long timeout_cycles;          // flag waits longer than this poison the exchange; <= 0: wait forever
};
int pw_xsend(const Layout& L, const FieldList& fl, int dir, const XArgs& a, cudaStream_t st);
int pw_xrecv(const Layout& L, const FieldList& fl, int dir, const XArgs& a, cudaStream_t st);
// Fortran (im+11)(jm+11)(km+11) box <-> padded device box is done with cudaMemcpy3D in api.cu

// ---- upwind-biased compact convection (upwind.cu), conschm = '543c' -------------------------
// fields of the lazily allocated upwind pool (same Layout as every other field)
enum UpSlot {
  UP_FSW = 0,     // 10 Steger-Warming split fluxes F+(5), F-(5) at nodes
  UP_FHC = 10,    // 10 compact interface fluxes fhcp(5), fhcm(5); interface i is stored at node i
  UP_FH = 20,     // 5  limited interface flux Fh
  UP_SSF = 25,    // 1  Ducros sensor
  UP_LSH = 26,    // 1  lshock as 0/1
  UP_TOTAL = 27
};
struct UpwindArgs {
  Box box;              // interfaces treated: is-1..ie along the direction, s..e in the two others
  int s[3], e[3];       // is..ke
  int lss, lee;         // node range of the split fluxes along the direction (solver.F90:1313-1327)
  int dim, ntype;       // im/jm/km and npdc of the direction
  int lchardecomp, sson;
  int explicit_recons;  // 1: convrsduwd (explicit reconstruction recons_exp), 0: convrsdcmp (compact flux)
  int recon_schem;      // recons_exp scheme: -1, 0 (linear), 1 WENO, 2 WENO-Z, 3 MP, 5 MP-LD, 6 ROUND
  double bfacmpld;
  const double* crinod; // critical nodes of the crash control as 0/1 (nullptr: none), src/solver.F90:1456-1481
};
int uw_sw_split(const Layout& L, const double* pool, double* up, const Thermo& th, int dir, int lss, int lee,
                cudaStream_t st);
int uw_interface_flux(const Layout& L, const double* pool, double* up, const Thermo& th, int dir, const UpwindArgs& a,
                      cudaStream_t st);
int uw_fhdiff(const Layout& L, double* pool, const double* up, int dir, const UpwindArgs& a, int dst0, int rmw_mask,
              cudaStream_t st);
int uw_ducros_ssf(const Layout& L, const double* pool, double* up, const int npdc[3], cudaStream_t st);
int uw_ducros_flag(const Layout& L, double* up, const int npdc[3], double shkcrt, cudaStream_t st);
