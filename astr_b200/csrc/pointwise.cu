// astr_b200/csrc/pointwise.cu -- pointwise and surface kernels of the RK stage (sm_100a).
// All are one-thread-per-node, i fastest (coalesced 8-byte accesses on 128-byte aligned
// rows); they are pure streaming kernels bounded by HBM bandwidth.
#include "pointwise.cuh"

namespace {

constexpr int PW_T = 128;

// Thread -> node of a box.  Wide boxes: one CTA per 128 consecutive i of a (j,k) row.  Thin boxes (the
// i-direction halo slabs, <= 16 nodes wide): the CTA's threads run over (i fastest, then j), so that a
// 128-thread CTA covers 128/w rows instead of one row with w active threads.
constexpr int PW_THIN = 16;
__device__ __forceinline__ bool box_node(const Box& b, int& i, int& j, int& k) {
  const int w = b.hi[0] - b.lo[0] + 1;
  k = b.lo[2] + blockIdx.z;
  if (w <= PW_THIN) {
    const int t = blockIdx.x * PW_T + threadIdx.x;
    i = b.lo[0] + t % w;
    j = b.lo[1] + t / w;
    return j <= b.hi[1];
  }
  i = b.lo[0] + blockIdx.x * PW_T + threadIdx.x;
  j = b.lo[1] + blockIdx.y;
  return i <= b.hi[0];
}
inline dim3 box_grid(const Box& b) {
  const int w = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1, nk = b.hi[2] - b.lo[2] + 1;
  if (w <= PW_THIN) return dim3((w * nj + PW_T - 1) / PW_T, 1, nk);
  return dim3((w + PW_T - 1) / PW_T, nj, nk);
}
inline bool box_empty(const Box& b) {
  return b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2];
}

// ---------------------------------------------------------------------------------
// periodic wrap of a single-block direction: array{3,4,5}d_sendrecv with one block
// (src/parallel.F90:4169-4174), qswap (:4876-4884), array3d_sync
// ---------------------------------------------------------------------------------
template <int DIR>
__global__ void k_halo_wrap(const Layout L, const FieldList fl, const int mode) {
  // threads run over the two other indices (a fastest)
  const int na = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int nb = (DIR == 2) ? L.jm + 1 : L.km + 1;
  double* f = fl.f[blockIdx.z];
  if (DIR == 0) {
    // i direction: the halo nodes of one row are contiguous -- 8 lanes per row (lane l < 5
    // copies halo node l+1 of both sides), so a warp touches 4 rows x 40 bytes per access
    // instead of 32 rows x 8 bytes
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = t >> 3, l = (t & 7) + 1;
    const int b = blockIdx.y;
    if (a >= na || b >= nb) return;
    double* p = f + L.idx(0, a, b);
    const int dm = L.im;
    if (mode != XMODE_SYNC && l <= ASTR_HM) {
      p[-l] = p[dm - l];
      p[dm + l] = p[l];
    }
    if (mode != XMODE_SWAP && l == 8) {
      const double v = 0.5 * (p[0] + p[dm]);
      p[0] = v;
      p[dm] = v;
    }
    return;
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (a >= na || b >= nb) return;
  long long base, sl;
  int dm;
  if (DIR == 1) { base = L.idx(a, 0, b); sl = L.sj; dm = L.jm; }
  else { base = L.idx(a, b, 0); sl = L.sk; dm = L.km; }
  double* p = f + base;
  if (dm == 0) {
    // 2-D block (ka==0): every plane -hm..hm is a copy of plane 0 (src/parallel.F90:4296-4299, :5161-5164)
    if (mode != XMODE_SYNC) {
      const double v = p[0];
#pragma unroll
      for (int l = 1; l <= ASTR_HM; ++l) { p[-l * sl] = v; p[l * sl] = v; }
    }
    return;
  }
  if (mode != XMODE_SYNC) {
#pragma unroll
    for (int l = 1; l <= ASTR_HM; ++l) {
      p[-l * sl] = p[(dm - l) * sl];
      p[(dm + l) * sl] = p[l * sl];
    }
  }
  if (mode != XMODE_SWAP) {
    const double v = 0.5 * (p[0] + p[dm * sl]);
    p[0] = v;
    p[dm * sl] = v;
  }
}

// ---------------------------------------------------------------------------------
// q2fvar_3da + thermal_3d (src/fludyna.F90:545-634, :136-179), nondimensional, no species
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void q2fvar_node(double* __restrict__ pool, const long long fs, const long long x,
                                            const Thermo& th) {
  const double r = pool[(S_Q + 0) * fs + x];
  const double u = pool[(S_Q + 1) * fs + x] / r;
  const double v = pool[(S_Q + 2) * fs + x] / r;
  const double w = pool[(S_Q + 3) * fs + x] / r;
  const double p = (pool[(S_Q + 4) * fs + x] - 0.5 * r * (u * u + v * v + w * w)) / th.const6;
  pool[S_RHO * fs + x] = r;
  pool[(S_VEL + 0) * fs + x] = u;
  pool[(S_VEL + 1) * fs + x] = v;
  pool[(S_VEL + 2) * fs + x] = w;
  pool[S_PRS * fs + x] = p;
  pool[S_TMP * fs + x] = th.T_of(p, r);
}

__global__ void k_q2fvar(const Layout L, double* __restrict__ pool, const Thermo th, const Box b) {
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  q2fvar_node(pool, L.fstride, L.idx(i, j, k), th);
}

// ---------------------------------------------------------------------------------
// dvel / dtmp from the raw xi-derivatives: gradcal's accumulation
// dvel(m,n) = sum_d d(u_m)/d(xi_d) * dxi(d,n)   (src/comsolver.F90:280-484)
// ---------------------------------------------------------------------------------
struct Grad { double dv[3][3], dt[3]; };

__device__ __forceinline__ void load_grad(const double* __restrict__ pool, const long long fs,
                                          const long long x, Grad& g) {
  double dxi[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int n = 0; n < 3; ++n) dxi[d][n] = pool[(S_DXI + 3 * d + n) * fs + x];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const double r0 = pool[(S_RAW + 0 + m) * fs + x];
    const double r1 = pool[(S_RAW + 4 + m) * fs + x];
    const double r2 = pool[(S_RAW + 8 + m) * fs + x];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const double v = r0 * dxi[0][n] + r1 * dxi[1][n] + r2 * dxi[2][n];
      if (m < 3) g.dv[m][n] = v; else g.dt[n] = v;
    }
  }
}

__global__ void k_materialise_grad(const Layout L, const double* __restrict__ pool, double* __restrict__ out) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  Grad g;
  load_grad(pool, fs, x, g);
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int n = 0; n < 3; ++n) out[(3 * m + n) * fs + x] = g.dv[m][n];
#pragma unroll
  for (int n = 0; n < 3; ++n) out[(9 + n) * fs + x] = g.dt[n];
  out[12 * fs + x] = g.dv[2][1] - g.dv[1][2];   // src/solver.F90:2456-2458
  out[13 * fs + x] = g.dv[0][2] - g.dv[2][0];
  out[14 * fs + x] = g.dv[1][0] - g.dv[0][1];
}

// ---------------------------------------------------------------------------------
// viscous stress and heat flux, diffrsdcal6 pointwise part (src/solver.F90:2433-2602)
// with miucal (src/fludyna.F90:791-812)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void visc_node(const double* __restrict__ pool, const long long fs, const long long x,
                                          const Thermo& th, double (&sg)[6], double (&qf)[3]) {
  Grad g;
  load_grad(pool, fs, x, g);
  const double t = pool[S_TMP * fs + x];
  const double miu = th.miu(t);
  const double s11 = g.dv[0][0];
  const double s12 = 0.5 * (g.dv[0][1] + g.dv[1][0]);
  const double s13 = 0.5 * (g.dv[0][2] + g.dv[2][0]);
  const double s22 = g.dv[1][1];
  const double s23 = 0.5 * (g.dv[1][2] + g.dv[2][1]);
  const double s33 = g.dv[2][2];
  const double skk = (1.0 / 3.0) * (s11 + s22 + s33);
  const double miu2 = 2.0 * miu;
  const double hcc = th.hcc(miu);
  sg[0] = miu2 * (s11 - skk);
  sg[1] = miu2 * s12;
  sg[2] = miu2 * s13;
  sg[3] = miu2 * (s22 - skk);
  sg[4] = miu2 * s23;
  sg[5] = miu2 * (s33 - skk);
  const double u = pool[(S_VEL + 0) * fs + x], v = pool[(S_VEL + 1) * fs + x], w = pool[(S_VEL + 2) * fs + x];
  qf[0] = hcc * g.dt[0] + sg[0] * u + sg[1] * v + sg[2] * w;
  qf[1] = hcc * g.dt[1] + sg[1] * u + sg[3] * v + sg[4] * w;
  qf[2] = hcc * g.dt[2] + sg[2] * u + sg[4] * v + sg[5] * w;
}

__global__ void k_visc(const Layout L, double* __restrict__ pool, const Thermo th, const Box b) {
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  double sg[6], qf[3];
  visc_node(pool, fs, x, th, sg, qf);
#pragma unroll
  for (int n = 0; n < 6; ++n) pool[(S_SIGMA + n) * fs + x] = sg[n];
#pragma unroll
  for (int n = 0; n < 3; ++n) pool[(S_QFLUX + n) * fs + x] = qf[n];
}

// ---------------------------------------------------------------------------------
// combined flux G_d = Fv_d - Fc_d on a box.
//   Fc: convrsdcal6 (src/solver.F90:2200-2224), Fv: diffrsdcal6 (src/solver.F90:2623-2680)
// the derivative is linear, so  -d(Fc)/dxi + d(Fv)/dxi = d(G)/dxi : 5 line solves per
// direction instead of 9.
// ---------------------------------------------------------------------------------
template <int DMASK>
__device__ __forceinline__ void flux_node(double* __restrict__ pool, const long long fs, const long long x,
                                          const int (&ijk)[3], const FluxRanges& fr, const int diffterm,
                                          const double (&sg)[6], const double (&qf)[3]) {
  const double q0 = pool[(S_Q + 0) * fs + x], q1 = pool[(S_Q + 1) * fs + x], q2 = pool[(S_Q + 2) * fs + x],
               q3 = pool[(S_Q + 3) * fs + x], q4 = pool[(S_Q + 4) * fs + x];
  const double u = pool[(S_VEL + 0) * fs + x], v = pool[(S_VEL + 1) * fs + x], w = pool[(S_VEL + 2) * fs + x];
  const double p = pool[S_PRS * fs + x];
  const double jac = pool[S_JAC * fs + x];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (!(DMASK & (1 << d))) continue;
    const double d1 = pool[(S_DXI + 3 * d + 0) * fs + x];
    const double d2 = pool[(S_DXI + 3 * d + 1) * fs + x];
    const double d3 = pool[(S_DXI + 3 * d + 2) * fs + x];
    // convective part only on pencils inside the is:ie / js:je / ks:ke ranges of the two
    // other directions
    const int o1 = (d == 0) ? 1 : 0, o2 = (d == 2) ? 1 : 2;
    const bool conv = ijk[o1] >= fr.s[o1] && ijk[o1] <= fr.e[o1] && ijk[o2] >= fr.s[o2] && ijk[o2] <= fr.e[o2];
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0, g4 = 0.0;
    if (diffterm) {
      g1 = (sg[0] * d1 + sg[1] * d2 + sg[2] * d3) * jac;
      g2 = (sg[1] * d1 + sg[3] * d2 + sg[4] * d3) * jac;
      g3 = (sg[2] * d1 + sg[4] * d2 + sg[5] * d3) * jac;
      g4 = (qf[0] * d1 + qf[1] * d2 + qf[2] * d3) * jac;
    }
    if (conv) {
      const double uu = d1 * u + d2 * v + d3 * w;
      g0 = g0 - jac * q0 * uu;
      g1 = g1 - jac * (q1 * uu + d1 * p);
      g2 = g2 - jac * (q2 * uu + d2 * p);
      g3 = g3 - jac * (q3 * uu + d3 * p);
      g4 = g4 - jac * (q4 + p) * uu;
    }
    pool[(S_G + 5 * d + 0) * fs + x] = g0;
    pool[(S_G + 5 * d + 1) * fs + x] = g1;
    pool[(S_G + 5 * d + 2) * fs + x] = g2;
    pool[(S_G + 5 * d + 3) * fs + x] = g3;
    pool[(S_G + 5 * d + 4) * fs + x] = g4;
  }
}

template <int DMASK>
__global__ void k_flux(const Layout L, double* __restrict__ pool, const Box b, const FluxRanges fr,
                       const int diffterm) {
  int ijk[3];
  if (!box_node(b, ijk[0], ijk[1], ijk[2])) return;
  const long long fs = L.fstride, x = L.idx(ijk[0], ijk[1], ijk[2]);
  double sg[6] = {0, 0, 0, 0, 0, 0}, qf[3] = {0, 0, 0};
  if (diffterm) {
#pragma unroll
    for (int n = 0; n < 6; ++n) sg[n] = pool[(S_SIGMA + n) * fs + x];
#pragma unroll
    for (int n = 0; n < 3; ++n) qf[n] = pool[(S_QFLUX + n) * fs + x];
  }
  flux_node<DMASK>(pool, fs, x, ijk, fr, diffterm, sg, qf);
}

// Fused interior pass: sigma / qflux stay in registers between the viscous-stress evaluation and
// the flux assembly of all three directions.  They are written to memory only on the shells
// within ASTR_HM nodes of a block face, which is all the halo exchange (solver.F90:2604-2606)
// and the halo-slab flux passes read.
template <int DMASK>
__global__ void k_visc_flux(const Layout L, double* __restrict__ pool, const Thermo th, const FluxRanges fr,
                            const int store_shell) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int ijk[3];
  if (!box_node(b, ijk[0], ijk[1], ijk[2])) return;
  const long long fs = L.fstride, x = L.idx(ijk[0], ijk[1], ijk[2]);
  double sg[6], qf[3];
  visc_node(pool, fs, x, th, sg, qf);
  const bool shell = ijk[0] <= ASTR_HM || ijk[0] >= L.im - ASTR_HM || ijk[1] <= ASTR_HM || ijk[1] >= L.jm - ASTR_HM ||
                     ijk[2] <= ASTR_HM || ijk[2] >= L.km - ASTR_HM;
  if (shell && store_shell) {
#pragma unroll
    for (int n = 0; n < 6; ++n) pool[(S_SIGMA + n) * fs + x] = sg[n];
#pragma unroll
    for (int n = 0; n < 3; ++n) pool[(S_QFLUX + n) * fs + x] = qf[n];
  }
  flux_node<DMASK>(pool, fs, x, ijk, fr, 1, sg, qf);
}

// ---------------------------------------------------------------------------------
// RK3-TVD / RK4 update (src/mainloop.F90:427-476) fused with updatefvar (src/fludyna.F90:191)
// ---------------------------------------------------------------------------------
__global__ void k_rk_update(const Layout L, double* __restrict__ pool, const Thermo th, const RkCoef rk,
                            const double* __restrict__ src) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  const double jac = pool[S_JAC * fs + x];
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    const double q = pool[(S_Q + m) * fs + x];
    double qs;
    if (rk.first) {
      qs = q * jac;                               // mainloop.F90:429-433
      pool[(S_QSAVE + m) * fs + x] = qs;
    } else {
      qs = pool[(S_QSAVE + m) * fs + x];
    }
    // qrhs = sum over directions of d(G_d)/d(xi_d); after rhscal the G slots hold the derivatives
    double rhs;
    if (rk.rhs_in_g)
      rhs = pool[(S_G + m) * fs + x] + pool[(S_G + 5 + m) * fs + x] + pool[(S_G + 10 + m) * fs + x];
    else
      rhs = pool[(S_QRHS + m) * fs + x];
    // src_chan (src/solver.F90:341-352), still pending when qrhs lives in the G slots
    if (src != nullptr && m > 0) rhs = rhs + src[m - 1] * jac;
    double vv;
    if (rk.scheme == 4) {
      double* av = rk.rhsav + m * fs + x;
      if (!rk.last) {
        vv = qs + rk.c1 * rk.dt * rhs;
        *av = rk.first ? rk.c2 * rhs : *av + rk.c2 * rhs;       // rhsav = 0 at rkstep 1 (mainloop.F90:436)
      } else {
        vv = qs + rk.c1 * rk.dt * (rhs + *av);
      }
    } else {
      vv = rk.c1 * qs + rk.c2 * q * jac + rk.c3 * rhs * rk.dt;
    }
    pool[(S_Q + m) * fs + x] = vv / jac;
  }
  if (rk.with_fvar) q2fvar_node(pool, fs, x, th);
}

// qrhs(0:im,0:jm,0:km,:) as an array of its own (staged API / tests / source terms)
__global__ void k_sum_qrhs(const Layout L, double* __restrict__ pool, const double* __restrict__ src) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  const double jac = pool[S_JAC * fs + x];
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    double rhs = pool[(S_G + m) * fs + x] + pool[(S_G + 5 + m) * fs + x] + pool[(S_G + 10 + m) * fs + x];
    if (src != nullptr && m > 0) rhs = rhs + src[m - 1] * jac;
    pool[(S_QRHS + m) * fs + x] = rhs;
  }
}

// ---------------------------------------------------------------------------------
// src_chan bulk integrals (src/solver.F90:317-339): trapezoidal rule in y of q(1:4) over
// nodes 1..im,1..jm,1..km; one CTA per (j,k) row, fixed-order tree => deterministic.
// yc = x(:,:,:,2) in the common Layout.
// ---------------------------------------------------------------------------------
__global__ void k_bulk_rows(const Layout L, const double* __restrict__ pool, const double* __restrict__ yc,
                            double* __restrict__ partial, const int k1) {
  const int j = 1 + blockIdx.x, k = k1 + blockIdx.y;    // k1 = 0 for 2-D blocks (solver.F90:306-315)
  const long long fs = L.fstride;
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 1 + threadIdx.x; i <= L.im; i += PW_T) {
    const long long x = L.idx(i, j, k), xm = x - L.sj;
    const double dy = yc[x] - yc[xm];
#pragma unroll
    for (int m = 0; m < 4; ++m) a[m] += 0.5 * (pool[(S_Q + m) * fs + xm] + pool[(S_Q + m) * fs + x]) * dy;
  }
  __shared__ double s[4][PW_T];
#pragma unroll
  for (int m = 0; m < 4; ++m) s[m][threadIdx.x] = a[m];
  __syncthreads();
  for (int h = PW_T / 2; h > 0; h >>= 1) {
    if (threadIdx.x < h)
#pragma unroll
      for (int m = 0; m < 4; ++m) s[m][threadIdx.x] += s[m][threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x < 4) {
    const long long row = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    partial[4 * row + threadIdx.x] = s[threadIdx.x][0];
  }
}
__global__ void k_bulk_final(const double* __restrict__ partial, const long long nrows, double* __restrict__ out4) {
  __shared__ double s[4][256];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long r = threadIdx.x; r < nrows; r += 256)
#pragma unroll
    for (int m = 0; m < 4; ++m) a[m] += partial[4 * r + m];
#pragma unroll
  for (int m = 0; m < 4; ++m) s[m][threadIdx.x] = a[m];
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h)
#pragma unroll
      for (int m = 0; m < 4; ++m) s[m][threadIdx.x] += s[m][threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x < 4) out4[threadIdx.x] = s[threadIdx.x][0];
}
// bulk4 = psum'ed (robulk, rho*u1 .., rho*u3 integrals) -> src4 = (force, force.ubulk)  (:336-339,:349)
__global__ void k_src_coef(const double* __restrict__ bulk4, const double f0, const double f1, const double f2,
                           double* __restrict__ src4) {
  const double u1 = bulk4[1] / bulk4[0], u2 = bulk4[2] / bulk4[0], u3 = bulk4[3] / bulk4[0];
  src4[0] = f0; src4[1] = f1; src4[2] = f2;
  src4[3] = f0 * u1 + f1 * u2 + f2 * u3;
}

// ---------------------------------------------------------------------------------
// explicit_central -> diff6ec (src/derivative.F90:319-413): explicit 6th-order first
// derivative with the 2-4 boundary closure on physical-boundary ends (ntype 1/2/4).
// One thread per node, i fastest; neighbours along j/k come from L1/L2 (a plane of a
// 512^3 field is 2 MB).  Out of place only (in != out).
// ---------------------------------------------------------------------------------
template <int DIR>
__global__ void k_diff6e(const Layout L, const SweepArgs a) {
  const int i = blockIdx.x * PW_T + threadIdx.x;
  const int j = blockIdx.y;
  const int nk = L.km + 1;
  const int k = blockIdx.z % nk, f = blockIdx.z / nk;
  if (i > L.im) return;
  const int l = (DIR == 0) ? i : (DIR == 1 ? j : k);
  const bool inside = (l >= a.o_lo && l <= a.o_hi);
  if (!inside && a.epi != EPI_STOREZ) return;
  const long long x = L.idx(i, j, k);
  double* __restrict__ out = a.out[f];
  if (!inside) { out[x] = 0.0; return; }
  const double* __restrict__ v = a.in[f] + x;
  const long long s = (DIR == 0) ? 1 : (DIR == 1 ? L.sj : L.sk);
  const int n = a.op.n, nt = a.op.ntype;
  const bool lo = (nt == 1 || nt == 4), hi = (nt == 2 || nt == 4);
  double d;
  if (lo && l == 0) d = -0.5 * v[2 * s] + 2.0 * v[s] - 1.5 * v[0];
  else if (lo && l == 1) d = 0.5 * (v[s] - v[-s]);
  else if (lo && l == 2) d = (2.0 / 3.0) * (v[s] - v[-s]) - (1.0 / 12.0) * (v[2 * s] - v[-2 * s]);
  else if (hi && l == n) d = 0.5 * v[-2 * s] - 2.0 * v[-s] + 1.5 * v[0];
  else if (hi && l == n - 1) d = 0.5 * (v[s] - v[-s]);
  else if (hi && l == n - 2) d = (2.0 / 3.0) * (v[s] - v[-s]) - (1.0 / 12.0) * (v[2 * s] - v[-2 * s]);
  else d = 0.75 * (v[s] - v[-s]) - 0.15 * (v[2 * s] - v[-2 * s]) + (1.0 / 60.0) * (v[3 * s] - v[-3 * s]);
  out[x] = (a.epi == EPI_ADD) ? out[x] + d : d;
}

// ---------------------------------------------------------------------------------
// noslip(ndir,tw) (src/bc.F90:6306-6723; nondimen, no species, turbmode none, no wall
// blowing): on the face node u=0, T=tw, p=(4 p1 - p2)/3 from the two interior neighbours,
// rho=thermal(p,T) (fludyna.F90:60), q=fvar2q(rho,vel,p) (fludyna.F90:312-376).
// Threads run over the two other indices (a fastest).
// ---------------------------------------------------------------------------------
template <int DIR>
__global__ void k_noslip(const Layout L, double* __restrict__ pool, const Thermo th, const int side, const double tw) {
  const int na = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int a = blockIdx.x * PW_T + threadIdx.x, b = blockIdx.y;
  if (a >= na) return;
  const int dm = (DIR == 0) ? L.im : (DIR == 1 ? L.jm : L.km);
  const int l = side ? dm : 0;
  const long long sd = (DIR == 0) ? 1 : (DIR == 1 ? L.sj : L.sk), sg = side ? -sd : sd;
  const long long x = (DIR == 0) ? L.idx(l, a, b) : (DIR == 1 ? L.idx(a, l, b) : L.idx(a, b, l));
  const long long fs = L.fstride;
  const double pe = (1.0 / 3.0) * (4.0 * pool[S_PRS * fs + x + sg] - pool[S_PRS * fs + x + 2 * sg]);
  const double rho = th.rho_of(pe, tw);
  pool[(S_VEL + 0) * fs + x] = 0.0; pool[(S_VEL + 1) * fs + x] = 0.0; pool[(S_VEL + 2) * fs + x] = 0.0;
  pool[S_PRS * fs + x] = pe;
  pool[S_TMP * fs + x] = tw;
  pool[S_RHO * fs + x] = rho;
  pool[(S_Q + 0) * fs + x] = rho;
  pool[(S_Q + 1) * fs + x] = 0.0; pool[(S_Q + 2) * fs + x] = 0.0; pool[(S_Q + 3) * fs + x] = 0.0;
  // fvar2q with pressure (nondimen) or temperature (dimensional), bc.F90:6331-6348; velocity is zero
  pool[(S_Q + 4) * fs + x] = th.nondimen ? pe * th.const6 : rho * (tw * th.cotem());
}

// src_chan body force (src/solver.F90:341-352): qrhs(2:4)+=force*jacob,
// qrhs(5)+=(force.ubulk)*jacob on all nodes 0..im,0..jm,0..km ; fe = force.ubulk
__global__ void k_add_force(const Layout L, double* __restrict__ pool, const double* __restrict__ src) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  const double jac = pool[S_JAC * fs + x];
#pragma unroll
  for (int m = 1; m < 5; ++m) pool[(S_QRHS + m) * fs + x] = pool[(S_QRHS + m) * fs + x] + src[m - 1] * jac;
}

// ---------------------------------------------------------------------------------
// Per-step diagnostics of rkfirst / steploop as block-level partial results (the caller applies psum / pmax
// and the reference's normalisation): one CTA per (j,k) row, fixed-order trees => deterministic.
//   RED_TGV : kenergycal, enstophycal, diss_rate_cal (src/statistic.F90:938, :871, :994), nodes 1..im,1..jm,1..km
//   RED_CFL : cflcal's deltai, deltaj, deltak (src/commcal.F90:27-74), nodes 0..im,0..jm,0..km, maxima
//   RED_CHAN: massfluxchan (src/statistic.F90:1437-1476) and fbcxchan (:1303-1367); rows j = 1..jm, the rows next
//             to a wall this block owns also carry the wall friction of j = 0 / j = jm
// ---------------------------------------------------------------------------------
enum { RED_TGV = 0, RED_CFL = 1, RED_CHAN = 2 };
template <int KIND>
__global__ void k_reduce_rows(const Layout L, const double* __restrict__ pool, const double* __restrict__ yc, const Thermo th,
                              double* __restrict__ partial, const int j0, const int k0, const int wall_lo, const int wall_hi) {
  const int j = j0 + blockIdx.x, k = k0 + blockIdx.y;
  const long long fs = L.fstride;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int i = (KIND == RED_CFL ? 0 : 1) + threadIdx.x; i <= L.im; i += PW_T) {
    const long long x = L.idx(i, j, k);
    if (KIND == RED_TGV) {
      Grad g;
      load_grad(pool, fs, x, g);
      const double r = pool[S_RHO * fs + x];
      const double u = pool[(S_VEL + 0) * fs + x], v = pool[(S_VEL + 1) * fs + x], w = pool[(S_VEL + 2) * fs + x];
      a0 += r * (u * u + v * v + w * w);
      const double o1 = g.dv[2][1] - g.dv[1][2], o2 = g.dv[0][2] - g.dv[2][0], o3 = g.dv[1][0] - g.dv[0][1];
      a1 += r * (o1 * o1 + o2 * o2 + o3 * o3);
      const double miu = th.miu(pool[S_TMP * fs + x]);
      const double s11 = g.dv[0][0], s12 = 0.5 * (g.dv[0][1] + g.dv[1][0]), s13 = 0.5 * (g.dv[0][2] + g.dv[2][0]);
      const double s22 = g.dv[1][1], s23 = 0.5 * (g.dv[1][2] + g.dv[2][1]), s33 = g.dv[2][2];
      const double div = s11 + s22 + s33;
      a2 += 2.0 * miu * (s11 * s11 + s22 * s22 + s33 * s33 + 2.0 * (s12 * s12 + s13 * s13 + s23 * s23) -
                         (1.0 / 3.0) * (div * div));
    } else if (KIND == RED_CFL) {
      const double u = pool[(S_VEL + 0) * fs + x], v = pool[(S_VEL + 1) * fs + x], w = pool[(S_VEL + 2) * fs + x];
      const double css = th.sos(pool[S_TMP * fs + x]);
      double bar[3], cs[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double d1 = pool[(S_DXI + 3 * d + 0) * fs + x], d2 = pool[(S_DXI + 3 * d + 1) * fs + x],
                     d3 = pool[(S_DXI + 3 * d + 2) * fs + x];
        bar[d] = d1 * u + d2 * v + d3 * w;
        cs[d] = css * sqrt(d1 * d1 + d2 * d2 + d3 * d3);
      }
      a0 = fmax(fmax(a0, bar[0]), fmax(bar[0] - cs[0], bar[0] + cs[0]));
      a1 = fmax(fmax(a1, bar[1]), fmax(bar[1] - cs[1], bar[1] + cs[1]));
      a2 = fmax(fmax(a2, bar[2]), fmax(bar[2] - cs[2], bar[2] + cs[2]));
    } else {
      const long long xm = x - L.sj;
      const double dy = yc[x] - yc[xm];
      a0 += 0.5 * (pool[(S_Q + 1) * fs + x] + pool[(S_Q + 1) * fs + xm]) * dy;
      if (wall_lo && j == 1) {
        Grad g;
        load_grad(pool, fs, xm, g);
        a1 += th.miu(pool[S_TMP * fs + xm]) * g.dv[0][1];
      }
      if (wall_hi && j == L.jm) {
        Grad g;
        load_grad(pool, fs, x, g);
        a1 -= th.miu(pool[S_TMP * fs + x]) * g.dv[0][1];
      }
    }
  }
  __shared__ double s0[PW_T], s1[PW_T], s2[PW_T];
  s0[threadIdx.x] = a0; s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
  __syncthreads();
  for (int h = PW_T / 2; h > 0; h >>= 1) {
    if (threadIdx.x < h) {
      if (KIND == RED_CFL) {
        s0[threadIdx.x] = fmax(s0[threadIdx.x], s0[threadIdx.x + h]);
        s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + h]);
        s2[threadIdx.x] = fmax(s2[threadIdx.x], s2[threadIdx.x + h]);
      } else {
        s0[threadIdx.x] += s0[threadIdx.x + h]; s1[threadIdx.x] += s1[threadIdx.x + h]; s2[threadIdx.x] += s2[threadIdx.x + h];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long row = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    partial[3 * row] = s0[0]; partial[3 * row + 1] = s1[0]; partial[3 * row + 2] = s2[0];
  }
}
template <bool ISMAX>
__global__ void k_reduce_final(const double* __restrict__ partial, const long long nrows, double* __restrict__ out3) {
  __shared__ double s[3][256];
  double a[3] = {0.0, 0.0, 0.0};
  for (long long r = threadIdx.x; r < nrows; r += 256)
#pragma unroll
    for (int m = 0; m < 3; ++m) a[m] = ISMAX ? fmax(a[m], partial[3 * r + m]) : a[m] + partial[3 * r + m];
#pragma unroll
  for (int m = 0; m < 3; ++m) s[m][threadIdx.x] = a[m];
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h)
#pragma unroll
      for (int m = 0; m < 3; ++m)
        s[m][threadIdx.x] = ISMAX ? fmax(s[m][threadIdx.x], s[m][threadIdx.x + h]) : s[m][threadIdx.x] + s[m][threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x < 3) out3[threadIdx.x] = s[threadIdx.x][0];
}

// ---------------------------------------------------------------------------------
// face pack / unpack (multi-block exchange).  Buffer order: [field][p2][l][p1], p1 fastest.
// side 0 packs planes l (sent to the low neighbour), side 1 packs planes dm-l.
// unpack side 1 (from the high neighbour's side-0 buffer) -> dm+l, l==0 averages node dm;
// unpack side 0 (from the low neighbour's side-1 buffer)  -> -l,   l==0 averages node 0.
// (src/parallel.F90:4180-4218 dataswap, :4941-4990 qswap)
// ---------------------------------------------------------------------------------
template <int DIR, bool PACK>
__global__ void k_face(const Layout L, const FieldList fl, const int side, const int l0, const int l1,
                       double* __restrict__ buf) {
  const int n1 = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int n2 = (DIR == 2) ? L.jm + 1 : L.km + 1;
  const int np = l1 - l0 + 1;
  const int p1 = blockIdx.x * blockDim.x + threadIdx.x;
  if (p1 >= n1) return;
  const int p2 = blockIdx.y / np, l = l0 + blockIdx.y % np;
  double* f = fl.f[blockIdx.z];
  const int dm = (DIR == 0) ? L.im : (DIR == 1 ? L.jm : L.km);
  const long long bi = (((long long)blockIdx.z * n2 + p2) * np + (l - l0)) * n1 + p1;
  int node;
  if (PACK) node = side ? dm - l : l;
  else node = side ? dm + l : -l;
  long long x;
  if (DIR == 0) x = L.idx(node, p1, p2);
  else if (DIR == 1) x = L.idx(p1, node, p2);
  else x = L.idx(p1, p2, node);
  if (PACK) buf[bi] = f[x];
  else if (l == 0) f[x] = 0.5 * (f[x] + buf[bi]);
  else f[x] = buf[bi];
}

// ---------------------------------------------------------------------------------
// inflow(1) (src/bc.F90:1366-1562), outflow(2) / outflow(4) (:3404-3617), farfield(3..6) (:3008-3392),
// slipadibwall(3 | 4) (:7231-7430); no species.  extrapolate(v1,v2,dv=0) = (4 v1 - v2)/3 (commfunc.F90:277-285).
// Threads run over the two other indices of the face (a fastest).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double extrap2(double v1, double v2) { return (1.0 / 3.0) * (4.0 * v1 - v2 - 2.0 * 0.0); }

template <int DIR>
__global__ void k_bcface(const Layout L, double* __restrict__ pool, const Thermo th, const BcArgs a) {
  const int na = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int p1 = blockIdx.x * PW_T + threadIdx.x, p2 = blockIdx.y;
  if (p1 >= na) return;
  const int dm = (DIR == 0) ? L.im : (DIR == 1 ? L.jm : L.km);
  const int l = a.side ? dm : 0;
  const long long sd = (DIR == 0) ? 1 : (DIR == 1 ? L.sj : L.sk), sg = a.side ? -sd : sd;
  const long long x = (DIR == 0) ? L.idx(l, p1, p2) : (DIR == 1 ? L.idx(p1, l, p2) : L.idx(p1, p2, l));
  const long long fs = L.fstride;
  double* rho = pool + S_RHO * fs; double* prs = pool + S_PRS * fs; double* tmp = pool + S_TMP * fs;
  double* v0 = pool + (S_VEL + 0) * fs; double* v1 = pool + (S_VEL + 1) * fs; double* v2 = pool + (S_VEL + 2) * fs;
  bool from_T = true;       // fvar2q with temperature (true) or pressure (false)
  if (a.kind == 11) {       // inflow, face i=0
    const int j = p1, k = p2, nj = L.jm + 1, nk = L.km + 1;
    const double vin0 = a.vel_in[j + (long long)nj * k], vin1 = a.vel_in[j + (long long)nj * (k + (long long)nk)],
                 vin2 = a.vel_in[j + (long long)nj * (k + 2LL * nk)];
    const double rho_ref = rho[x + sg];
    const double css = th.sos(a.tmp_prof[j]);
    v1[x] = vin1; v2[x] = vin2;
    tmp[x] = a.tmp_in[j + (long long)nj * k];
    const double pe = extrap2(prs[x + sg], prs[x + 2 * sg]);
    const double ue = extrap2(v0[x + sg], v0[x + 2 * sg]);
    const double malo = vin0 / css;
    const double blend = 0.5 * (tanh((malo - 1.0) * 6.0) + 1.0);
    const double p = (0.5 * (a.pinf + pe) + 0.5 * rho_ref * css * (vin0 - ue)) * (1.0 - blend) + a.pinf * blend;
    prs[x] = p;
    v0[x] = vin0 + (a.pinf - p) / rho_ref / css;
    rho[x] = th.rho_of(p, tmp[x]);
  } else if (a.kind == 51 && !(DIR == 1 && a.side == 1)) {
    // farfield(3 | 5 | 6), bc.F90:3024-3104, :3235-3390: subsonic characteristic inflow / outflow against the free stream
    double* vv[3] = {v0, v1, v2};
    const double ext[3] = {extrap2(v0[x + sg], v0[x + 2 * sg]), extrap2(v1[x + sg], v1[x + 2 * sg]),
                           extrap2(v2[x + sg], v2[x + 2 * sg])};
    const double pe = extrap2(prs[x + sg], prs[x + 2 * sg]), roe = extrap2(rho[x + sg], rho[x + 2 * sg]);
    const double css = th.sos(tmp[x]);
    const double csse = extrap2(th.sos(tmp[x + sg]), th.sos(tmp[x + 2 * sg]));
    const double vn = vv[DIR][x], vne = ext[DIR], vninf = a.vinf[DIR];
    const bool inflow = a.side ? (vn <= 0.0) : (vn >= 0.0);
    if (inflow) {
      const double rho0 = rho[x];
      double vnew, pnew;
      if (!a.side) {
        vnew = 0.5 * (a.pinf - pe) / (rho0 * css) + 0.5 * (vninf + vne);
        pnew = 0.5 * (a.pinf + pe) + 0.5 * rho0 * css * (vninf - vne);
      } else {
        vnew = -0.5 * (a.pinf - pe) / (rho0 * css) + 0.5 * (vninf + vne);
        pnew = 0.5 * (a.pinf + pe) - 0.5 * rho0 * css * (vninf - vne);
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) vv[m][x] = a.vinf[m];
      vv[DIR][x] = vnew;
      prs[x] = pnew;
      rho[x] = a.roinf * pow(pnew / a.pinf, 1.0 / th.gamma);
    } else {
      const double pnew = a.pinf;
      prs[x] = pnew;
      rho[x] = roe + (pnew - pe) / csse / csse;
#pragma unroll
      for (int m = 0; m < 3; ++m) vv[m][x] = ext[m];
      vv[DIR][x] = a.side ? vne + (pe - pnew) / roe / csse : vne - (pe - pnew) / roe / csse;
    }
    tmp[x] = th.T_of(prs[x], rho[x]);
    from_T = false;
  } else if (a.kind == 421) {              // slipadibwall(3 | 4), bc.F90:7231-7430: slip, adiabatic
    const double pe = extrap2(prs[x + sg], prs[x + 2 * sg]);
    const double te = extrap2(tmp[x + sg], tmp[x + 2 * sg]);
    const double ue = extrap2(v0[x + sg], v0[x + 2 * sg]);
    const double ve = extrap2(v1[x + sg], v1[x + 2 * sg]);
    v0[x] = ue; v1[x] = a.side ? ve : 0.0; v2[x] = 0.0;      // jmax keeps the extrapolated v (bc.F90:7386-7390)
    tmp[x] = te; prs[x] = pe;
    rho[x] = th.rho_of(pe, te);
    from_T = false;
  } else if (a.kind == 21 && DIR == 0) {   // outflow at imax: first-order copy
    v0[x] = v0[x + sg]; v1[x] = v1[x + sg]; v2[x] = v2[x + sg];
    prs[x] = prs[x + sg]; tmp[x] = tmp[x + sg];
    rho[x] = th.rho_of(prs[x], tmp[x]);
  } else {                  // outflow at jmax (21) / farfield at jmax (51): second-order extrapolation
    const double ue = extrap2(v0[x + sg], v0[x + 2 * sg]), ve = extrap2(v1[x + sg], v1[x + 2 * sg]),
                 we = extrap2(v2[x + sg], v2[x + 2 * sg]);
    const double pe = extrap2(prs[x + sg], prs[x + 2 * sg]), roe = extrap2(rho[x + sg], rho[x + 2 * sg]);
    if (a.kind == 51) {
      prs[x] = pe; rho[x] = roe;
    } else {
      const double css = th.sos(tmp[x]);
      const double ub = v1[x];
      if (ub >= css) { prs[x] = pe; rho[x] = roe; }
      else {
        const double te = extrap2(tmp[x + sg], tmp[x + 2 * sg]);
        const double alpha = 0.25;
        const double p = (prs[x] + alpha * a.deltat * a.pinf + rho[x] * css * (ve - v1[x])) / (1.0 + alpha * a.deltat);
        prs[x] = p;
        tmp[x] = te;
        rho[x] = th.rho_of(p, te);
      }
      from_T = !th.nondimen;          // bc.F90:3596-3605
    }
    v0[x] = ue; v1[x] = ve; v2[x] = we;
    if (a.kind == 51 || th.nondimen) tmp[x] = th.T_of(prs[x], rho[x]);
    else rho[x] = th.rho_of(prs[x], tmp[x]);      // dimensional outflow: density from p and the (old or extrapolated) T
  }
  const double r = rho[x], u = v0[x], v = v1[x], w = v2[x];
  pool[(S_Q + 0) * fs + x] = r; pool[(S_Q + 1) * fs + x] = r * u; pool[(S_Q + 2) * fs + x] = r * v;
  pool[(S_Q + 3) * fs + x] = r * w;
  const double var1 = 0.5 * (u * u + v * v + w * w);
  pool[(S_Q + 4) * fs + x] = from_T ? r * (tmp[x] * th.cotem() + var1) : prs[x] * th.const6 + r * var1;
}

// ---------------------------------------------------------------------------------
// host-layout (dense Fortran box (im+11)(jm+11)(km+11)) staging buffer -> padded device field.  Uploads go
// through one contiguous pinned-memory DMA (54 GB/s on the bench box) and this re-pitch, instead of a pitched
// cudaMemcpy2D (30 GB/s).
// ---------------------------------------------------------------------------------
__global__ void k_repitch(const Layout L, const double* __restrict__ src, double* __restrict__ dst) {
  const int w = L.im + 1 + 2 * ASTR_HM;
  const int i = blockIdx.x * PW_T + threadIdx.x;
  if (i >= w) return;
  const long long row = (long long)blockIdx.z * gridDim.y + blockIdx.y;      // j' + njt*k'
  dst[(ASTR_IOFF - ASTR_HM) + i + (long long)L.pitch * row] = src[i + (long long)w * row];
}

// ---------------------------------------------------------------------------------
// checkpoint staging: the datasets of writeflfed / readcheckpoint are node arrays without halos
// (rho(0:im,0:jm,0:km) ..., src/readwrite.F90:1448-1453, :1974-1984)
// ---------------------------------------------------------------------------------
template <bool PACK>
__global__ void k_dense(const Layout L, double* __restrict__ field, double* __restrict__ dense) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long t = (long long)i + (long long)(L.im + 1) * ((long long)j + (long long)(L.jm + 1) * k);
  if (PACK) dense[t] = field[L.idx(i, j, k)];
  else field[L.idx(i, j, k)] = dense[t];
}
// updateq (src/fludyna.F90:254-300): fvar2q from density, velocity and TEMPERATURE (:425-429)
__global__ void k_updateq(const Layout L, double* __restrict__ pool, const Thermo th) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  const double r = pool[S_RHO * fs + x], u = pool[S_VEL * fs + x], v = pool[(S_VEL + 1) * fs + x], w = pool[(S_VEL + 2) * fs + x];
  pool[S_Q * fs + x] = r; pool[(S_Q + 1) * fs + x] = r * u; pool[(S_Q + 2) * fs + x] = r * v; pool[(S_Q + 3) * fs + x] = r * w;
  const double var1 = 0.5 * (u * u + v * v + w * w);
  pool[(S_Q + 4) * fs + x] = r * (pool[S_TMP * fs + x] * th.cotem() + var1);
}

// ---------------------------------------------------------------------------------
// spongefilter_layer (src/sponge_layer.F90:67-319): damped 7-point average of q over the layer box,
// Jacobi style -- pass 1 writes the qrhs slots (dead between the RK update and the next rhscal), pass 2
// copies them back.  coef is the box-shaped sponge_damp_coef in Fortran order.
// ---------------------------------------------------------------------------------
template <bool COPYBACK>
__global__ void k_sponge(const Layout L, double* __restrict__ pool, const Box b, const double* __restrict__ coef) {
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  if (COPYBACK) {
#pragma unroll
    for (int n = 0; n < 5; ++n) pool[(S_Q + n) * fs + x] = pool[(S_QRHS + n) * fs + x];
    return;
  }
  const int ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1;
  const double var1 = coef[(long long)(i - b.lo[0]) + (long long)ni * ((j - b.lo[1]) + (long long)nj * (k - b.lo[2]))];
#pragma unroll
  for (int n = 0; n < 5; ++n) {
    const double* q = pool + (S_Q + n) * fs + x;
    pool[(S_QRHS + n) * fs + x] =
        (1.0 - var1) * q[0] + (1.0 / 6.0) * var1 * (q[1] + q[-1] + q[L.sj] + q[-L.sj] + q[L.sk] + q[-L.sk]);
  }
}

// ---------------------------------------------------------------------------------
// crash control (lcracon), src/mainloop.F90:709-1198.  nodestat <= 0 everywhere (no immersed body).
// ---------------------------------------------------------------------------------
// crashcheck (:725-768): fluid nodes whose density is not >= 0 become critical nodes
__global__ void k_crashcheck(const Layout L, const double* __restrict__ pool, double* __restrict__ crinod,
                             unsigned long long* __restrict__ count) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long x = L.idx(i, j, k);
  if (!(pool[S_Q * L.fstride + x] >= 0.0)) { crinod[x] = 1.0; atomicAdd(count, 1ull); }
}
// crinod_expansion (:985-1033): every flagged node of -1..dim+1 flags its 27-neighbourhood; count = 27 x flagged
template <bool COPYBACK>
__global__ void k_crinod_dilate(const Layout L, double* __restrict__ crinod, double* __restrict__ tmp,
                                unsigned long long* __restrict__ count) {
  Box b = {{-2, -2, -2}, {L.im + 2, L.jm + 2, L.km + 2}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long x = L.idx(i, j, k);
  if (COPYBACK) { crinod[x] = tmp[x]; return; }
  bool any = false;
  for (int kk = -1; kk <= 1; ++kk)
    for (int jj = -1; jj <= 1; ++jj)
      for (int ii = -1; ii <= 1; ++ii) {
        const int i1 = i + ii, j1 = j + jj, k1 = k + kk;
        if (i1 < -1 || i1 > L.im + 1 || j1 < -1 || j1 > L.jm + 1 || k1 < -1 || k1 > L.km + 1) continue;
        any = any || (crinod[L.idx(i1, j1, k1)] != 0.0);
      }
  tmp[x] = any ? 1.0 : 0.0;
  if (i >= -1 && i <= L.im + 1 && j >= -1 && j <= L.jm + 1 && k >= -1 && k <= L.km + 1 && crinod[x] != 0.0)
    atomicAdd(count, 27ull);
}
// crashfix (:1084-1095), detection: rho / prs / tmp under eps; the repairs follow in storage order (k_crashfix_apply)
__global__ void k_crashfix_flag(const Layout L, const double* __restrict__ pool, double* __restrict__ crinod,
                                const double eps_rho, const double eps_prs, const double eps_tmp,
                                unsigned long long* __restrict__ count, long long* __restrict__ list, const long long cap) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long fs = L.fstride, x = L.idx(i, j, k);
  if (pool[S_RHO * fs + x] >= eps_rho && pool[S_PRS * fs + x] >= eps_prs && pool[S_TMP * fs + x] >= eps_tmp) return;
  crinod[x] = 1.0;
  const unsigned long long slot = atomicAdd(count, 1ull);
  if ((long long)slot < cap) list[slot] = (long long)i + (long long)(L.im + 1) * ((long long)j + (long long)(L.jm + 1) * k);
}
// crashfix (:1096-1140), repairs: one thread walks the sorted list, so that a later node sees the earlier repairs
// exactly as the reference's i-fastest loop does.  The list is empty in a healthy run and short in a sick one.
__global__ void k_crashfix_apply(const Layout L, double* __restrict__ pool, const Thermo th,
                                 const long long* __restrict__ list, const CrashFixArgs a,
                                 unsigned long long* __restrict__ fixed) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long fs = L.fstride;
  unsigned long long counter = 0;
  for (long long t = 0; t < a.n; ++t) {
    const long long key = list[t];
    const int i = (int)(key % (L.im + 1)), j = (int)((key / (L.im + 1)) % (L.jm + 1)), k = (int)(key / ((long long)(L.im + 1) * (L.jm + 1)));
    double qavg[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int norm = 0;
    for (int kk = -1; kk <= 1; ++kk)
      for (int jj = -1; jj <= 1; ++jj)
        for (int ii = -1; ii <= 1; ++ii) {
          if (ii == 0 && jj == 0 && kk == 0) continue;
          if (a.g0[0] + i + ii < 0 || a.g0[0] + i + ii > a.ia || a.g0[1] + j + jj < 0 || a.g0[1] + j + jj > a.ja) continue;
          const long long y = L.idx(i + ii, j + jj, k + kk);
          if (pool[S_RHO * fs + y] >= 0.0 && pool[S_PRS * fs + y] >= 0.0 && pool[S_TMP * fs + y] >= 0.0) {
#pragma unroll
            for (int m = 0; m < 5; ++m) qavg[m] = qavg[m] + pool[(S_Q + m) * fs + y];
            norm += 1;
          }
        }
    if (norm >= 1) {
      const long long x = L.idx(i, j, k);
#pragma unroll
      for (int m = 0; m < 5; ++m) pool[(S_Q + m) * fs + x] = qavg[m] / (double)norm;
      q2fvar_node(pool, fs, x, th);
      counter += 1;
    }
  }
  *fixed = counter;
}
// nf fields of the common Layout, nodes of a box only (databakup restores q(0:im,0:jm,0:km,:), :941)
__global__ void k_copy_box(const Layout L, double* __restrict__ dst, const double* __restrict__ src, const int nf, const Box b) {
  int i, j, k;
  if (!box_node(b, i, j, k)) return;
  const long long x = L.idx(i, j, k);
  for (int m = 0; m < nf; ++m) dst[m * L.fstride + x] = src[m * L.fstride + x];
}

// ---------------------------------------------------------------------------------
// Fused pack + peer-to-peer halo exchange (replaces pack -> ncclSend/Recv -> unpack).
// SEND: every CTA first waits until the neighbour has consumed what this rank wrote into the
// neighbour's receive window last time (ack flag, written by the neighbour into THIS rank's
// memory), then stores its share of the face planes straight into the neighbour's window over
// NVLink; the last CTA of a side to finish publishes `ready = seq` in the neighbour's memory.
// RECV: every CTA waits for `ready >= seq` in its own memory, unpacks from its own window
// (L1-bypassing loads), and the last CTA acknowledges to the neighbour.
// Same element order, averaging of shared nodes and plane ranges as k_face above; for i-faces
// the plane index runs fastest so that the strided gather touches each sector once.
// A wait that outlasts a.timeout_cycles (cfg.xchg_timeout_ms; default 60 s, < 0 waits forever like ncclRecv)
// POISONS the exchange: *err is set, the CTA neither moves data nor signals, and every later exchange kernel
// of this rank returns at once; the neighbours then time out in turn.  The host sees the sticky error at the
// next API call that synchronises (p2p_check in api.cu).  Nothing is ever unpacked from, or written over, a
// window whose flag did not arrive.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Window order (both sides agree on it): DIR 0: [field][k][j][plane], plane fastest, so that the strided gather of
// an i face touches each sector once; DIR 1, 2: [field][p2][plane][i], i fastest.  One CTA per (field, p2) slice
// of np * n1 contiguous window values (5 130 CTAs of ~20 KB for a 5-field 513^2 face pair; round 1 ran 77 000 CTAs
// of 2 KB, each paying the flag poll, the system fence and the counter): the threads run along the slice, so every
// warp instruction stores 256 contiguous bytes into the neighbour's window.
template <int DIR, bool SEND>
__global__ void __launch_bounds__(256) k_xface(const Layout L, const FieldList fl, const XArgs a) {
  const int side = blockIdx.z;
  const XSide& xs = a.s[side];
  if (!xs.active) return;
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    int ok = (*reinterpret_cast<volatile unsigned int*>(a.err) == 0u);
    const long long t0 = clock64();
    while (ok && ld_acquire_sys(xs.wait_flag) < xs.wait_val) {
      if (a.timeout_cycles > 0 && clock64() - t0 > a.timeout_cycles) { atomicExch(a.err, 1u); ok = 0; }
      else if (*reinterpret_cast<volatile unsigned int*>(a.err) != 0u) ok = 0;    // another CTA gave up
      else __nanosleep(64);
    }
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) return;      // poisoned: no data phase, no signal
  const int n1 = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int n2 = (DIR == 2) ? L.jm + 1 : L.km + 1;
  const int np = a.l1 - a.l0 + 1;
  const int dm = (DIR == 0) ? L.im : (DIR == 1 ? L.jm : L.km);
  // one CTA per (field, p2, side): np * n1 values, contiguous in the window
  const int p2 = blockIdx.x, fld = blockIdx.y;
  double* __restrict__ f = fl.f[fld];
  const long long wb = ((long long)fld * n2 + p2) * (long long)n1 * np;
  if (DIR == 0) {
    // thread t moves window element t = j * np + plane: consecutive threads write consecutive window addresses
    // (dense 256-byte remote stores per warp instruction) and read the np neighbouring nodes of a few rows
    const int total = n1 * np;
    const int dj = 256 / np, dl = 256 - dj * np;
    int j = threadIdx.x / np, lr = threadIdx.x - j * np;
    for (int t = threadIdx.x; t < total; t += 256) {
      const int l = a.l0 + lr;
      const int node = SEND ? (side ? dm - l : l) : (side ? dm + l : -l);
      const long long x = L.idx(node, j, p2);
      if (SEND) xs.remote[wb + t] = f[x];
      else {
        double v = __ldcg(xs.local + wb + t);
        if (l == 0) v = 0.5 * (f[x] + v);
        f[x] = v;
      }
      j += dj; lr += dl;
      if (lr >= np) { lr -= np; ++j; }
    }
  } else {
    for (int lr = 0; lr < np; ++lr) {
      const int l = a.l0 + lr;
      const int node = SEND ? (side ? dm - l : l) : (side ? dm + l : -l);
      const long long x0 = (DIR == 1) ? L.idx(0, node, p2) : L.idx(0, p2, node);
      const long long bi = wb + (long long)lr * n1;
#pragma unroll 3
      for (int i = threadIdx.x; i < n1; i += 256) {
        if (SEND) xs.remote[bi + i] = f[x0 + i];
        else {
          double v = __ldcg(xs.local + bi + i);
          if (l == 0) v = 0.5 * (f[x0 + i] + v);
          f[x0 + i] = v;
        }
      }
    }
  }
  // one system-scope release per CTA: the barrier orders the CTA's stores before thread 0's fence
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(xs.counter, 1u);
    if (prev == gridDim.x * gridDim.y - 1) {
      *xs.counter = 0u;
      __threadfence_system();
      st_release_sys(xs.signal_flag, xs.signal_val);
    }
  }
}

}  // namespace

#define LAUNCH_CHECK()                        \
  do {                                        \
    astr_count_launch();                      \
    CUDA_OK(cudaGetLastError());              \
  } while (0)

int pw_halo_wrap(const Layout& L, const FieldList& fl, int dir, int mode, cudaStream_t st) {
  const int na = (dir == 0) ? L.jm + 1 : L.im + 1;
  const int nb = (dir == 2) ? L.jm + 1 : L.km + 1;
  dim3 grid((na + PW_T - 1) / PW_T, nb, fl.nf);
  if (dir == 0) { grid.x = (na * 8 + PW_T - 1) / PW_T; k_halo_wrap<0><<<grid, PW_T, 0, st>>>(L, fl, mode); }
  else if (dir == 1) k_halo_wrap<1><<<grid, PW_T, 0, st>>>(L, fl, mode);
  else k_halo_wrap<2><<<grid, PW_T, 0, st>>>(L, fl, mode);
  LAUNCH_CHECK();
  return 0;
}

int pw_q2fvar(const Layout& L, double* pool, const Thermo& th, const Box& b, cudaStream_t st) {
  if (box_empty(b)) return 0;
  k_q2fvar<<<box_grid(b), PW_T, 0, st>>>(L, pool, th, b);
  LAUNCH_CHECK();
  return 0;
}

int pw_visc(const Layout& L, double* pool, const Thermo& th, const Box& b, cudaStream_t st) {
  if (box_empty(b)) return 0;
  k_visc<<<box_grid(b), PW_T, 0, st>>>(L, pool, th, b);
  LAUNCH_CHECK();
  return 0;
}

int pw_visc_flux(const Layout& L, double* pool, const Thermo& th, const FluxRanges& fr, int ndims, int store_shell,
                 cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  if (ndims == 3) k_visc_flux<7><<<box_grid(b), PW_T, 0, st>>>(L, pool, th, fr, store_shell);
  else k_visc_flux<3><<<box_grid(b), PW_T, 0, st>>>(L, pool, th, fr, store_shell);    // 2-D: no zeta flux (solver.F90:2776)
  LAUNCH_CHECK();
  return 0;
}

int pw_materialise_grad(const Layout& L, double* pool, double* out, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_materialise_grad<<<box_grid(b), PW_T, 0, st>>>(L, pool, out);
  LAUNCH_CHECK();
  return 0;
}

int pw_flux(const Layout& L, double* pool, const Box& b, int dmask, const FluxRanges& fr, int diffterm,
            cudaStream_t st) {
  if (box_empty(b)) return 0;
  dim3 grid = box_grid(b);
  switch (dmask) {
    case 1: k_flux<1><<<grid, PW_T, 0, st>>>(L, pool, b, fr, diffterm); break;
    case 2: k_flux<2><<<grid, PW_T, 0, st>>>(L, pool, b, fr, diffterm); break;
    case 4: k_flux<4><<<grid, PW_T, 0, st>>>(L, pool, b, fr, diffterm); break;
    case 3: k_flux<3><<<grid, PW_T, 0, st>>>(L, pool, b, fr, diffterm); break;
    case 7: k_flux<7><<<grid, PW_T, 0, st>>>(L, pool, b, fr, diffterm); break;
    default: return astr_fail_msg("pw_flux: unsupported direction mask");
  }
  LAUNCH_CHECK();
  return 0;
}

int pw_rk_update(const Layout& L, double* pool, const Thermo& th, const RkCoef& rk, const double* src,
                 cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_rk_update<<<box_grid(b), PW_T, 0, st>>>(L, pool, th, rk, src);
  LAUNCH_CHECK();
  return 0;
}

int pw_sum_qrhs(const Layout& L, double* pool, const double* src, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_sum_qrhs<<<box_grid(b), PW_T, 0, st>>>(L, pool, src);
  LAUNCH_CHECK();
  return 0;
}

int pw_bulk(const Layout& L, const double* pool, const double* yc, double* partial, double* out4, cudaStream_t st) {
  const int nk = L.km == 0 ? 1 : L.km;
  dim3 grid(L.jm, nk);
  k_bulk_rows<<<grid, PW_T, 0, st>>>(L, pool, yc, partial, L.km == 0 ? 0 : 1);
  LAUNCH_CHECK();
  k_bulk_final<<<1, 256, 0, st>>>(partial, (long long)L.jm * nk, out4);
  LAUNCH_CHECK();
  return 0;
}
int pw_src_coef(const double* bulk4, const double force[3], double* src4, cudaStream_t st) {
  k_src_coef<<<1, 1, 0, st>>>(bulk4, force[0], force[1], force[2], src4);
  LAUNCH_CHECK();
  return 0;
}

int pw_diff6e(int dir, const SweepArgs& a, cudaStream_t st) {
  const Layout& L = a.L;
  for (int f = 0; f < a.nf; ++f)
    if (a.in[f] == a.out[f]) return astr_fail_msg("explicit derivative cannot run in place");
  dim3 grid((L.im + PW_T) / PW_T, L.jm + 1, (L.km + 1) * a.nf);
  if (dir == 0) k_diff6e<0><<<grid, PW_T, 0, st>>>(L, a);
  else if (dir == 1) k_diff6e<1><<<grid, PW_T, 0, st>>>(L, a);
  else k_diff6e<2><<<grid, PW_T, 0, st>>>(L, a);
  LAUNCH_CHECK();
  return 0;
}

int pw_repitch(const Layout& L, const double* stage, double* field, cudaStream_t st) {
  dim3 grid((L.im + 1 + 2 * ASTR_HM + PW_T - 1) / PW_T, L.njt, L.nkt);
  k_repitch<<<grid, PW_T, 0, st>>>(L, stage, field);
  LAUNCH_CHECK();
  return 0;
}

int pw_dense(const Layout& L, double* field, double* dense, bool pack, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  if (pack) k_dense<true><<<box_grid(b), PW_T, 0, st>>>(L, field, dense);
  else k_dense<false><<<box_grid(b), PW_T, 0, st>>>(L, field, dense);
  LAUNCH_CHECK();
  return 0;
}
int pw_updateq(const Layout& L, double* pool, const Thermo& th, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_updateq<<<box_grid(b), PW_T, 0, st>>>(L, pool, th);
  LAUNCH_CHECK();
  return 0;
}

int pw_crashcheck(const Layout& L, const double* pool, double* crinod, unsigned long long* count, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_crashcheck<<<box_grid(b), PW_T, 0, st>>>(L, pool, crinod, count);
  LAUNCH_CHECK();
  return 0;
}
int pw_crinod_dilate(const Layout& L, double* crinod, double* tmp, unsigned long long* count, cudaStream_t st) {
  Box b = {{-2, -2, -2}, {L.im + 2, L.jm + 2, L.km + 2}};
  k_crinod_dilate<false><<<box_grid(b), PW_T, 0, st>>>(L, crinod, tmp, count);
  LAUNCH_CHECK();
  k_crinod_dilate<true><<<box_grid(b), PW_T, 0, st>>>(L, crinod, tmp, count);
  LAUNCH_CHECK();
  return 0;
}
int pw_crashfix_flag(const Layout& L, const double* pool, double* crinod, double eps_rho, double eps_prs, double eps_tmp,
                     unsigned long long* count, long long* list, long long cap, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_crashfix_flag<<<box_grid(b), PW_T, 0, st>>>(L, pool, crinod, eps_rho, eps_prs, eps_tmp, count, list, cap);
  LAUNCH_CHECK();
  return 0;
}
int pw_crashfix_apply(const Layout& L, double* pool, const Thermo& th, const long long* list, const CrashFixArgs& a,
                      unsigned long long* fixed, cudaStream_t st) {
  k_crashfix_apply<<<1, 32, 0, st>>>(L, pool, th, list, a, fixed);
  LAUNCH_CHECK();
  return 0;
}
int pw_copy_box(const Layout& L, double* dst, const double* src, int nf, const Box& b, cudaStream_t st) {
  if (box_empty(b)) return 0;
  k_copy_box<<<box_grid(b), PW_T, 0, st>>>(L, dst, src, nf, b);
  LAUNCH_CHECK();
  return 0;
}

int pw_sponge(const Layout& L, double* pool, const Box& b, const double* coef, cudaStream_t st) {
  if (box_empty(b)) return 0;
  k_sponge<false><<<box_grid(b), PW_T, 0, st>>>(L, pool, b, coef);
  LAUNCH_CHECK();
  k_sponge<true><<<box_grid(b), PW_T, 0, st>>>(L, pool, b, coef);
  LAUNCH_CHECK();
  return 0;
}

int pw_bcface(const Layout& L, double* pool, const Thermo& th, int dir, const BcArgs& a, cudaStream_t st) {
  const int na = (dir == 0) ? L.jm + 1 : L.im + 1, nb = (dir == 2) ? L.jm + 1 : L.km + 1;
  dim3 grid((na + PW_T - 1) / PW_T, nb);
  if (dir == 0) k_bcface<0><<<grid, PW_T, 0, st>>>(L, pool, th, a);
  else if (dir == 1) k_bcface<1><<<grid, PW_T, 0, st>>>(L, pool, th, a);
  else k_bcface<2><<<grid, PW_T, 0, st>>>(L, pool, th, a);
  LAUNCH_CHECK();
  return 0;
}

int pw_noslip(const Layout& L, double* pool, const Thermo& th, int dir, int side, double tw, cudaStream_t st) {
  const int na = (dir == 0) ? L.jm + 1 : L.im + 1, nb = (dir == 2) ? L.jm + 1 : L.km + 1;
  dim3 grid((na + PW_T - 1) / PW_T, nb);
  if (dir == 0) k_noslip<0><<<grid, PW_T, 0, st>>>(L, pool, th, side, tw);
  else if (dir == 1) k_noslip<1><<<grid, PW_T, 0, st>>>(L, pool, th, side, tw);
  else k_noslip<2><<<grid, PW_T, 0, st>>>(L, pool, th, side, tw);
  LAUNCH_CHECK();
  return 0;
}

int pw_add_force(const Layout& L, double* pool, const double* src, cudaStream_t st) {
  Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
  k_add_force<<<box_grid(b), PW_T, 0, st>>>(L, pool, src);
  LAUNCH_CHECK();
  return 0;
}

// what: 0 KE / enstrophy / dissipation sums, 1 CFL maxima, 2 channel mass flux / wall friction sums -> out3
int pw_reduce(const Layout& L, double* pool, const double* yc, const Thermo& th, int what, int wall_lo, int wall_hi,
              double* partial, double* out3, cudaStream_t st) {
  if (what == 0) {
    dim3 grid(L.jm, L.km);
    k_reduce_rows<RED_TGV><<<grid, PW_T, 0, st>>>(L, pool, yc, th, partial, 1, 1, 0, 0);
    LAUNCH_CHECK();
    k_reduce_final<false><<<1, 256, 0, st>>>(partial, (long long)L.jm * L.km, out3);
  } else if (what == 1) {
    dim3 grid(L.jm + 1, L.km + 1);
    k_reduce_rows<RED_CFL><<<grid, PW_T, 0, st>>>(L, pool, yc, th, partial, 0, 0, 0, 0);
    LAUNCH_CHECK();
    k_reduce_final<true><<<1, 256, 0, st>>>(partial, (long long)(L.jm + 1) * (L.km + 1), out3);
  } else {
    const int nk = L.km == 0 ? 1 : L.km;               // 2-D blocks: k = 0 only (statistic.F90:1447-1451)
    dim3 grid(L.jm, nk);
    k_reduce_rows<RED_CHAN><<<grid, PW_T, 0, st>>>(L, pool, yc, th, partial, 1, L.km == 0 ? 0 : 1, wall_lo, wall_hi);
    LAUNCH_CHECK();
    k_reduce_final<false><<<1, 256, 0, st>>>(partial, (long long)L.jm * nk, out3);
  }
  LAUNCH_CHECK();
  return 0;
}

template <bool PACK>
static int face_launch(const Layout& L, const FieldList& fl, int dir, int side, int l0, int l1, double* buf,
                       cudaStream_t st) {
  const int n1 = (dir == 0) ? L.jm + 1 : L.im + 1;
  const int n2 = (dir == 2) ? L.jm + 1 : L.km + 1;
  const int np = l1 - l0 + 1;
  dim3 grid((n1 + PW_T - 1) / PW_T, n2 * np, fl.nf);
  if (dir == 0) k_face<0, PACK><<<grid, PW_T, 0, st>>>(L, fl, side, l0, l1, buf);
  else if (dir == 1) k_face<1, PACK><<<grid, PW_T, 0, st>>>(L, fl, side, l0, l1, buf);
  else k_face<2, PACK><<<grid, PW_T, 0, st>>>(L, fl, side, l0, l1, buf);
  LAUNCH_CHECK();
  return 0;
}
template <bool SEND>
static int xface_launch(const Layout& L, const FieldList& fl, int dir, const XArgs& a, cudaStream_t st) {
  const int n1 = (dir == 0) ? L.jm + 1 : L.im + 1;
  const int n2 = (dir == 2) ? L.jm + 1 : L.km + 1;
  const int np = a.l1 - a.l0 + 1;
  // one CTA per (field, p2) slice of a window (np * n1 values), both sides in one launch
  dim3 grid(n2, fl.nf, 2);
  (void)n1; (void)np;
  if (dir == 0) k_xface<0, SEND><<<grid, 256, 0, st>>>(L, fl, a);
  else if (dir == 1) k_xface<1, SEND><<<grid, 256, 0, st>>>(L, fl, a);
  else k_xface<2, SEND><<<grid, 256, 0, st>>>(L, fl, a);
  LAUNCH_CHECK();
  return 0;
}
int pw_xsend(const Layout& L, const FieldList& fl, int dir, const XArgs& a, cudaStream_t st) {
  return xface_launch<true>(L, fl, dir, a, st);
}
int pw_xrecv(const Layout& L, const FieldList& fl, int dir, const XArgs& a, cudaStream_t st) {
  return xface_launch<false>(L, fl, dir, a, st);
}

int pw_pack(const Layout& L, const FieldList& fl, int dir, int side, int l0, int l1, double* buf,
            cudaStream_t st) {
  return face_launch<true>(L, fl, dir, side, l0, l1, buf, st);
}
int pw_unpack(const Layout& L, const FieldList& fl, int dir, int side, int l0, int l1, const double* buf,
              cudaStream_t st) {
  return face_launch<false>(L, fl, dir, side, l0, l1, const_cast<double*>(buf), st);
}
