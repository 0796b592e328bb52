// astr_b200/csrc/geom.cu -- grid metrics on the device: dx/dxi with the compact operator,
// Jacobian, conservative-form dxi/dx (src/geom.F90:99-700, ndims==3 branch), and the
// coordinate halo fill gridsendrecv (src/parallel.F90:2780-3035).
#include "geom.cuh"

namespace {
constexpr int GT = 128;

// ---- gridsendrecv ------------------------------------------------------------------
// threads over the two other indices; blockIdx.z = coordinate component
template <int DIR>
__device__ __forceinline__ void face_base(const Layout& L, int a, int b, long long& base, long long& sl, int& dm) {
  if (DIR == 0) { base = L.idx(0, a, b); sl = 1; dm = L.im; }
  else if (DIR == 1) { base = L.idx(a, 0, b); sl = L.sj; dm = L.jm; }
  else { base = L.idx(a, b, 0); sl = L.sk; dm = L.km; }
}
struct X3 { double* x[3]; };

// mode 0: single block (parallel.F90:2861-2866); mode 1: pack offsets; mode 2: unpack
template <int DIR>
__global__ void k_xhalo(const Layout L, const X3 xs, const int mode, double* __restrict__ buf_lo,
                        double* __restrict__ buf_hi, const double* __restrict__ from_lo,
                        const double* __restrict__ from_hi) {
  const int na = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int nb = (DIR == 2) ? L.jm + 1 : L.km + 1;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, m = blockIdx.z;
  if (a >= na || b >= nb) return;
  long long base, sl; int dm;
  face_base<DIR>(L, a, b, base, sl, dm);
  double* p = xs.x[m] + base;
  const long long bi = (((long long)m * nb + b) * ASTR_HM) * na + a;  // + (n-1)*na
  const double x0 = p[0], xm = p[dm * sl];
#pragma unroll
  for (int n = 1; n <= ASTR_HM; ++n) {
    if (mode == 0) {
      p[(dm + n) * sl] = xm + (p[n * sl] - x0);
      p[-n * sl] = x0 - (xm - p[(dm - n) * sl]);
    } else if (mode == 1) {
      buf_lo[bi + (long long)(n - 1) * na] = p[n * sl] - x0;          // sendbuf1 (:2812)
      buf_hi[bi + (long long)(n - 1) * na] = p[(dm - n) * sl] - xm;   // sendbuf2 (:2813)
    } else {
      if (from_lo) p[-n * sl] = from_lo[bi + (long long)(n - 1) * na] + x0;     // neighbour's sendbuf2
      else p[-n * sl] = 2.0 * x0 - p[n * sl];                                   // even reflection
      if (from_hi) p[(dm + n) * sl] = from_hi[bi + (long long)(n - 1) * na] + xm; // neighbour's sendbuf1
      else p[(dm + n) * sl] = 2.0 * xm - p[(dm - n) * sl];
    }
  }
}

// ---- Jacobian (geom.F90:340-357) -----------------------------------------------------
struct DX9 { const double* d[9]; };  // d[3*m+n] = d x_m / d xi_n
__global__ void k_jacobian(const Layout L, const DX9 dx, double* __restrict__ jac) {
  const int i = blockIdx.x * GT + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
  if (i > L.im) return;
  const long long x = L.idx(i, j, k);
#define DXV(m, n) dx.d[3 * ((m)-1) + ((n)-1)][x]
  jac[x] = DXV(1, 1) * DXV(2, 2) * DXV(3, 3) + DXV(1, 2) * DXV(2, 3) * DXV(3, 1) +
           DXV(1, 3) * DXV(2, 1) * DXV(3, 2) - DXV(1, 3) * DXV(2, 2) * DXV(3, 1) -
           DXV(1, 2) * DXV(2, 1) * DXV(3, 3) - DXV(1, 1) * DXV(2, 3) * DXV(3, 2);
#undef DXV
}

// ---- 2-D blocks (ndims==2): Jacobian (geom.F90:371-375) and dxi (:520-527) from the four in-plane dx ----------
__global__ void k_jacobian2d(const Layout L, const DX9 dx, double* __restrict__ jac) {
  const int i = blockIdx.x * GT + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
  if (i > L.im) return;
  const long long x = L.idx(i, j, k);
  jac[x] = dx.d[0][x] * dx.d[4][x] - dx.d[1][x] * dx.d[3][x];      // dx(1,1) dx(2,2) - dx(1,2) dx(2,1)
}
struct F9w { double* f[9]; };
__global__ void k_dxi2d(const Layout L, const DX9 dx, const F9w f) {
  const int i = blockIdx.x * GT + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
  if (i > L.im) return;
  const long long x = L.idx(i, j, k);
  f.f[0][x] = dx.d[4][x];        // dxi(1,1) =  dx(2,2)
  f.f[1][x] = -dx.d[1][x];       // dxi(1,2) = -dx(1,2)
  f.f[3][x] = -dx.d[3][x];       // dxi(2,1) = -dx(2,1)
  f.f[4][x] = dx.d[0][x];        // dxi(2,2) =  dx(1,1)
}

// ---- phi = 0.5*(dx(m1,n1)*x(c1) - dx(m2,n2)*x(c2)) on a box (geom.F90:402-510) ------------
struct PhiTerm { const double *da, *xa, *db, *xb; double* out; };
struct PhiTerms { PhiTerm t[6]; };
__global__ void k_phi(const Layout L, const PhiTerms pt, const Box b) {
  const int i = b.lo[0] + blockIdx.x * GT + threadIdx.x;
  if (i > b.hi[0]) return;
  const int jj = blockIdx.y % (b.hi[1] - b.lo[1] + 1), t = blockIdx.y / (b.hi[1] - b.lo[1] + 1);
  const long long x = L.idx(i, b.lo[1] + jj, b.lo[2] + blockIdx.z);
  const PhiTerm& p = pt.t[t];
  p.out[x] = 0.5 * (p.da[x] * p.xa[x] - p.db[x] * p.xb[x]);
}

// ---- dxi /= jacob on 0..n (geom.F90:662-666) ---------------------------------------------
struct F9 { double* f[9]; };
__global__ void k_div_jac(const Layout L, const F9 f, const double* __restrict__ jac) {
  const int i = blockIdx.x * GT + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
  if (i > L.im) return;
  const long long x = L.idx(i, j, k);
  const double jv = jac[x];
#pragma unroll
  for (int n = 0; n < 9; ++n) f.f[n][x] = f.f[n][x] / jv;
}

template <int DIR>
int xhalo_launch(const Layout& L, double* x3[3], int mode, double* bl, double* bh, const double* fl,
                 const double* fh, cudaStream_t st) {
  const int na = (DIR == 0) ? L.jm + 1 : L.im + 1;
  const int nb = (DIR == 2) ? L.jm + 1 : L.km + 1;
  X3 xs = {{x3[0], x3[1], x3[2]}};
  dim3 grid((na + GT - 1) / GT, nb, 3);
  k_xhalo<DIR><<<grid, GT, 0, st>>>(L, xs, mode, bl, bh, fl, fh);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}
int xhalo_any(const Layout& L, double* x3[3], int d, int mode, double* bl, double* bh, const double* fl,
              const double* fh, cudaStream_t st) {
  if (d == 0) return xhalo_launch<0>(L, x3, mode, bl, bh, fl, fh, st);
  if (d == 1) return xhalo_launch<1>(L, x3, mode, bl, bh, fl, fh, st);
  return xhalo_launch<2>(L, x3, mode, bl, bh, fl, fh, st);
}
}  // namespace

int geom_xhalo_single(const Layout& L, double* x3[3], int d, cudaStream_t st) {
  return xhalo_any(L, x3, d, 0, nullptr, nullptr, nullptr, nullptr, st);
}
int geom_xhalo_pack(const Layout& L, double* x3[3], int d, double* buf_lo, double* buf_hi, cudaStream_t st) {
  return xhalo_any(L, x3, d, 1, buf_lo, buf_hi, nullptr, nullptr, st);
}
int geom_xhalo_unpack(const Layout& L, double* x3[3], int d, const double* from_lo, const double* from_hi,
                      cudaStream_t st) {
  return xhalo_any(L, x3, d, 2, nullptr, nullptr, from_lo, from_hi, st);
}

#define TRYG(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

int geom_gridgeom(const Layout& L, const astr_cfg& cfg, cudaStream_t st) {
  const int dimv[3] = {L.im, L.jm, L.km};
  // 1. coordinate halos
  for (int d = 0; d < cfg.ndims; ++d) TRYG(astr_xhalo_exchange(d));   // 2-D: the k planes are never differentiated
  if (cfg.ndims == 2)     // d x_m / d zeta = 0: the slots may hold something else from an earlier use of the scratch pool
    for (int m = 0; m < 3; ++m)
      CUDA_OK(cudaMemsetAsync(astr_slot_ptr(S_SCR + 3 * m + 2), 0, (size_t)L.fstride * sizeof(double), st));
  // 2. dx(m,n) = d x_m / d xi_n -> scratch slot S_SCR + 3m+n   (geom.F90:130-164; 2-D blocks: no zeta sweeps, the
  //    scratch slots are zero-initialised)
  const int nd = cfg.ndims;
  for (int d = 0; d < nd; ++d) {
    int in[3], out[3];
    for (int m = 0; m < 3; ++m) { in[m] = S_G + m; out[m] = S_SCR + 3 * m + d; }
    TRYG(astr_sweep_slots(d, OP_DERIV, in, out, 3, EPI_STORE, 0, dimv[d]));
  }
  int dxs[9];
  for (int n = 0; n < 9; ++n) dxs[n] = S_SCR + n;
  for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(dxs, 9, d, XMODE_SWAP));     // geom.F90:333
  // 3. Jacobian + its halos and shared-node sync (:340-385)
  {
    DX9 dx;
    for (int n = 0; n < 9; ++n) dx.d[n] = astr_slot_ptr(S_SCR + n);
    dim3 grid((L.im + GT) / GT, L.jm + 1, L.km + 1);
    if (nd == 2) k_jacobian2d<<<grid, GT, 0, st>>>(L, dx, astr_slot_ptr(S_JAC));
    else k_jacobian<<<grid, GT, 0, st>>>(L, dx, astr_slot_ptr(S_JAC));
    astr_count_launch();
    CUDA_OK(cudaGetLastError());
    int js = S_JAC;
    for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(&js, 1, d, XMODE_SWAP));
    for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(&js, 1, d, XMODE_SYNC));
  }
  // 4. conservative-form metrics (:399-519): dxi(a,b) accumulates two sweeps
  struct Term { int d, a, b, m1, n1, c1, m2, n2, c2; };
  static const Term terms[18] = {
      {0, 2, 1, 2, 3, 3, 3, 3, 2}, {0, 2, 2, 3, 3, 1, 1, 3, 3}, {0, 2, 3, 1, 3, 2, 2, 3, 1},
      {0, 3, 1, 3, 2, 2, 2, 2, 3}, {0, 3, 2, 1, 2, 3, 3, 2, 1}, {0, 3, 3, 2, 2, 1, 1, 2, 2},
      {1, 1, 1, 3, 3, 2, 2, 3, 3}, {1, 1, 2, 1, 3, 3, 3, 3, 1}, {1, 1, 3, 2, 3, 1, 1, 3, 2},
      {1, 3, 1, 2, 1, 3, 3, 1, 2}, {1, 3, 2, 3, 1, 1, 1, 1, 3}, {1, 3, 3, 1, 1, 2, 2, 1, 1},
      {2, 1, 1, 2, 2, 3, 3, 2, 2}, {2, 1, 2, 3, 2, 1, 1, 2, 3}, {2, 1, 3, 1, 2, 2, 2, 2, 1},
      {2, 2, 1, 3, 1, 2, 2, 1, 3}, {2, 2, 2, 1, 1, 3, 3, 1, 1}, {2, 2, 3, 2, 1, 1, 1, 1, 2}};
  for (int n = 0; n < 9; ++n)
    CUDA_OK(cudaMemsetAsync(astr_slot_ptr(S_DXI + n), 0, (size_t)L.fstride * sizeof(double), st));
  bool seen[9] = {false, false, false, false, false, false, false, false, false};
  if (nd == 2) {      // geom.F90:520-527: the in-plane metrics are plain copies, everything else stays zero
    DX9 dx;
    for (int n = 0; n < 9; ++n) dx.d[n] = astr_slot_ptr(S_SCR + n);
    F9w f;
    for (int n = 0; n < 9; ++n) f.f[n] = astr_slot_ptr(S_DXI + n);
    dim3 grid((L.im + GT) / GT, L.jm + 1, L.km + 1);
    k_dxi2d<<<grid, GT, 0, st>>>(L, dx, f);
    astr_count_launch();
    CUDA_OK(cudaGetLastError());
  }
  for (int d = 0; d < 3 && nd == 3; ++d) {
    PhiTerms pt;
    int in[6], out[6];
    bool add[6];
    for (int t = 0; t < 6; ++t) {
      const Term& T = terms[6 * d + t];
      pt.t[t].da = astr_slot_ptr(S_SCR + 3 * (T.m1 - 1) + (T.n1 - 1));
      pt.t[t].xa = astr_slot_ptr(S_G + T.c1 - 1);
      pt.t[t].db = astr_slot_ptr(S_SCR + 3 * (T.m2 - 1) + (T.n2 - 1));
      pt.t[t].xb = astr_slot_ptr(S_G + T.c2 - 1);
      pt.t[t].out = astr_slot_ptr(S_RAW + t);
      in[t] = S_RAW + t;
      out[t] = S_DXI + 3 * (T.a - 1) + (T.b - 1);
      add[t] = seen[3 * (T.a - 1) + (T.b - 1)];
      seen[3 * (T.a - 1) + (T.b - 1)] = true;
    }
    Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
    b.lo[d] = -ASTR_HM; b.hi[d] = dimv[d] + ASTR_HM;
    dim3 grid((b.hi[0] - b.lo[0] + GT) / GT, (b.hi[1] - b.lo[1] + 1) * 6, b.hi[2] - b.lo[2] + 1);
    k_phi<<<grid, GT, 0, st>>>(L, pt, b);
    astr_count_launch();
    CUDA_OK(cudaGetLastError());
    // terms 0..2 and 3..5 each share one epilogue
    for (int h = 0; h < 2; ++h)
      TRYG(astr_sweep_slots(d, OP_DERIV, in + 3 * h, out + 3 * h, 3, add[3 * h] ? EPI_ADD : EPI_STORE, 0, dimv[d]));
  }
  int dxi[9];
  for (int n = 0; n < 9; ++n) dxi[n] = S_DXI + n;
  for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(dxi, 9, d, XMODE_SWAP));     // :541
  {
    F9 f;
    for (int n = 0; n < 9; ++n) f.f[n] = astr_slot_ptr(S_DXI + n);
    dim3 grid((L.im + GT) / GT, L.jm + 1, L.km + 1);
    k_div_jac<<<grid, GT, 0, st>>>(L, f, astr_slot_ptr(S_JAC));                       // :662-666
    astr_count_launch();
    CUDA_OK(cudaGetLastError());
  }
  for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(dxi, 9, d, XMODE_SWAP));     // :674
  for (int d = 0; d < 3; ++d) TRYG(astr_exchange_slots(dxi, 9, d, XMODE_SYNC));     // :676-680
  (void)cfg;
  return 0;
}
