// astr_b200/csrc/sweep.cu -- the batched tridiagonal line-solve engine (sm_100a).
//
// One kernel family serves every 1-D compact operator of the reference:
//   OP_DERIV  : 6th-order compact first derivative, `fds%central` = df_compact +
//               compact_fd_rhs (src/derivative.F90:171-198, :210-306)
//   OP_FILTER : 10th-order compact low-pass filter, compact_filter + compact_filter_rhs
//               (src/filter.F90:112-144, :156-285)
// both solved with the reference's pre-factored unit-diagonal Thomas recurrences
// (src/commfunc.F90:790-813) -- same boundary (ntype 1/2/4) and interface (ntype 3)
// closure rows, same coefficient tables.
//
// Mapping.  A CTA owns a BUNDLE of 16 neighbouring pencils and the
// whole line (nodes -5..n+5) of each, staged once in shared memory with 16-byte
// cp.async.cg granules (HBM traffic = 1 read + 1 write per node, the algorithmic
// minimum).  j/k sweeps: the 16 pencils are 16 consecutive i => every line position is
// one aligned 128-byte row.  i sweeps: 16 consecutive j rows, each a contiguous line.
// The line is cut in C chunks; thread (pencil p, chunk c) runs the Thomas forward and
// backward recurrences on its chunk IN PLACE with zero carry, the true carries are
// recovered from the C chunk-end values with the pre-computed propagation products
// pf/qb (exact algebra, no truncation), and the last correction x = g + qb*xin is folded
// into the coalesced write-out loop, which also applies the epilogue (store / add).
// The sequential dependence is therefore n/C long instead of n, and 3 CTAs per SM
// overlap load, solve and store phases.
#include "common.cuh"
#include <cstdio>

__constant__ FilterCoef c_fc;

int astr_set_filter_coef(const FilterCoef& fc) {
  cudaError_t e = cudaMemcpyToSymbol(c_fc, &fc, sizeof(FilterCoef));
  if (e != cudaSuccess) return astr_fail("cudaMemcpyToSymbol(c_fc)", e, __FILE__, __LINE__);
  return 0;
}

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

template <int OP> struct OpTraits;
template <> struct OpTraits<OP_DERIV> { static constexpr int H = 2, W = 5; };
template <> struct OpTraits<OP_FILTER> { static constexpr int H = 5, W = 11; };

// Interior right-hand side at phase t (t mod W is a compile-time constant after
// unrolling): the value of node m+k sits in window slot (t+H+k) mod W.
template <int OP>
__device__ __forceinline__ double interior_rhs(const double (&w)[OpTraits<OP>::W], int t) {
  constexpr int H = OpTraits<OP>::H, W = OpTraits<OP>::W;
#define WS(k) w[(t + H + (k)) % W]
  if (OP == OP_DERIV) {
    // src/derivative.F90:296-304
    const double var1 = WS(1) - WS(-1);
    const double var2 = WS(2) - WS(-2);
    return (7.0 / 9.0) * var1 + (1.0 / 36.0) * var2;
  } else {
    // src/filter.F90:271-283
    const double var0 = WS(0) + WS(0);
    const double var1 = WS(1) + WS(-1);
    const double var2 = WS(2) + WS(-2);
    const double var3 = WS(3) + WS(-3);
    const double var4 = WS(4) + WS(-4);
    const double var5 = WS(5) + WS(-5);
    return c_fc.coef10i[0] * var0 + c_fc.coef10i[1] * var1 + c_fc.coef10i[2] * var2 +
           c_fc.coef10i[3] * var3 + c_fc.coef10i[4] * var4 + c_fc.coef10i[5] * var5;
  }
#undef WS
}

// Closure rows.  F(node) reads the pristine line from shared memory.  sf[k] is row k,
// sl[k] is row nrows-nsl+k.
template <int OP, class FN>
__device__ __forceinline__ void closure_first(FN F, int ntype, int n, double (&sf)[5]) {
  const bool phys = (ntype == 1 || ntype == 4);
  if (OP == OP_DERIV) {
    if (phys) {  // src/derivative.F90:230-248
      sf[0] = -2.5 * F(0) + 2.0 * F(1) + 0.5 * F(2);
      sf[1] = 0.75 * (F(2) - F(0));
    } else {     // :250-260, row of ghost node -1
      sf[0] = 0.75 * (F(0) - F(-2)) - 0.15 * (F(1) - F(-3)) + (1.0 / 60.0) * (F(2) - F(-4));
    }
  } else {
    if (phys) {  // src/filter.F90:176-204
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 6; ++j) v = v + c_fc.coefb[k][j] * F(j);
        sf[k] = v;
      }
      {
        const double v0 = F(3) + F(3), v1 = F(4) + F(2), v2 = F(5) + F(1), v3 = F(6) + F(0);
        sf[3] = c_fc.coef6i[0] * v0 + c_fc.coef6i[1] * v1 + c_fc.coef6i[2] * v2 + c_fc.coef6i[3] * v3;
      }
      {
        const double v0 = F(4) + F(4), v1 = F(5) + F(3), v2 = F(6) + F(2), v3 = F(7) + F(1),
                     v4 = F(8) + F(0);
        sf[4] = c_fc.coef8i[0] * v0 + c_fc.coef8i[1] * v1 + c_fc.coef8i[2] * v2 +
                c_fc.coef8i[3] * v3 + c_fc.coef8i[4] * v4;
      }
    } else {     // :206-218, ghost rows -3..-1 against the fixed window f(-5..5)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 10; ++j) v = v + c_fc.coefh[k][j] * F(-5 + j);
        sf[k] = v;
      }
    }
  }
}

template <int OP, class FN>
__device__ __forceinline__ void closure_last(FN F, int ntype, int n, double (&sl)[5]) {
  const bool phys = (ntype == 2 || ntype == 4);
  if (OP == OP_DERIV) {
    if (phys) {  // src/derivative.F90:264-281
      sl[0] = 0.75 * (F(n) - F(n - 2));
      sl[1] = 2.5 * F(n) - 2.0 * F(n - 1) - 0.5 * F(n - 2);
    } else {     // :283-292, row of ghost node n+1
      const int j = n + 1;
      sl[0] = 0.75 * (F(j + 1) - F(j - 1)) - 0.15 * (F(j + 2) - F(j - 2)) +
              (1.0 / 60.0) * (F(j + 3) - F(j - 3));
    }
  } else {
    if (phys) {  // src/filter.F90:222-249 ; rows n-4, n-3, n-2, n-1, n
      {
        const int j = n - 4;
        const double v0 = F(j) + F(j), v1 = F(j + 1) + F(j - 1), v2 = F(j + 2) + F(j - 2),
                     v3 = F(j + 3) + F(j - 3), v4 = F(j + 4) + F(j - 4);
        sl[0] = c_fc.coef8i[0] * v0 + c_fc.coef8i[1] * v1 + c_fc.coef8i[2] * v2 +
                c_fc.coef8i[3] * v3 + c_fc.coef8i[4] * v4;
      }
      {
        const int j = n - 3;
        const double v0 = F(j) + F(j), v1 = F(j + 1) + F(j - 1), v2 = F(j + 2) + F(j - 2),
                     v3 = F(j + 3) + F(j - 3);
        sl[1] = c_fc.coef6i[0] * v0 + c_fc.coef6i[1] * v1 + c_fc.coef6i[2] * v2 + c_fc.coef6i[3] * v3;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // node n-k uses coefb[k]
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 6; ++j) v = v + c_fc.coefb[k][j] * F(n - j);
        sl[4 - k] = v;
      }
    } else {     // :251-261 ; ghost rows n+1..n+3 ; node n+3-k uses coefh[k]
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 10; ++j) v = v + c_fc.coefh[k][j] * F(n + 5 - j);
        sl[2 - k] = v;
      }
    }
  }
}

template <int DIR, int OP, int C>
__global__ void __launch_bounds__(ASTR_BW* C) sweep_kernel(const SweepArgs a) {
  constexpr int H = OpTraits<OP>::H, W = OpTraits<OP>::W;
  constexpr int T = ASTR_BW * C;
  extern __shared__ __align__(16) double sm[];

  const int tid = threadIdx.x;
  const int n = a.op.n;
  const int nrows = a.op.nrows;
  const int first_node = a.op.first_node;
  const int sp = a.sp;
  const int ntile = (DIR == 0) ? ASTR_BW * sp : (n + 11) * ASTR_BW;
  double* EE = sm + ntile;
  double* GS = EE + C * ASTR_BW;
  double* XIN = GS + C * ASTR_BW;
  unsigned char* ch = reinterpret_cast<unsigned char*>(XIN + C * ASTR_BW);

  // ---- thread -> (pencil, chunk) ----------------------------------------------------
  int p, c;
  if (DIR == 0 && C > 1) {
    // half-warp = 8 pencils x 2 chunks: with sp == 2 (mod 16) and chunk starts of
    // different parity the 16 addresses of a half-warp fall in distinct 8-byte banks.
    p = (tid & 7) + 8 * ((tid >> 4) & 1);
    c = ((tid >> 3) & 1) + 2 * (tid >> 5);
  } else {
    p = tid & (ASTR_BW - 1);
    c = tid >> 4;
  }
  constexpr int SL = (DIR == 0) ? 1 : ASTR_BW;
  double* tile = sm + ((DIR == 0) ? (p * sp + 1) : p);  // tile[(node+5)*SL] = f(node)
  auto F = [&](int node) -> double { return tile[(node + 5) * SL]; };

  // ---- where this bundle lives in global memory -------------------------------------
  const Layout& L = a.L;
  const double* __restrict__ gin = a.in[blockIdx.z];
  double* __restrict__ gout = a.out[blockIdx.z];
  long long gbase, gl;   // element offset of (node 0, pencil 0); stride of a line step
  int pmax;              // last valid pencil index inside the bundle
  if (DIR == 0) {
    const int j0 = blockIdx.x * ASTR_BW, k = blockIdx.y;
    gbase = L.idx(0, j0, k); gl = 1; pmax = L.jm - j0;
  } else if (DIR == 1) {
    const int i0 = blockIdx.x * ASTR_BW, k = blockIdx.y;
    gbase = L.idx(i0, 0, k); gl = L.sj; pmax = L.im - i0;
  } else {
    const int i0 = blockIdx.x * ASTR_BW, j = blockIdx.y;
    gbase = L.idx(i0, j, 0); gl = L.sk; pmax = L.im - i0;
  }

  // ---- stage the bundle: global -> shared, 16 bytes per cp.async ---------------------
  if (DIR == 0) {
    constexpr int LW = (T < 32) ? T : 32;
    const int lane = tid % LW, wp = tid / LW;
    const int w2 = (n + 13) >> 1;  // nodes -6 .. n+5 (+1 pad when n is odd)
    for (int pp = wp; pp < ASTR_BW; pp += T / LW) {
      if (pp > pmax) continue;
      const double* src = gin + gbase + L.sj * pp - 6;
      double* dst = sm + pp * sp;
      for (int w = lane; w < w2; w += LW) cp_async16(dst + 2 * w, src + 2 * w);
    }
  } else {
    const double* src0 = gin + gbase - 5 * gl;
    const int tot = (n + 11) * 8;
    for (int e = tid; e < tot; e += T) {
      const int s = e >> 3, q = e & 7;
      cp_async16(sm + s * ASTR_BW + 2 * q, src0 + (long long)s * gl + 2 * q);
    }
  }
  for (int r = tid; r < nrows; r += T) ch[r] = (unsigned char)(((r + 1) * C - 1) / nrows);
  cp_async_wait_all();
  __syncthreads();

  // ---- phase 0: everything that must see the pristine line ---------------------------
  const int ra = (c * nrows) / C, rb = ((c + 1) * nrows) / C - 1;
  const int nsf = a.op.nsf, nsl = a.op.nsl;
  const int ri = (c == 0) ? nsf : ra;                    // first interior row of the chunk
  const int re = (c == C - 1) ? (nrows - 1 - nsl) : rb;  // last interior row of the chunk
  double w[W];
  {
    const int m0 = first_node + ri;
#pragma unroll
    for (int s = 0; s < 2 * H; ++s) w[s] = F(m0 - H + s);
    w[2 * H] = 0.0;
  }
  double sf[5] = {0, 0, 0, 0, 0}, sl[5] = {0, 0, 0, 0, 0};
  if (c == 0) closure_first<OP>(F, a.op.ntype, n, sf);
  if (c == C - 1) closure_last<OP>(F, a.op.ntype, n, sl);
  __syncthreads();

  const double* __restrict__ ac1 = a.op.ac1;
  const double* __restrict__ ac2 = a.op.ac2;
  const double* __restrict__ ac3 = a.op.ac3;
  const double* __restrict__ pf = a.op.pf;
  const double* __restrict__ qb = a.op.qb;

  // ---- phase 1: forward elimination with zero carry-in -------------------------------
  // e(r) = d(r)*ac2(r) - e(r-1)*ac3(r)        (src/commfunc.F90:802-804)
  double eprev = 0.0;
  double eh[H];
  {
    int r = ra;
    if (c == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < nsf) {
          const double e = sf[k] * __ldg(ac2 + r) - eprev * __ldg(ac3 + r);
          tile[(first_node + r + 5) * SL] = e;
          eprev = e;
          ++r;
        }
    }
    // the first H interior rows are kept in registers until every thread is done with
    // its right-hand overlap (those positions are the neighbour chunk's first nodes)
#pragma unroll
    for (int t = 0; t < H; ++t) {
      const int node = first_node + r;
      w[(t + 2 * H) % W] = F(node + H);
      const double d = interior_rhs<OP>(w, t);
      const double e = d * __ldg(ac2 + r) - eprev * __ldg(ac3 + r);
      eh[t] = e;
      eprev = e;
      ++r;
    }
    while (r <= re) {
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const int t = H + u;
        if (r <= re) {
          const int node = first_node + r;
          w[(t + 2 * H) % W] = F(node + H);
          const double d = interior_rhs<OP>(w, t);
          const double e = d * __ldg(ac2 + r) - eprev * __ldg(ac3 + r);
          tile[(node + 5) * SL] = e;
          eprev = e;
          ++r;
        }
      }
    }
    if (c == C - 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < nsl) {
          const double e = sl[k] * __ldg(ac2 + r) - eprev * __ldg(ac3 + r);
          tile[(first_node + r + 5) * SL] = e;
          eprev = e;
          ++r;
        }
    }
  }
  EE[c * ASTR_BW + p] = eprev;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < H; ++t) tile[(first_node + ri + t + 5) * SL] = eh[t];
  // true carry into this chunk: d'(ra-1)
  double cin = 0.0;
  for (int cc = 0; cc < c; ++cc) {
    const int rbc = ((cc + 1) * nrows) / C - 1;
    cin = EE[cc * ASTR_BW + p] + __ldg(pf + rbc) * cin;
  }

  // ---- phase 2: back substitution with zero carry-in ---------------------------------
  // x(r) = d'(r) - ac1(r)*x(r+1)              (src/commfunc.F90:808-810)
  {
    double gnext = 0.0;
#pragma unroll 4
    for (int r = rb; r >= ra; --r) {
      double* ps = tile + (first_node + r + 5) * SL;
      const double dp = *ps + __ldg(pf + r) * cin;
      const double g = dp - __ldg(ac1 + r) * gnext;
      *ps = g;
      gnext = g;
    }
    GS[c * ASTR_BW + p] = gnext;
  }
  __syncthreads();
  {
    double xin = 0.0;
    for (int cc = C - 1; cc > c; --cc) {
      const int rac = (cc * nrows) / C;
      xin = GS[cc * ASTR_BW + p] + __ldg(qb + rac) * xin;
    }
    XIN[c * ASTR_BW + p] = xin;
  }
  __syncthreads();

  // ---- phase 3: coalesced write-out, x = g + qb*xin, epilogue -------------------------
  const int epi = a.epi;
  const int o_lo = a.o_lo, o_hi = a.o_hi;
  const int w_lo = (epi == EPI_STOREZ) ? 0 : o_lo;
  const int w_hi = (epi == EPI_STOREZ) ? n : o_hi;
  if (DIR == 0) {
    constexpr int LW = (T < 32) ? T : 32;
    const int lane = tid % LW, wp = tid / LW;
    for (int pp = wp; pp < ASTR_BW; pp += T / LW) {
      if (pp > pmax) continue;
      double* orow = gout + gbase + L.sj * pp;
      const double* srow = sm + pp * sp + 6;
#pragma unroll 4
      for (int node = w_lo + lane; node <= w_hi; node += LW) {
        const int r = node - first_node;
        double x = srow[node] + __ldg(qb + r) * XIN[ch[r] * ASTR_BW + pp];
        if (epi == EPI_ADD) x = orow[node] + x;
        else if (epi == EPI_STOREZ && (node < o_lo || node > o_hi)) x = 0.0;
        orow[node] = x;
      }
    }
  } else {
    const int pp = tid & (ASTR_BW - 1);
    if (pp <= pmax) {
      double* ocol = gout + gbase + pp;
#pragma unroll 4
      for (int node = w_lo + (tid >> 4); node <= w_hi; node += C) {
        const int r = node - first_node;
        double x = sm[(node + 5) * ASTR_BW + pp] + __ldg(qb + r) * XIN[ch[r] * ASTR_BW + pp];
        double* po = ocol + (long long)node * gl;
        if (epi == EPI_ADD) x = *po + x;
        else if (epi == EPI_STOREZ && (node < o_lo || node > o_hi)) x = 0.0;
        *po = x;
      }
    }
  }
}

template <int DIR, int OP, int C>
int launch_one(const SweepArgs& a, cudaStream_t st) {
  int sp = 0;
  const size_t smem = astr_sweep_smem_bytes(DIR, a.op.n, C, &sp);
  static bool attr_done = false;
  static size_t attr_smem = 0;
  if (!attr_done || smem > attr_smem) {
    CUDA_OK(cudaFuncSetAttribute(sweep_kernel<DIR, OP, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    CUDA_OK(cudaFuncSetAttribute(sweep_kernel<DIR, OP, C>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    attr_done = true;
    attr_smem = smem;
  }
  SweepArgs b = a;
  b.sp = sp;
  const Layout& L = a.L;
  dim3 grid;
  if (DIR == 0) grid = dim3((L.jm + ASTR_BW) / ASTR_BW, L.km + 1, a.nf);
  else if (DIR == 1) grid = dim3((L.im + ASTR_BW) / ASTR_BW, L.km + 1, a.nf);
  else grid = dim3((L.im + ASTR_BW) / ASTR_BW, L.jm + 1, a.nf);
  sweep_kernel<DIR, OP, C><<<grid, ASTR_BW * C, smem, st>>>(b);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <int DIR, int OP>
int launch_c(const SweepArgs& a, cudaStream_t st) {
  switch (a.op.C) {
    case 1: return launch_one<DIR, OP, 1>(a, st);
    case 2: return launch_one<DIR, OP, 2>(a, st);
    case 4: return launch_one<DIR, OP, 4>(a, st);
    case 8: return launch_one<DIR, OP, 8>(a, st);
    default: return astr_fail_msg("sweep: unsupported chunk count");
  }
}

}  // namespace

// Largest supported chunk count whose chunks are all >= 12 rows (the closure rows and the
// H-row register prologue must fit in the first / last chunk).
int astr_sweep_max_chunks(int nrows) {
  int best = 0;
  const int cand[4] = {1, 2, 4, 8};
  for (int k = 0; k < 4; ++k)
    if (nrows / cand[k] >= 12) best = cand[k];
  return best;
}

size_t astr_sweep_smem_bytes(int dir, int n, int C, int* sp_out) {
  int sp = 0;
  size_t tile;
  if (dir == 0) {
    int wdt = n + 12 + (n & 1);  // nodes -6..n+5, even count
    sp = wdt;
    while ((sp & 15) != 2) ++sp;
    tile = (size_t)ASTR_BW * sp;
  } else {
    tile = (size_t)(n + 11) * ASTR_BW;
  }
  if (sp_out) *sp_out = sp;
  size_t bytes = (tile + 3 * (size_t)C * ASTR_BW) * sizeof(double) + (size_t)(n + 16);
  return (bytes + 15) & ~(size_t)15;
}

int astr_launch_sweep(int dir, int optype, const SweepArgs& a, cudaStream_t st) {
  if (a.nf < 1 || a.nf > ASTR_MAXF) return astr_fail_msg("sweep: bad field count");
  if (optype == OP_DERIV) {
    if (dir == 0) return launch_c<0, OP_DERIV>(a, st);
    if (dir == 1) return launch_c<1, OP_DERIV>(a, st);
    return launch_c<2, OP_DERIV>(a, st);
  } else {
    if (dir == 0) return launch_c<0, OP_FILTER>(a, st);
    if (dir == 1) return launch_c<1, OP_FILTER>(a, st);
    return launch_c<2, OP_FILTER>(a, st);
  }
}
