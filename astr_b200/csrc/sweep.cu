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
// The sequential dependence is therefore n/C long instead of n.  The kernel is persistent
// (one CTA per SM) with up to 3 independent 128-thread groups, each cycling through
// load -> solve -> store on its own bundle buffer, so the phases overlap inside the SM.
#include "common.cuh"
#include <cstdio>

__constant__ FilterCoef c_fc;

int astr_set_filter_coef(const FilterCoef& fc) {
  cudaError_t e = cudaMemcpyToSymbol(c_fc, &fc, sizeof(FilterCoef));
  if (e != cudaSuccess) return astr_fail("cudaMemcpyToSymbol(c_fc)", e, __FILE__, __LINE__);
  return 0;
}

// compact_flux_rhs interior coefficients (src/flux.F90:247-262) for the current bfacmpld
__constant__ double c_flx[4];
int astr_set_flux_coef(double b) {
  const double h[4] = {1.0 / 18.0 - (1.0 / 36.0) * b, 19.0 / 18.0 - (9.0 / 36.0) * b, 5.0 / 9.0 + (9.0 / 36.0) * b,
                       (1.0 / 36.0) * b};
  cudaError_t e = cudaMemcpyToSymbol(c_flx, h, sizeof h);
  if (e != cudaSuccess) return astr_fail("cudaMemcpyToSymbol(c_flx)", e, __FILE__, __LINE__);
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
template <> struct OpTraits<OP_FLUXP> { static constexpr int H = 2, W = 5; };
template <> struct OpTraits<OP_FLUXM> { static constexpr int H = 2, W = 5; };

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
  } else if (OP == OP_FLUXP) {
    // src/flux.F90:247-253
    return c_flx[0] * WS(-1) + c_flx[1] * WS(0) + c_flx[2] * WS(1) + c_flx[3] * WS(2);
  } else if (OP == OP_FLUXM) {
    // src/flux.F90:255-261
    return c_flx[0] * WS(2) + c_flx[1] * WS(1) + c_flx[2] * WS(0) + c_flx[3] * WS(-1);
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
  if (OP == OP_FLUXP || OP == OP_FLUXM) {
    if (phys) {  // src/flux.F90:189-199, rows -1 and 0
      sf[0] = 2.5 * F(0) + 0.5 * F(1);
      sf[1] = 0.75 * F(0) + 0.75 * F(1);
    } else {     // :205-209, row -2: 6th-order explicit
      const double var1 = F(-2) + F(-1), var2 = F(-3) + F(0), var3 = F(-4) + F(1);
      sf[0] = (37.0 / 60.0) * var1 - (2.0 / 15.0) * var2 + (1.0 / 60.0) * var3;
    }
  } else if (OP == OP_DERIV) {
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
  if (OP == OP_FLUXP || OP == OP_FLUXM) {
    if (phys) {  // src/flux.F90:217-225, rows n-1 and n
      sl[0] = 0.75 * F(n) + 0.75 * F(n - 1);
      sl[1] = 2.5 * F(n) + 0.5 * F(n - 1);
    } else {     // :231-236, row n+1
      const int j = n + 1;
      const double var1 = F(j) + F(j + 1), var2 = F(j - 1) + F(j + 2), var3 = F(j - 2) + F(j + 3);
      sl[0] = (37.0 / 60.0) * var1 - (2.0 / 15.0) * var2 + (1.0 / 60.0) * var3;
    }
  } else if (OP == OP_DERIV) {
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

template <int NG>
__device__ __forceinline__ void group_sync(int id, int nthreads) {
  if (NG == 1) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent kernel: one CTA per SM slot, NG independent groups of 16*C threads.  Every group
// owns one bundle buffer and walks the bundle list with stride gridDim.x*NG, so that while one
// group runs its recurrences another waits for its cp.async stage and a third streams its
// results out: the load / solve / store phases of different bundles overlap inside the SM.
// The operator tables (ac1, ac2, ac3, pf, qb) are staged in shared memory once per CTA.
template <int DIR, int OP, int C, int NG>
__global__ void __launch_bounds__(ASTR_BW* C* NG, 1) sweep_kernel(const SweepArgs a) {
  constexpr int H = OpTraits<OP>::H, W = OpTraits<OP>::W;
  constexpr int T = ASTR_BW * C;  // threads per group
  extern __shared__ __align__(16) double sm_all[];

  const int n = a.op.n;
  const int nrows = a.op.nrows;
  const int first_node = a.op.first_node;
  const int sp = a.sp;
  const int ntile = (DIR == 0) ? ASTR_BW * sp : (n + 11) * ASTR_BW;
  const int nr8 = (nrows + 1) & ~1;
  // shared memory: 5 tables | chunk-of-row bytes | NG x (tile, EE, GS, XIN)
  const double* __restrict__ ac1 = sm_all;
  const double* __restrict__ ac2 = sm_all + nr8;
  const double* __restrict__ ac3 = sm_all + 2 * nr8;
  const double* __restrict__ pf = sm_all + 3 * nr8;
  const double* __restrict__ qb = sm_all + 4 * nr8;
  unsigned char* ch = reinterpret_cast<unsigned char*>(sm_all + 5 * nr8);
  const int chd = ((nrows + 15) & ~15) / 8;  // doubles occupied by ch
  const int grp = threadIdx.x / T;
  const int tid = threadIdx.x - grp * T;
  double* sm = sm_all + 5 * nr8 + chd + (size_t)grp * (ntile + 2 * C * ASTR_BW);
  double* EE = sm + ntile;
  double* GS = EE + C * ASTR_BW;
  double* XIN = EE;  // EE is dead once every thread has passed the barrier after phase 2

  {
    double* tab = sm_all;
    for (int r = threadIdx.x; r < nrows; r += T * NG) {
      tab[r] = a.op.ac1[r];
      tab[nr8 + r] = a.op.ac2[r];
      tab[2 * nr8 + r] = a.op.ac3[r];
      tab[3 * nr8 + r] = a.op.pf[r];
      tab[4 * nr8 + r] = a.op.qb[r];
      {
        int cc = 0;
        while (cc + 1 < C && r >= chunk_start(cc + 1, nrows, C)) ++cc;
        ch[r] = (unsigned char)cc;
      }
    }
    // padding rows (the paired write-out may look one row past the end)
    for (int r = nrows + threadIdx.x; r < ((nrows + 15) & ~15); r += T * NG) ch[r] = 0;
  }
  __syncthreads();

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

  const Layout& L = a.L;
  const int nbx = ((DIR == 0 ? L.jm : L.im) + ASTR_BW) / ASTR_BW;
  const int nby = (DIR == 2 ? L.jm : L.km) + 1;
  const int nbundles = nbx * nby * a.nf;

  const int ra = chunk_start(c, nrows, C), rb = chunk_start(c + 1, nrows, C) - 1;
  const int nsf = a.op.nsf, nsl = a.op.nsl;
  const int ri = (c == 0) ? nsf : ra;                    // first interior row of the chunk
  const int re = (c == C - 1) ? (nrows - 1 - nsl) : rb;  // last interior row of the chunk
  const int epi = a.epi;
  const int o_lo = a.o_lo, o_hi = a.o_hi;
  const int w_lo = (epi == EPI_STOREZ) ? 0 : o_lo;
  const int w_hi = (epi == EPI_STOREZ) ? n : o_hi;

  for (int bnd = blockIdx.x * NG + grp; bnd < nbundles; bnd += gridDim.x * NG) {
    // ---- where this bundle lives in global memory -----------------------------------
    const int bx = bnd % nbx, by = (bnd / nbx) % nby, bz = bnd / (nbx * nby);
    const double* __restrict__ gin = a.in[bz];
    double* __restrict__ gout = a.out[bz];
    long long gbase, gl;   // element offset of (node 0, pencil 0); stride of a line step
    int pmax;              // last valid pencil index inside the bundle
    if (DIR == 0) {
      gbase = L.idx(0, bx * ASTR_BW, by); gl = 1; pmax = L.jm - bx * ASTR_BW;
    } else if (DIR == 1) {
      gbase = L.idx(bx * ASTR_BW, 0, by); gl = L.sj; pmax = L.im - bx * ASTR_BW;
    } else {
      gbase = L.idx(bx * ASTR_BW, by, 0); gl = L.sk; pmax = L.im - bx * ASTR_BW;
    }

    // ---- stage the bundle: global -> shared, 16 bytes per cp.async -------------------
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
      // thread -> (16-byte column q, row s0 + k*RS): a warp fetches 4 full 128-byte rows
      constexpr int RS = T / 8;
      const int q = tid & 7, s0 = tid >> 3;
      const double* src = gin + gbase + (long long)(s0 - 5) * gl + 2 * q;
      double* dst = sm + s0 * ASTR_BW + 2 * q;
      const long long sstep = (long long)RS * gl;
#pragma unroll 4
      for (int srow = s0; srow < n + 11; srow += RS) {
        cp_async16(dst, src);
        dst += RS * ASTR_BW;
        src += sstep;
      }
    }
    cp_async_wait_all();
    group_sync<NG>(grp + 1, T);

    // ---- phase 0: everything that must see the pristine line -------------------------
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
    group_sync<NG>(grp + 1, T);

    // ---- phase 1: forward elimination with zero carry-in -----------------------------
    // e(r) = d(r)*ac2(r) - e(r-1)*ac3(r)        (src/commfunc.F90:802-804)
    double eprev = 0.0;
    double eh[H];
    {
      int r = ra;
      if (c == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k < nsf) {
            const double e = __fma_rn(-eprev, ac3[r], sf[k] * ac2[r]);
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
        const double e = __fma_rn(-eprev, ac3[r], d * ac2[r]);
        eh[t] = e;
        eprev = e;
        ++r;
      }
      // full groups of W rows without guards: the loads of a group are independent of the
      // recurrence and get hoisted, only the FMA chain is sequential
      const int nfull = (re - r + 1) / W;
      for (int gi = 0; gi < nfull; ++gi) {
        double* ps = tile + (first_node + r + 5) * SL;
        // every load of the group first (explicitly: the compiler cannot move shared loads
        // above the in-place stores), then the arithmetic; only one FMA per row is sequential
        double fn[W], c2[W], c3[W];
#pragma unroll
        for (int u = 0; u < W; ++u) {
          fn[u] = ps[(u + H) * SL];
          c2[u] = ac2[r + u];
          c3[u] = ac3[r + u];
        }
#pragma unroll
        for (int u = 0; u < W; ++u) {
          const int t = H + u;
          w[(t + 2 * H) % W] = fn[u];
          const double d = interior_rhs<OP>(w, t);
          const double e = __fma_rn(-eprev, c3[u], d * c2[u]);
          ps[u * SL] = e;
          eprev = e;
        }
        r += W;
      }
#pragma unroll
      for (int u = 0; u < W - 1; ++u) {
        const int t = H + u;
        if (r <= re) {
          const int node = first_node + r;
          w[(t + 2 * H) % W] = F(node + H);
          const double d = interior_rhs<OP>(w, t);
          const double e = __fma_rn(-eprev, ac3[r], d * ac2[r]);
          tile[(node + 5) * SL] = e;
          eprev = e;
          ++r;
        }
      }
      if (c == C - 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k < nsl) {
            const double e = __fma_rn(-eprev, ac3[r], sl[k] * ac2[r]);
            tile[(first_node + r + 5) * SL] = e;
            eprev = e;
            ++r;
          }
      }
    }
    EE[c * ASTR_BW + p] = eprev;
    group_sync<NG>(grp + 1, T);
#pragma unroll
    for (int t = 0; t < H; ++t) tile[(first_node + ri + t + 5) * SL] = eh[t];
    // true carry into this chunk: d'(ra-1)
    double cin = 0.0;
    for (int cc = 0; cc < c; ++cc) {
      const int rbc = chunk_start(cc + 1, nrows, C) - 1;
      cin = EE[cc * ASTR_BW + p] + pf[rbc] * cin;
    }

    // ---- phase 2: back substitution with zero carry-in -------------------------------
    // x(r) = d'(r) - ac1(r)*x(r+1)              (src/commfunc.F90:808-810)
    {
      double gnext = 0.0;
      int r = rb;
      const int nfull = (rb - ra + 1) / 8;
      for (int gi = 0; gi < nfull; ++gi) {
        double* ps = tile + (first_node + r + 5) * SL;
        double ev[8], pv[8], av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ev[u] = ps[-u * SL];
          pv[u] = pf[r - u];
          av[u] = ac1[r - u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double dp = __fma_rn(pv[u], cin, ev[u]);
          const double g = __fma_rn(-av[u], gnext, dp);
          ps[-u * SL] = g;
          gnext = g;
        }
        r -= 8;
      }
      for (; r >= ra; --r) {
        double* ps = tile + (first_node + r + 5) * SL;
        const double dp = __fma_rn(pf[r], cin, *ps);
        const double g = __fma_rn(-ac1[r], gnext, dp);
        *ps = g;
        gnext = g;
      }
      GS[c * ASTR_BW + p] = gnext;
    }
    group_sync<NG>(grp + 1, T);
    {
      double xin = 0.0;
      for (int cc = C - 1; cc > c; --cc) {
        const int rac = chunk_start(cc, nrows, C);
        xin = GS[cc * ASTR_BW + p] + qb[rac] * xin;
      }
      XIN[c * ASTR_BW + p] = xin;
    }
    group_sync<NG>(grp + 1, T);

    // ---- phase 3: coalesced write-out, x = g + qb*xin, epilogue ----------------------
    // chunk by chunk (xin is per chunk), 4 rows per batch with the accumulate loads up front
    const int r_lo = w_lo - first_node, r_hi = w_hi - first_node;
    if (DIR == 0) {
      // lanes run along the line in aligned node pairs (16-byte accesses, full 128-byte lines);
      // the chunk of each node comes from the byte table
      constexpr int LW = (T < 32) ? T : 32;
      const int lane = tid % LW, wp = tid / LW;
      const int n_lo = w_lo & ~1;
      for (int pp = wp; pp < ASTR_BW; pp += T / LW) {
        if (pp > pmax) continue;
        double* orow = gout + gbase + L.sj * pp;                   // orow[node]
        const double* srow = sm + pp * sp + 6;                     // srow[node]
        const double* xrow = XIN + pp;
        for (int nd0 = n_lo + 2 * lane; nd0 <= w_hi; nd0 += 8 * LW) {
          double2 old[4];
          if (epi == EPI_ADD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int nd = nd0 + 2 * k * LW;
              if (nd <= w_hi) {
                if (nd >= w_lo && nd + 1 <= w_hi) old[k] = *reinterpret_cast<const double2*>(orow + nd);
                else { old[k].x = (nd >= w_lo) ? orow[nd] : 0.0; old[k].y = (nd + 1 <= w_hi) ? orow[nd + 1] : 0.0; }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int nd = nd0 + 2 * k * LW;
            if (nd <= w_hi) {
              const int r = nd - first_node;
              const int rx = max(r, 0);   // nd may sit one below the first row (its half is not stored)
              const double2 gv = *reinterpret_cast<const double2*>(srow + nd);
              double2 x;
              x.x = __fma_rn(qb[rx], xrow[ch[rx] * ASTR_BW], gv.x);
              x.y = __fma_rn(qb[r + 1], xrow[ch[r + 1] * ASTR_BW], gv.y);
              if (epi == EPI_ADD) { x.x = old[k].x + x.x; x.y = old[k].y + x.y; }
              else if (epi == EPI_STOREZ) {
                if (nd < o_lo || nd > o_hi) x.x = 0.0;
                if (nd + 1 < o_lo || nd + 1 > o_hi) x.y = 0.0;
              }
              if (nd >= w_lo && nd + 1 <= w_hi) *reinterpret_cast<double2*>(orow + nd) = x;
              else if (nd >= w_lo) orow[nd] = x.x;
              else if (nd + 1 <= w_hi) orow[nd + 1] = x.y;
            }
          }
        }
      }
    } else {
      // thread -> pencil pair (16-byte column) x row slot: a warp writes 4 full 128-byte rows
      constexpr int RS = T / 8;
      const int pp = 2 * (tid & 7), rs = tid >> 3;
      if (pp <= pmax) {
        const bool pair = (pp + 1 <= pmax);
        double* ocol = gout + gbase + pp + (long long)first_node * gl;   // ocol[r*gl]
        const double* scol = sm + (first_node + 5) * ASTR_BW + pp;        // scol[r*16]
        for (int cc = 0; cc < C; ++cc) {
          const double2 xin = *reinterpret_cast<const double2*>(XIN + cc * ASTR_BW + pp);
          const int rlo = max(chunk_start(cc, nrows, C), r_lo), rhi = min(chunk_start(cc + 1, nrows, C) - 1, r_hi);
          for (int r0 = rlo + rs; r0 <= rhi; r0 += 4 * RS) {
            double2 old[4];
            if (epi == EPI_ADD) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (r0 + k * RS <= rhi) {
                  const double* po = ocol + (long long)(r0 + k * RS) * gl;
                  if (pair) old[k] = *reinterpret_cast<const double2*>(po);
                  else { old[k].x = *po; old[k].y = 0.0; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = r0 + k * RS;
              if (r <= rhi) {
                const double2 gv = *reinterpret_cast<const double2*>(scol + r * ASTR_BW);
                const double qv = qb[r];
                double2 x;
                x.x = __fma_rn(qv, xin.x, gv.x);
                x.y = __fma_rn(qv, xin.y, gv.y);
                if (epi == EPI_ADD) { x.x = old[k].x + x.x; x.y = old[k].y + x.y; }
                else if (epi == EPI_STOREZ && (r < o_lo - first_node || r > o_hi - first_node)) { x.x = 0.0; x.y = 0.0; }
                double* po = ocol + (long long)r * gl;
                if (pair) *reinterpret_cast<double2*>(po) = x;
                else *po = x.x;
              }
            }
          }
        }
      }
    }
    // the buffer is overwritten by the next bundle's cp.async: all reads must be done
    group_sync<NG>(grp + 1, T);
  }
}

static int g_num_sms = 0;

template <int DIR, int OP, int C, int NG>
int launch_one(const SweepArgs& a, cudaStream_t st) {
  int sp = 0;
  const size_t smem = astr_sweep_smem_bytes(DIR, a.op.n, C, NG, &sp);
  static size_t attr_smem = 0;
  static int occ = 0;
  auto kern = sweep_kernel<DIR, OP, C, NG>;
  if (smem > attr_smem || occ == 0) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    attr_smem = smem;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ASTR_BW * C * NG, smem));
    if (occ < 1) return astr_fail_msg("sweep: kernel does not fit on an SM");
    if (!g_num_sms) {
      int dev = 0;
      CUDA_OK(cudaGetDevice(&dev));
      CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
  }
  SweepArgs b = a;
  b.sp = sp;
  const Layout& L = a.L;
  const int nbx = ((DIR == 0 ? L.jm : L.im) + ASTR_BW) / ASTR_BW;
  const int nby = (DIR == 2 ? L.jm : L.km) + 1;
  const long long nbundles = (long long)nbx * nby * a.nf;
  long long grid = (long long)g_num_sms * occ;
  const long long need = (nbundles + NG - 1) / NG;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, ASTR_BW * C * NG, smem, st>>>(b);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}


// -------------------------------------------------------------------------------------------------
// Warp-per-line engine for the i direction (the line is contiguous in memory).
// A warp owns one line: coalesced 16-byte loads bring nodes -6..n+5 into a per-warp shared line buffer,
// lane c then owns rows [chunk32_start(c), chunk32_start(c+1)) (<= 17) in REGISTERS: right-hand sides from
// a rotating window over the buffer (one LDS per row), forward and backward recurrences of the reference
// (src/commfunc.F90:790-813) with zero carries, the true carries from two warp-level scans of affine maps
// (5 shuffle steps each) instead of block barriers, solution back into the buffer and out with coalesced
// 16-byte stores.  No CTA-level synchronisation inside the line loop: warps are independent, so the SM
// hides latency with 16 warps of independent lines.  Same tables, closures and epilogues as above.
// -------------------------------------------------------------------------------------------------
constexpr int W3_WARPS = 8;           // warps per CTA
constexpr int W3_PADF = 8;            // buffer index = node + W3_PADF (window reads of closure rows may reach node -8)

template <int OP>
__global__ void __launch_bounds__(W3_WARPS * 32, 2) sweep3_kernel(const SweepArgs a) {
  constexpr int H = OpTraits<OP>::H, W = OpTraits<OP>::W;
  constexpr int LCH = ASTR_W3_LCH;
  extern __shared__ __align__(16) double sm_all[];
  const int n = a.op.n, nrows = a.op.nrows, first_node = a.op.first_node;
  const int nr8 = (nrows + 1) & ~1;
  const double* __restrict__ ac1 = sm_all;
  const double* __restrict__ ac2 = sm_all + nr8;
  const double* __restrict__ ac3 = sm_all + 2 * nr8;
  const double* __restrict__ pf = sm_all + 3 * nr8;
  const double* __restrict__ qb = sm_all + 4 * nr8;
  const int lbn = (n + 2 * W3_PADF + 2 + 1) & ~1;         // doubles per line buffer (even)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* LBa = sm_all + 5 * nr8 + (size_t)warp * 2 * lbn;   // two line buffers per warp: the next line is
  double* LBb = LBa + lbn;                                   // fetched (cp.async) under the solve of this one
  for (int r = threadIdx.x; r < nrows; r += W3_WARPS * 32) {
    sm_all[r] = a.op.ac1[r];
    sm_all[nr8 + r] = a.op.ac2[r];
    sm_all[2 * nr8 + r] = a.op.ac3[r];
    sm_all[3 * nr8 + r] = a.op.pf32[r];
    sm_all[4 * nr8 + r] = a.op.qb32[r];
  }
  // window reads of closure rows may touch the pads: keep them finite
  for (int t = lane; t < 2 * lbn; t += 32) LBa[t] = 0.0;
  __syncthreads();

  const Layout& L = a.L;
  const int ra = chunk32_start(lane, nrows), rb = chunk32_start(lane + 1, nrows) - 1;
  const int nsf = a.op.nsf, nsl = a.op.nsl;
  const int epi = a.epi, o_lo = a.o_lo, o_hi = a.o_hi;
  const int w_lo = (epi == EPI_STOREZ) ? 0 : o_lo;
  const int w_hi = (epi == EPI_STOREZ) ? n : o_hi;
  const long long nlines = (long long)a.nf * (L.km + 1) * (L.jm + 1);
  const long long wstride = (long long)gridDim.x * W3_WARPS;
  const int npair = (n + 13) >> 1;                         // node pairs -6.. (n+5 or n+6)
  double* LB = LBa;
  auto F = [&](int node) -> double { return LB[node + W3_PADF]; };
  auto row_of = [&](long long ln, int& f) -> long long {
    const int j = (int)(ln % (L.jm + 1));
    const int k = (int)((ln / (L.jm + 1)) % (L.km + 1));
    f = (int)(ln / ((long long)(L.jm + 1) * (L.km + 1)));
    return L.idx(0, j, k);
  };
  // line -> buffer: 16-byte cp.async (node -6 is 16-byte aligned in global memory; W3_PADF - 6 is even)
  auto fetch = [&](long long ln, double* buf) {
    int f;
    const long long off = row_of(ln, f);
    const double* __restrict__ grow = a.in[f] + off;
    for (int w = lane; w < npair; w += 32) cp_async16(buf + (W3_PADF - 6) + 2 * w, grow - 6 + 2 * w);
  };

  const long long ln0 = (long long)blockIdx.x * W3_WARPS + warp;
  if (ln0 < nlines) fetch(ln0, LB);
  for (long long ln = ln0; ln < nlines; ln += wstride) {
    int f;
    const long long off = row_of(ln, f);
    double* __restrict__ orow = a.out[f] + off;
    cp_async_wait_all();
    __syncwarp();
    if (ln + wstride < nlines) fetch(ln + wstride, LB == LBa ? LBb : LBa);   // in flight during the solve
    // ---- right-hand sides and forward elimination with zero carry-in ---------------------------------
    double sf[5] = {0, 0, 0, 0, 0}, sl[5] = {0, 0, 0, 0, 0};
    if (lane == 0) closure_first<OP>(F, a.op.ntype, n, sf);
    if (rb >= nrows - nsl && ra <= rb) closure_last<OP>(F, a.op.ntype, n, sl);   // lanes that own tail closure rows
    double e[LCH];
    double w[W];
    {
      const double* p = LB + W3_PADF + first_node + ra - H;
#pragma unroll
      for (int s = 0; s < 2 * H; ++s) w[s] = p[s];
      w[2 * H] = 0.0;
    }
    double eprev = 0.0;
#pragma unroll
    for (int s = 0; s < LCH; ++s) {
      const int r = ra + s;
      if (r <= rb) {
        w[(s + 2 * H) % W] = LB[W3_PADF + first_node + r + H];
        double d = interior_rhs<OP>(w, s);
        if (lane == 0 && s < nsf) d = sf[s < 5 ? s : 4];
        if (r >= nrows - nsl) {
          const int kk = r - (nrows - nsl);
          d = kk == 0 ? sl[0] : kk == 1 ? sl[1] : kk == 2 ? sl[2] : kk == 3 ? sl[3] : sl[4];
        }
        const double ev = __fma_rn(-eprev, ac3[r], d * ac2[r]);
        e[s] = ev;
        eprev = ev;
      } else {
        e[s] = 0.0;
      }
    }
    // ---- carry into each chunk: d'(ra-1) = EE(c-1) + pf(rb(c-1)) * d'(ra(c-1)-1)  (exclusive scan of x -> A x + B) --
    double cin;
    {
      double A = (ra <= rb) ? pf[rb] : 1.0, B = eprev;     // an empty chunk is the identity map
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double Ap = __shfl_up_sync(0xffffffffu, A, off), Bp = __shfl_up_sync(0xffffffffu, B, off);
        if (lane >= off) { B = __fma_rn(A, Bp, B); A = A * Ap; }
      }
      cin = __shfl_up_sync(0xffffffffu, B, 1);
      if (lane == 0) cin = 0.0;
    }
    // ---- back substitution with zero carry-in ---------------------------------------------------------
    double gnext = 0.0;
#pragma unroll
    for (int s = LCH - 1; s >= 0; --s) {
      const int r = ra + s;
      if (r <= rb) {
        const double dp = __fma_rn(pf[r], cin, e[s]);
        const double g = __fma_rn(-ac1[r], gnext, dp);
        e[s] = g;
        gnext = g;
      }
    }
    // ---- x(rb+1) of each chunk: x(ra(c)) = GS(c) + qb(ra(c)) * x(rb(c)+1)  (suffix scan from the last chunk) -----
    double xin;
    {
      double A = (ra <= rb) ? qb[ra] : 1.0, B = gnext;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double Ap = __shfl_down_sync(0xffffffffu, A, off), Bp = __shfl_down_sync(0xffffffffu, B, off);
        if (lane + off < 32) { B = __fma_rn(A, Bp, B); A = A * Ap; }
      }
      xin = __shfl_down_sync(0xffffffffu, B, 1);
      if (lane == 31) xin = 0.0;
    }
    __syncwarp();                      // every lane is done reading the pristine line
#pragma unroll
    for (int s = 0; s < LCH; ++s) {
      const int r = ra + s;
      if (r <= rb) LB[W3_PADF + first_node + r] = __fma_rn(qb[r], xin, e[s]);
    }
    __syncwarp();
    // ---- buffer -> global, aligned node pairs, epilogue ------------------------------------------------
    for (int nd = (w_lo & ~1) + 2 * lane; nd <= w_hi; nd += 64) {
      const bool vx = nd >= w_lo, vy = nd + 1 <= w_hi;
      double2 x = *reinterpret_cast<const double2*>(LB + W3_PADF + nd);
      if (epi == EPI_ADD) {
        if (vx) x.x = orow[nd] + x.x;
        if (vy) x.y = orow[nd + 1] + x.y;
      } else if (epi == EPI_STOREZ) {
        if (nd < o_lo || nd > o_hi) x.x = 0.0;
        if (nd + 1 < o_lo || nd + 1 > o_hi) x.y = 0.0;
      }
      if (vx && vy) *reinterpret_cast<double2*>(orow + nd) = x;
      else if (vx) orow[nd] = x.x;
      else if (vy) orow[nd + 1] = x.y;
    }
    __syncwarp();                      // this buffer is refilled two lines from now
    LB = (LB == LBa) ? LBb : LBa;
  }
}


// -------------------------------------------------------------------------------------------------
// EXPERIMENTAL (ASTR_SWEEP_W3=2; parity-tested, not yet timed): the warp-per-line engine with the overheads
// the SASS of sweep3_kernel shows removed.  The operator tables are padded to 32 x 17 rows with neutral rows (ac2 = 1,
// ac1 = ac3 = 0), so every lane runs the same 17 unguarded rows; per-lane base pointers turn all table and
// buffer accesses of the unrolled rows into immediate offsets; the few rows whose right-hand side is a tail
// closure (or padding) are recomputed afterwards by the lanes that own them.
// -------------------------------------------------------------------------------------------------
constexpr int W3P_LBN = 560;          // doubles per line buffer: node -8 .. first_node + 543 + H

template <int OP>
__global__ void __launch_bounds__(W3_WARPS * 32, 2) sweep3p_kernel(const SweepArgs a) {
  constexpr int H = OpTraits<OP>::H, W = OpTraits<OP>::W;
  constexpr int LCH = ASTR_W3_LCH, NR = ASTR_W3_ROWS;
  extern __shared__ __align__(16) double sm_all[];
  const int n = a.op.n, nrows = a.op.nrows, first_node = a.op.first_node;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* LBa = sm_all + 5 * NR + (size_t)warp * 2 * W3P_LBN;
  double* LBb = LBa + W3P_LBN;
  for (int r = threadIdx.x; r < 5 * NR; r += W3_WARPS * 32) sm_all[r] = a.op.pad5[r];
  for (int t = lane; t < 2 * W3P_LBN; t += 32) LBa[t] = 0.0;
  __syncthreads();
  // per-lane views of the tables: entry s of lane c is row 17 c + s
  const double* __restrict__ t1 = sm_all + LCH * lane;
  const double* __restrict__ t2 = sm_all + NR + LCH * lane;
  const double* __restrict__ t3 = sm_all + 2 * NR + LCH * lane;
  const double* __restrict__ tp = sm_all + 3 * NR + LCH * lane;
  const double* __restrict__ tq = sm_all + 4 * NR + LCH * lane;

  const Layout& L = a.L;
  const int nsf = a.op.nsf, nsl = a.op.nsl;
  const int epi = a.epi, o_lo = a.o_lo, o_hi = a.o_hi;
  const int w_lo = (epi == EPI_STOREZ) ? 0 : o_lo;
  const int w_hi = (epi == EPI_STOREZ) ? n : o_hi;
  const long long nlines = (long long)a.nf * (L.km + 1) * (L.jm + 1);
  const long long wstride = (long long)gridDim.x * W3_WARPS;
  const int npair = (n + 13) >> 1;
  // k = row - (nrows - nsl): k < 0 regular row, 0 <= k < nsl tail closure row, k >= nsl padding
  const int k0 = LCH * lane - (nrows - nsl);
  const bool fix_tail = k0 + LCH - 1 >= 0;
  double* LB = LBa;
  auto F = [&](int node) -> double { return LB[node + W3_PADF]; };
  auto row_of = [&](long long ln, int& f) -> long long {
    const int j = (int)(ln % (L.jm + 1));
    const int k = (int)((ln / (L.jm + 1)) % (L.km + 1));
    f = (int)(ln / ((long long)(L.jm + 1) * (L.km + 1)));
    return L.idx(0, j, k);
  };
  auto fetch = [&](long long ln, double* buf) {
    int f;
    const long long off = row_of(ln, f);
    const double* __restrict__ grow = a.in[f] + off;
    for (int w = lane; w < npair; w += 32) cp_async16(buf + (W3_PADF - 6) + 2 * w, grow - 6 + 2 * w);
  };

  const long long ln0 = (long long)blockIdx.x * W3_WARPS + warp;
  if (ln0 < nlines) fetch(ln0, LB);
  for (long long ln = ln0; ln < nlines; ln += wstride) {
    int f;
    const long long off = row_of(ln, f);
    double* __restrict__ orow = a.out[f] + off;
    cp_async_wait_all();
    __syncwarp();
    if (ln + wstride < nlines) fetch(ln + wstride, LB == LBa ? LBb : LBa);
    double* __restrict__ lw = LB + W3_PADF + first_node + LCH * lane;      // lw[s] = f(node of row 17 lane + s)
    double sf[5] = {0, 0, 0, 0, 0}, sl[5] = {0, 0, 0, 0, 0};
    if (lane == 0) closure_first<OP>(F, a.op.ntype, n, sf);
    if (fix_tail && k0 < nsl) closure_last<OP>(F, a.op.ntype, n, sl);
    double e[LCH];
    double w[W];
#pragma unroll
    for (int s = 0; s < 2 * H; ++s) w[s] = lw[s - H];
    w[2 * H] = 0.0;
    double eprev = 0.0;
#pragma unroll
    for (int s = 0; s < LCH; ++s) {
      w[(s + 2 * H) % W] = lw[s + H];
      double d = interior_rhs<OP>(w, s);
      if (s < 5) { if (lane == 0 && s < nsf) d = sf[s]; }
      const double ev = __fma_rn(-eprev, t3[s], d * t2[s]);
      e[s] = ev;
      eprev = ev;
    }
    if (fix_tail) {      // the last lanes: tail closure rows and padding get their own right-hand sides
      double ep = 0.0;
#pragma unroll
      for (int s = 0; s < LCH; ++s) {
        const int k = k0 + s;
        if (k < 0) ep = e[s];
        else {
          const double d = k == 0 ? sl[0] : k == 1 ? sl[1] : k == 2 ? sl[2] : k == 3 ? sl[3] : k == 4 ? sl[4] : 0.0;
          ep = __fma_rn(-ep, t3[s], (k < nsl ? d : 0.0) * t2[s]);
          e[s] = ep;
        }
      }
      eprev = ep;
    }
    double cin;
    {
      double A = tp[LCH - 1], B = eprev;
#pragma unroll
      for (int off2 = 1; off2 < 32; off2 <<= 1) {
        const double Ap = __shfl_up_sync(0xffffffffu, A, off2), Bp = __shfl_up_sync(0xffffffffu, B, off2);
        if (lane >= off2) { B = __fma_rn(A, Bp, B); A = A * Ap; }
      }
      cin = __shfl_up_sync(0xffffffffu, B, 1);
      if (lane == 0) cin = 0.0;
    }
    double gnext = 0.0;
#pragma unroll
    for (int s = LCH - 1; s >= 0; --s) {
      const double dp = __fma_rn(tp[s], cin, e[s]);
      const double g = __fma_rn(-t1[s], gnext, dp);
      e[s] = g;
      gnext = g;
    }
    double xin;
    {
      double A = tq[0], B = gnext;
#pragma unroll
      for (int off2 = 1; off2 < 32; off2 <<= 1) {
        const double Ap = __shfl_down_sync(0xffffffffu, A, off2), Bp = __shfl_down_sync(0xffffffffu, B, off2);
        if (lane + off2 < 32) { B = __fma_rn(A, Bp, B); A = A * Ap; }
      }
      xin = __shfl_down_sync(0xffffffffu, B, 1);
      if (lane == 31) xin = 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < LCH; ++s) lw[s] = __fma_rn(tq[s], xin, e[s]);
    __syncwarp();
    for (int nd = (w_lo & ~1) + 2 * lane; nd <= w_hi; nd += 64) {
      const bool vx = nd >= w_lo, vy = nd + 1 <= w_hi;
      double2 x = *reinterpret_cast<const double2*>(LB + W3_PADF + nd);
      if (epi == EPI_ADD) {
        if (vx) x.x = orow[nd] + x.x;
        if (vy) x.y = orow[nd + 1] + x.y;
      } else if (epi == EPI_STOREZ) {
        if (nd < o_lo || nd > o_hi) x.x = 0.0;
        if (nd + 1 < o_lo || nd + 1 > o_hi) x.y = 0.0;
      }
      if (vx && vy) *reinterpret_cast<double2*>(orow + nd) = x;
      else if (vx) orow[nd] = x.x;
      else if (vy) orow[nd + 1] = x.y;
    }
    __syncwarp();
    LB = (LB == LBa) ? LBb : LBa;
  }
}

template <int OP, bool PADDED>
int launch3(const SweepArgs& a, cudaStream_t st) {
  auto kern = PADDED ? sweep3p_kernel<OP> : sweep3_kernel<OP>;
  const int nr8 = (a.op.nrows + 1) & ~1;
  const int lbn = (a.op.n + 2 * W3_PADF + 2 + 1) & ~1;
  const size_t smem = PADDED ? ((size_t)5 * ASTR_W3_ROWS + (size_t)W3_WARPS * 2 * W3P_LBN) * sizeof(double)
                             : ((size_t)5 * nr8 + (size_t)W3_WARPS * 2 * lbn) * sizeof(double);
  static size_t attr_smem = 0;
  static int occ = 0;
  if (smem > attr_smem || occ == 0) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_smem = smem;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W3_WARPS * 32, smem));
    if (occ < 1) return astr_fail_msg("sweep3: kernel does not fit on an SM");
    if (!g_num_sms) {
      int dev = 0;
      CUDA_OK(cudaGetDevice(&dev));
      CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
  }
  const Layout& L = a.L;
  const long long nlines = (long long)a.nf * (L.km + 1) * (L.jm + 1);
  long long grid = (long long)g_num_sms * occ;
  const long long need = (nlines + W3_WARPS - 1) / W3_WARPS;
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, W3_WARPS * 32, smem, st>>>(a);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <int DIR, int OP>
int launch_c(const SweepArgs& a, cudaStream_t st) {
  const int n = a.op.n;
  const size_t cap = 227 * 1024;
  switch (a.op.C) {
    case 1: return launch_one<DIR, OP, 1, 1>(a, st);
    case 2: return launch_one<DIR, OP, 2, 1>(a, st);
    case 4: return launch_one<DIR, OP, 4, 1>(a, st);
    case 8:
      if (astr_sweep_smem_bytes(DIR, n, 8, 3, nullptr) <= cap) return launch_one<DIR, OP, 8, 3>(a, st);
      if (astr_sweep_smem_bytes(DIR, n, 8, 2, nullptr) <= cap) return launch_one<DIR, OP, 8, 2>(a, st);
      if (astr_sweep_smem_bytes(DIR, n, 8, 1, nullptr) <= cap) return launch_one<DIR, OP, 8, 1>(a, st);
      return astr_fail_msg("sweep: line too long for shared memory");
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

size_t astr_sweep_smem_bytes(int dir, int n, int C, int NG, int* sp_out) {
  int sp = 0;
  size_t tile;
  if (dir == 0) {
    int wdt = n + 12 + (n & 1);  // nodes -6..n+5, even count
    sp = wdt;
    while ((sp & 15) != 2 && (sp & 15) != 14) ++sp;   // 2*p_lo*(sp/2) must hit 8 distinct even banks
    tile = (size_t)ASTR_BW * sp;
  } else {
    tile = (size_t)(n + 11) * ASTR_BW;
  }
  if (sp_out) *sp_out = sp;
  const size_t nrows_max = (size_t)n + 8;          // >= nrows rounded up to even
  const size_t tabs = 5 * nrows_max + ((nrows_max + 15) & ~(size_t)15) / 8;
  return (tabs + (size_t)NG * (tile + 2 * (size_t)C * ASTR_BW)) * sizeof(double);
}

// i direction only; lines of 8*32 .. 17*32 rows (every lane chunk then holds the H-row window and, at the two
// ends, the closure rows).  Returns -1 when it does not apply: the caller uses the shared-memory engine.
int astr_launch_sweep3(int optype, const SweepArgs& a, int variant, cudaStream_t st) {
  if (!a.op.pf32 || a.op.nrows < 8 * 32 || a.op.nrows > ASTR_W3_LCH * 32) return -1;   // < 256 rows: too few lanes busy
  if (a.nf < 1 || a.nf > ASTR_MAXF) return astr_fail_msg("sweep: bad field count");
  if (variant == 2 && a.op.pad5 && a.op.first_node >= -3 && a.op.n + 6 + W3_PADF < W3P_LBN) {   // experimental padded variant
    switch (optype) {
      case OP_DERIV: return launch3<OP_DERIV, true>(a, st);
      case OP_FILTER: return launch3<OP_FILTER, true>(a, st);
      case OP_FLUXP: return launch3<OP_FLUXP, true>(a, st);
      default: return launch3<OP_FLUXM, true>(a, st);
    }
  }
  switch (optype) {
    case OP_DERIV: return launch3<OP_DERIV, false>(a, st);
    case OP_FILTER: return launch3<OP_FILTER, false>(a, st);
    case OP_FLUXP: return launch3<OP_FLUXP, false>(a, st);
    default: return launch3<OP_FLUXM, false>(a, st);
  }
}

int astr_launch_sweep(int dir, int optype, const SweepArgs& a, cudaStream_t st) {
  if (a.nf < 1 || a.nf > ASTR_MAXF) return astr_fail_msg("sweep: bad field count");
  if (optype == OP_DERIV) {
    if (dir == 0) return launch_c<0, OP_DERIV>(a, st);
    if (dir == 1) return launch_c<1, OP_DERIV>(a, st);
    return launch_c<2, OP_DERIV>(a, st);
  } else if (optype == OP_FLUXP) {
    if (dir == 0) return launch_c<0, OP_FLUXP>(a, st);
    if (dir == 1) return launch_c<1, OP_FLUXP>(a, st);
    return launch_c<2, OP_FLUXP>(a, st);
  } else if (optype == OP_FLUXM) {
    if (dir == 0) return launch_c<0, OP_FLUXM>(a, st);
    if (dir == 1) return launch_c<1, OP_FLUXM>(a, st);
    return launch_c<2, OP_FLUXM>(a, st);
  } else {
    if (dir == 0) return launch_c<0, OP_FILTER>(a, st);
    if (dir == 1) return launch_c<1, OP_FILTER>(a, st);
    return launch_c<2, OP_FILTER>(a, st);
  }
}
