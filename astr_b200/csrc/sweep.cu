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
            const double e = fma_(-eprev, ac3[r], sf[k] * ac2[r]);
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
        const double e = fma_(-eprev, ac3[r], d * ac2[r]);
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
          const double e = fma_(-eprev, c3[u], d * c2[u]);
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
          const double e = fma_(-eprev, ac3[r], d * ac2[r]);
          tile[(node + 5) * SL] = e;
          eprev = e;
          ++r;
        }
      }
      if (c == C - 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k < nsl) {
            const double e = fma_(-eprev, ac3[r], sl[k] * ac2[r]);
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
          const double dp = fma_(pv[u], cin, ev[u]);
          const double g = fma_(-av[u], gnext, dp);
          ps[-u * SL] = g;
          gnext = g;
        }
        r -= 8;
      }
      for (; r >= ra; --r) {
        double* ps = tile + (first_node + r + 5) * SL;
        const double dp = fma_(pf[r], cin, *ps);
        const double g = fma_(-ac1[r], gnext, dp);
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
              x.x = fma_(qb[rx], xrow[ch[rx] * ASTR_BW], gv.x);
              x.y = fma_(qb[r + 1], xrow[ch[r + 1] * ASTR_BW], gv.y);
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
                x.x = fma_(qv, xin.x, gv.x);
                x.y = fma_(qv, xin.y, gv.y);
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
  static int occ = 0, attr_dev = -1;
  auto kern = sweep_kernel<DIR, OP, C, NG>;
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  if (dev != attr_dev) { attr_smem = 0; occ = 0; attr_dev = dev; g_num_sms = 0; }   // attributes are per device
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
