// astr_b200/csrc/sweep2.cu -- register-resident batched line solves (sm_100a), j and k sweeps.
//
// Operators: `fds%central` (src/derivative.F90:171-306) and `compact_filter`
// (src/filter.F90:112-285), same closures and coefficient tables as sweep.cu; the algebra
// (fresh-start chunk factorisation + two-sided reduced scan) is documented in linecore.h.
//
// Mapping.  A CTA owns a bundle of 32 neighbouring pencils (32 consecutive i => every line
// position is one contiguous 256-byte segment) and the whole line of each.
//  * TMA: one elected thread fetches the bundle as 2-3 `cp.async.bulk.tensor` boxes of
//    (32 pencils x <=256 line nodes) into a [node][32] shared-memory tile, completion on an
//    mbarrier.  The fetch of bundle b+1 is issued as soon as every warp has copied its rows of
//    bundle b into registers, so it runs under the arithmetic and the write-out of bundle b.
//  * warp w owns regular chunk w (<= 33 rows) of all 32 lines, lane = pencil: rows (+ stencil
//    overlap) come from the tile with immediate-offset, conflict-free LDS.64; the recurrences
//    run in registers; the coefficient tables live in __constant__ memory and are consumed as
//    immediate constant-bank operands.
//  * one named barrier per phase; the chunk boundary values of all elements of a pencil are
//    exchanged through 2 doubles per thread (S, S') and resolved by a redundant scan.
//  * the solution goes straight from registers to global memory, 256 contiguous bytes per warp
//    and row.  HBM traffic: 1 read + 1 write per node.
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

namespace {

// operator tables: written once per (operator, direction) by astr_sweep2_set_plan
__constant__ LinePlan c_plan[2][3];
__constant__ FilterCoef c_fc2;

struct Sweep2Args {
  Layout L;
  int NW;
  int rb, nbox;              // line nodes per TMA box, boxes per bundle
  int slot[ASTR_MAXF];       // 4th tensor coordinate of each input field
  double* out[ASTR_MAXF];
  int nf, epi, o_lo, o_hi;
};

__device__ __forceinline__ void cta_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double* row_ptr(double* p, unsigned gl8, int s) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(p) + (unsigned long long)gl8 * (unsigned)s);
}
__device__ __forceinline__ void st_if(double* p, double x, int on) {
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f64 [%0], %1; }" ::"l"(p), "d"(x), "r"(on) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

struct BundlePos { int i0, by, bz; };

// One instantiation per warp role, each with its own bundle loop (register allocation of the
// hot ROLE_MID path is not polluted by the closure code); all warps of the CTA meet at the
// same named barriers.
template <int DIR, int OP, int role>
__device__ __forceinline__ void role_loop(const Sweep2Args& a, const CUtensorMap* tm, const double* tile,
                                          double (*sS)[ASTR_EMAX][32], double (*sP)[ASTR_EMAX][32],
                                          unsigned long long* mbar) {
  constexpr int H = OpT<OP>::H;
  constexpr int L = ASTR_LMAX;
  constexpr int WN = L + 2 * H;

  const LinePlan& pl = c_plan[OP][DIR];
  const FilterCoef& fc = c_fc2;
  const Layout& Lay = a.L;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int NW = pl.NW, E = pl.E, n = pl.n;
  const int len = (role == ROLE_HEAD) ? pl.len0 : L;
  const int node0 = pl.first_node + plan_chunk_row(pl, w);
  const bool p0 = (pl.ntype == 1 || pl.ntype == 4), pm = (pl.ntype == 2 || pl.ntype == 4);
  const int epi = a.epi, o_lo = a.o_lo, o_hi = a.o_hi;
  const int w_lo = (epi == EPI_STOREZ) ? 0 : o_lo, w_hi = (epi == EPI_STOREZ) ? n : o_hi;
  // a chunk whose rows are all written unmodified needs no per-row range checks
  const bool plain = node0 >= o_lo && node0 + len - 1 <= o_hi;
  const int me = w + 1;
  const int nthreads = NW * 32;

  const int nbx = (Lay.im + 32) / 32;
  const int nby = (DIR == 2 ? Lay.jm : Lay.km) + 1;
  const int nbundles = nbx * nby * a.nf;
  const unsigned gl8 = (unsigned)((DIR == 1 ? Lay.sj : Lay.sk) * 8);   // bytes per line step
  const unsigned tile_bytes = (unsigned)(a.rb * a.nbox) * 256u;

  auto locate = [&](int bnd) {
    BundlePos p;
    p.i0 = (bnd % nbx) * 32; p.by = (bnd / nbx) % nby; p.bz = bnd / (nbx * nby);
    return p;
  };
  // elected thread: fetch one bundle into the tile.  Tensor coordinates (x, y, z, slot) =
  // (i + 16, j + 5, k + 5, field); out-of-range columns of a ragged bundle are zero-filled.
  auto fetch = [&](int bnd) {
    const BundlePos p = locate(bnd);
    mbar_expect_tx(mbar, tile_bytes);
    for (int b = 0; b < a.nbox; ++b) {
      void* dst = const_cast<double*>(tile) + (size_t)b * a.rb * 32;
      if (DIR == 1) tma_load_4d(dst, tm, p.i0 + ASTR_IOFF, b * a.rb, p.by + ASTR_HM, a.slot[p.bz], mbar);
      else tma_load_4d(dst, tm, p.i0 + ASTR_IOFF, p.by + ASTR_HM, b * a.rb, a.slot[p.bz], mbar);
    }
  };

  if (role == ROLE_HEAD && lane == 0 && (int)blockIdx.x < nbundles) fetch(blockIdx.x);

  // tl[node * 32] = f(node) of this thread's pencil
  const double* tl = tile + ASTR_HM * 32 + lane;
  unsigned phase = 0;
  int par = 0;
  for (int bnd = blockIdx.x; bnd < nbundles; bnd += gridDim.x, par ^= 1) {
    const BundlePos bp = locate(bnd);
    // lanes past the last pencil of a ragged bundle compute on halo / zero-filled columns and
    // write into the unused row padding (columns -16..-9 of the same rows): no predication
    const int valid = (bp.i0 + lane) <= Lay.im;
    const int icol = valid ? bp.i0 + lane : -ASTR_IOFF + (lane & 7);
    double* __restrict__ gout = a.out[bp.bz] + ((DIR == 1) ? Lay.idx(icol, 0, bp.by) : Lay.idx(icol, bp.by, 0));
    double* sSp = &sS[par][0][lane];
    double* sPp = &sP[par][0][lane];

    mbar_wait(mbar, phase);
    phase ^= 1;

    // ---- phase A: tile -> registers ----------------------------------------------------------
    double ov[4] = {0, 0, 0, 0};
    double sd[5] = {0, 0, 0, 0, 0};  // closure right-hand sides of this end
    if (role == ROLE_HEAD) {
      double hw[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) hw[k] = tl[(k - 5) * 32];
      closure_head<OP>(hw, p0, fc, sd);
    } else if (role == ROLE_TAIL) {
      double tw[14];
      const double* p = tl + (n - 8) * 32;
#pragma unroll
      for (int k = 0; k < 14; ++k) tw[k] = p[k * 32];
      closure_tail<OP>(tw, pm, fc, sd);
    }
    // the chunk (+ stencil overlap).  Window slots outside the halo (filter, interface ends:
    // 2 slots) only feed rows whose right-hand side is a closure row.
    double wv[WN];
    {
      const double* p = tl + (node0 - H) * 32;
#pragma unroll
      for (int s = 0; s < WN; ++s) {
        bool skip = false;
        if (OP == 1 && role == ROLE_HEAD && !p0 && s < 2) skip = true;
        if (OP == 1 && role == ROLE_TAIL && !pm && s >= WN - 2) skip = true;
        if (role == ROLE_HEAD && s >= len + 2 * H) skip = true;
        wv[s] = skip ? 0.0 : p[s * 32];
      }
    }
    cta_sync(1, nthreads);          // every warp holds its rows: the tile may be overwritten
    if (role == ROLE_HEAD && lane == 0 && bnd + (int)gridDim.x < nbundles) fetch(bnd + gridDim.x);

    // ---- phase B: eliminate, publish S / S' --------------------------------------------------
    double se[ASTR_SMAX] = {0, 0};   // forward-eliminated rows of the head / tail block
    if (role == ROLE_HEAD) {
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = (pl.sh == 1) ? sd[1 + k] : (k < 3 ? sd[(2 + k) < 5 ? 2 + k : 4] : 0.0);
      double yh, yt;
      spec_forward(pl.head, sd, se, yh, yt);
      sSp[0] = fma_(pl.el[0].gamma, yh, yt);
      sPp[0] = fma_(pl.el[0].gammap, yt, yh);
    } else if (role == ROLE_TAIL) {
      // rows nrows-nsl .. nrows-st-1 are regular rows with a closure right-hand side
      const int nov_t = pl.nsl - pl.st;
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = sd[k];
      double dt[ASTR_SMAX];
#pragma unroll
      for (int k = 0; k < ASTR_SMAX; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j == nov_t + k) v = sd[j];
        dt[k] = v;
      }
      double yh, yt;
      spec_forward(pl.tail, dt, se, yh, yt);
      sSp[(E - 1) * 32] = fma_(pl.el[E - 1].gamma, yh, yt);
      sPp[(E - 1) * 32] = fma_(pl.el[E - 1].gammap, yt, yh);
    }
    double e[L];
    {
      double yh, yt;
      chunk_forward<OP, role>(pl.reg, fc, wv, len, role == ROLE_HEAD ? p0 : pm, ov, e, yh, yt);
      sSp[me * 32] = fma_(pl.el[me].gamma, yh, yt);
      sPp[me * 32] = fma_(pl.el[me].gammap, yt, yh);
    }
    cta_sync(2, nthreads);

    // ---- phase C: boundary values of this chunk, solution, write-out -------------------------
    const ScanOut so = reduced_scan(
        pl, [&](int el) { return sSp[el * 32]; }, [&](int el) { return sPp[el * 32]; }, me);
    const double t_prev = scan_t_prev(pl, me, so.Pm1, so.Pb0);
    const double h_next = scan_h_next(pl, me, so.Pm, so.Pb1);

    auto put = [&](int node, double x) {
      if (node >= w_lo && node <= w_hi) {
        if (epi == EPI_STOREZ && (node < o_lo || node > o_hi)) x = 0.0;
        gout[(long long)node * (gl8 / 8)] = x;
      }
    };
    {
      // rows of chunk 0 can only fall below the written range, rows of the last chunk only
      // above it (the other end is at least one full chunk away)
      double* po = gout + (long long)node0 * (gl8 / 8);
      const int s_lo = w_lo - node0, s_hi = w_hi - node0, z_lo = o_lo - node0, z_hi = o_hi - node0;
      if (role == ROLE_MID || plain) {
        chunk_back<role == ROLE_HEAD ? ROLE_HEAD : ROLE_MID>(pl.reg, e, len, t_prev, h_next,
                                                             [&](int s, double x) { st_if(row_ptr(po, gl8, s), x, valid); });
      } else if (role == ROLE_HEAD) {
        chunk_back<ROLE_HEAD>(pl.reg, e, len, t_prev, h_next, [&](int s, double x) {
          if (s >= s_lo) *row_ptr(po, gl8, s) = (epi == EPI_STOREZ && s < z_lo) ? 0.0 : x;
        });
      } else {
        chunk_back<ROLE_MID>(pl.reg, e, len, t_prev, h_next, [&](int s, double x) {
          if (s <= s_hi) *row_ptr(po, gl8, s) = (epi == EPI_STOREZ && s > z_hi) ? 0.0 : x;
        });
      }
    }
    if (role == ROLE_HEAD) {   // element 0: P(0) = Pm1, P'(1) = Pb0 of element 1
      double x[ASTR_SMAX];
      spec_back(pl.head, se, 0.0, scan_h_next(pl, 0, so.Pm1, so.Pb0), x);
#pragma unroll
      for (int k = 0; k < ASTR_SMAX; ++k)
        if (k < pl.sh) put(pl.first_node + k, x[k]);
    } else if (role == ROLE_TAIL) {   // element E-1: P(E-2) = Pm, P'(E-1) = Pb1 of element E-2
      double x[ASTR_SMAX];
      spec_back(pl.tail, se, scan_t_prev(pl, E - 1, so.Pm, so.Pb1), 0.0, x);
#pragma unroll
      for (int k = 0; k < ASTR_SMAX; ++k)
        if (k < pl.st) put(pl.first_node + pl.nrows - pl.st + k, x[k]);
    }
    // sS/sP are double-buffered by bundle parity: a warp can run at most one barrier ahead of
    // the slowest one, so phase B of the next bundle never overwrites values still being read
  }
}

template <int DIR, int OP>
__global__ void __launch_bounds__(512, 1)
sweep2_kernel(const __grid_constant__ Sweep2Args a, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(128) double tile[];
  __shared__ double sS[2][ASTR_EMAX][32];
  __shared__ double sP[2][ASTR_EMAX][32];
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w == 0) role_loop<DIR, OP, ROLE_HEAD>(a, &tm, tile, sS, sP, &mbar);
  else if (w == a.NW - 1) role_loop<DIR, OP, ROLE_TAIL>(a, &tm, tile, sS, sP, &mbar);
  else role_loop<DIR, OP, ROLE_MID>(a, &tm, tile, sS, sP, &mbar);
}

int g_sms = 0;

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

struct PoolMap {
  const double* base = nullptr;
  int nslots = 0;
  CUtensorMap tm[3];        // per sweep direction (box shape differs); [0] unused
  int rb[3] = {0, 0, 0}, nbox[3] = {0, 0, 0};
};
PoolMap g_maps[2];

int make_map(PoolMap& pm, const Layout& L, int dir) {
  const int rows = (dir == 1 ? L.njt : L.nkt);
  const int nbox = (rows + 255) / 256;
  const int rb = (rows + nbox - 1) / nbox;
  pm.rb[dir] = rb; pm.nbox[dir] = nbox;
  const cuuint64_t gdim[4] = {(cuuint64_t)L.pitch, (cuuint64_t)L.njt, (cuuint64_t)L.nkt, (cuuint64_t)pm.nslots};
  const cuuint64_t gstr[3] = {(cuuint64_t)L.sj * 8, (cuuint64_t)L.sk * 8, (cuuint64_t)L.fstride * 8};
  const cuuint32_t box[4] = {32, (cuuint32_t)(dir == 1 ? rb : 1), (cuuint32_t)(dir == 2 ? rb : 1), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = g_encode(&pm.tm[dir], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(pm.base), gdim, gstr,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "sweep2: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return astr_fail_msg(buf);
  }
  return 0;
}

template <int DIR, int OP>
int launch2(Sweep2Args& a, const PoolMap& pm, cudaStream_t st) {
  auto kern = sweep2_kernel<DIR, OP>;
  const int threads = a.NW * 32;
  a.rb = pm.rb[DIR]; a.nbox = pm.nbox[DIR];
  const size_t smem = (size_t)a.rb * a.nbox * 256;
  static int occ_threads = 0, occ = 0;
  static size_t occ_smem = 0;
  if (occ_threads != threads || occ_smem != smem) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return astr_fail_msg("sweep2: kernel does not fit on an SM");
    occ_threads = threads; occ_smem = smem;
    if (!g_sms) {
      int dev = 0;
      CUDA_OK(cudaGetDevice(&dev));
      CUDA_OK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
  }
  const Layout& L = a.L;
  const long long nbx = (L.im + 32) / 32;
  const long long nby = (DIR == 2 ? L.jm : L.km) + 1;
  const long long nbundles = nbx * nby * a.nf;
  long long grid = (long long)g_sms * occ;
  if (grid > nbundles) grid = nbundles;
  kern<<<(unsigned)grid, threads, smem, st>>>(a, pm.tm[DIR]);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int astr_sweep2_set_plan(int dir, int optype, const LinePlan& plan, const FilterCoef& fc) {
  CUDA_OK(cudaMemcpyToSymbol(c_plan, &plan, sizeof(LinePlan), sizeof(LinePlan) * (size_t)(optype * 3 + dir)));
  CUDA_OK(cudaMemcpyToSymbol(c_fc2, &fc, sizeof(FilterCoef)));
  return 0;
}

// Registers an allocation of `nslots` fields (stride L.fstride) as TMA source `which` (0, 1).
int astr_sweep2_register_pool(int which, const double* base, int nslots, const Layout& L) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return astr_fail_msg("sweep2: cuTensorMapEncodeTiled unavailable");
    g_encode = (EncodeTiledFn)fn;
  }
  PoolMap& pm = g_maps[which];
  pm.base = base; pm.nslots = nslots;
  if (!base) return 0;
  for (int d = 1; d <= 2; ++d) {
    const int rc = make_map(pm, L, d);
    if (rc) return rc;
  }
  return 0;
}

// Returns -1 when the request is outside what these kernels cover (the caller then uses the
// shared-memory engine of sweep.cu), 0 on success, >0 on error.
int astr_launch_sweep2(int dir, int optype, const LinePlan& plan, const SweepArgs& s, cudaStream_t st) {
  if (dir == 0 || !plan.ok || s.epi == EPI_ADD) return -1;
  if (plan.NW * 32 > 512) return -1;
  // every input field must be a slot of one registered allocation
  const PoolMap* pm = nullptr;
  Sweep2Args a;
  for (int i = 0; i < s.nf; ++i) {
    const PoolMap* hit = nullptr;
    for (const PoolMap& m : g_maps) {
      if (!m.base) continue;
      const long long off = s.in[i] - m.base;
      if (off >= 0 && off % s.L.fstride == 0 && off / s.L.fstride < m.nslots) { hit = &m; a.slot[i] = (int)(off / s.L.fstride); }
    }
    if (!hit || (pm && hit != pm)) return -1;
    pm = hit;
  }
  if (!pm) return -1;
  if ((size_t)pm->rb[dir] * pm->nbox[dir] * 256 > 200 * 1024) return -1;
  a.L = s.L; a.NW = plan.NW;
  for (int i = 0; i < ASTR_MAXF; ++i) a.out[i] = s.out[i];
  a.nf = s.nf; a.epi = s.epi; a.o_lo = s.o_lo; a.o_hi = s.o_hi;
  if (optype == OP_DERIV) return dir == 1 ? launch2<1, OP_DERIV>(a, *pm, st) : launch2<2, OP_DERIV>(a, *pm, st);
  return dir == 1 ? launch2<1, OP_FILTER>(a, *pm, st) : launch2<2, OP_FILTER>(a, *pm, st);
}
