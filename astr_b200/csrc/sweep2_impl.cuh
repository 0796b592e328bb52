// astr_b200/csrc/sweep2.cu -- register-resident batched line solves (sm_100a), all three directions.
//
// Operators: `fds%central` (src/derivative.F90:171-306), `compact_filter` (src/filter.F90:112-285)
// and `flux_compact` (src/flux.F90:125-266): same closures and coefficient tables as the reference;
// the algebra (fresh-start chunk factorisation + truncated two-sided reduced scan) is documented in
// linecore.h.
//
// Mapping.  A CTA owns a BUNDLE of 32 neighbouring pencils and the whole line of each; lane = pencil.
//  * warps 0..NW-1: regular chunk w (ASTR_LMAX rows) of all 32 lines.  Rows (+ stencil overlap) are
//    copied once from the shared-memory tile into registers; the three recurrences run in registers
//    with the coefficient tables as immediate constant-bank operands.  All regular warps execute the
//    same straight-line code, so none of them is the straggler of a barrier.
//  * warp NW ("special warp"): the head and tail blocks (closure rows + the rows that do not fill a
//    chunk), and the producer: it refills the tile as soon as every warp has arrived on the `empty`
//    mbarrier -- the regular warps never block on a CTA barrier between reading the tile and
//    publishing their boundary sums.
//  * one named barrier per bundle: after it every element resolves its two boundary values from the
//    published sums S / S' of the W nearest elements.
// j / k sweeps (sweep2_kernel): 32 consecutive i => every line position is one contiguous 256-byte
//    segment.  TMA (`cp.async.bulk.tensor`, 2-3 boxes of 32 pencils x <=256 nodes) fills a [node][32]
//    tile, completion on the `full` mbarrier; the solution goes straight from registers to global
//    memory, 256 contiguous bytes per warp and row.
// i sweeps (sweep2i_kernel): 32 consecutive j, each line contiguous in memory.  One bulk copy
//    (`cp.async.bulk.shared.global`) per line fills a [line][node] tile whose row pitch is 2 (mod 16)
//    doubles, so that the lane-per-line 16-byte accesses (LDS.128 / STS.128) are conflict-free; the
//    solution is staged through a 16-line output tile and leaves as full, aligned 128-byte lines.
// HBM traffic: 1 read + 1 write per node.
//
// This header holds the kernels; it is compiled three times (sweep2_i.cu, sweep2_j.cu, sweep2_k.cu: one
// translation unit per sweep direction, built in parallel), each with its own copy of the constant-bank tables.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

#include "sweep2_args.cuh"

namespace {

// operator tables: written once per (operator, direction) by astr_sweep2_set_plan
__constant__ LinePlan c_plan[4];      // [operator]: the tables of this translation unit's direction
__constant__ FilterCoef c_fc2;

constexpr int ESZ = ASTR_EMAX + 2 * ASTR_WPAD;   // padded element slots of the S / S' exchange

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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// Window loads as volatile asm: they stay in program order in front of the `empty` arrive, so the tile is
// released (and refilled) before the elimination starts instead of after it (the compiler otherwise sinks the
// loads next to their uses to save registers).
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct BundlePos { int bx, by, bz; };
__device__ __forceinline__ BundlePos locate(unsigned bnd, const Sweep2Args& a) {
  BundlePos p;
  const unsigned q = fdiv(bnd, a.dx);
  p.bx = (int)(bnd - q * a.dx.d);
  const unsigned z = fdiv(bnd, a.dxy);
  p.bz = (int)z;
  p.by = (int)(q - z * (unsigned)a.nby);
  return p;
}

// Where a lane's pencil of a j/k bundle lives: offset of node 0 of the line, field, and whether it exists
struct LanePos { long long off; int bz, valid, packed; };
struct FalseT { static constexpr bool value = false; };
struct TrueT { static constexpr bool value = true; };
template <int DIR, int PEN, bool PACKED>
__device__ __forceinline__ LanePos lane_pos(unsigned bnd, int lane, const Sweep2Args& a) {
  const Layout& Lay = a.L;
  LanePos r;
  r.packed = PACKED;
  int icol, by;
  if (!PACKED) {
    const BundlePos bp = locate(bnd, a);
    const int i0 = bp.bx * PEN;
    // lanes past the last pencil of a ragged bundle compute on zero-filled columns; their stores are predicated off
    r.valid = (lane < PEN) && (i0 + lane) <= Lay.im;
    icol = r.valid ? i0 + lane : 0;
    by = bp.by; r.bz = bp.bz;
  } else {
    const unsigned q = bnd - (unsigned)a.nfull;
    const unsigned z = fdiv(q, a.dpk);
    const int pb = (int)(q - z * a.dpk.d);
    const int g = (int)fdiv((unsigned)lane, a.drr);
    const int x = lane - g * a.rag_r;
    by = pb * a.rag_G + g;
    r.valid = g < a.rag_G && by < a.nby;
    if (!r.valid) by = 0;
    icol = a.rag_i0 + x; r.bz = (int)z;
  }
  r.off = (DIR == 1) ? Lay.idx(icol, 0, by) : Lay.idx(icol, by, 0);
  return r;
}
__device__ __forceinline__ double ldg_f64(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// epilogue of one solution value: is node written, and with what
struct OutRange {
  int w_lo, w_hi, o_lo, o_hi, storez;
  __device__ __forceinline__ OutRange(const Sweep2Args& a, int n)
      : w_lo(a.epi == EPI_STOREZ ? 0 : a.o_lo), w_hi(a.epi == EPI_STOREZ ? n : a.o_hi), o_lo(a.o_lo), o_hi(a.o_hi),
        storez(a.epi == EPI_STOREZ) {}
  __device__ __forceinline__ bool writes(int node) const { return node >= w_lo && node <= w_hi; }
  __device__ __forceinline__ double value(int node, double x) const {
    return (storez && (node < o_lo || node > o_hi)) ? 0.0 : x;
  }
};

// head / tail block right-hand sides, dispatched on the (warp-uniform) end types
template <int OP, int HWN, int S>
__device__ __forceinline__ void head_rhs_any(bool p0, const double (&hw)[HWN], const FilterCoef& fc, int nsf, int len,
                                             double (&d)[S]) {
  if (p0) head_rhs<OP, true>(hw, fc, nsf, len, d);
  else head_rhs<OP, false>(hw, fc, nsf, len, d);
}
// window of the head block, in groups of 8 slots (groups behind `lim` are skipped, the rest is zero):
// LD(k) returns f(first_node - HB + k); the first 3 slots may lie in front of the halo (filter, interface end)
template <int HWN, class LD>
__device__ __forceinline__ void load_head_window(double (&hw)[HWN], int lim, int hnode0, LD ld) {
#pragma unroll
  for (int g = 0; g < (HWN + 7) / 8; ++g) {
    if (g * 8 < lim) {
#pragma unroll
      for (int k = g * 8; k < g * 8 + 8 && k < HWN; ++k) {
        bool in = k < lim;
        if (k < 3) in = in && (hnode0 + k >= -ASTR_HM);
        hw[k] = in ? ld(k) : 0.0;
      }
    } else {
#pragma unroll
      for (int k = g * 8; k < g * 8 + 8 && k < HWN; ++k) hw[k] = 0.0;
    }
  }
}
template <int OP>
__device__ __forceinline__ void tail_rhs_any(bool pm, const double (&tw)[16], const FilterCoef& fc, int extra,
                                             double (&d)[ASTR_TS]) {
  if (pm) tail_rhs<OP, true>(tw, fc, extra, d);
  else tail_rhs<OP, false>(tw, fc, extra, d);
}

// =============================================================================================
// j / k sweeps
// =============================================================================================
// PEN pencils per bundle (lanes >= PEN idle), NT tiles.  NT = 2: the tile of bundle b+1 is requested while the
// warps still read bundle b, so the TMA stream never pauses; chosen when two tiles fit in shared memory.
struct JKCtx {
  double* tile[2];           // [node + 5][PEN]
  double* sS[2];             // [parity][ESZ][PEN]
  double* sP[2];
  unsigned long long* full;  // [NT]
  unsigned long long* empty; // [NT]
  int nthreads, nbundles;
};

// SHORT: this warp owns chunk 0 of a plan whose first chunk has pl.ls0 < ASTR_LMAX rows (linecore.h)
// PSHORT: the plan has a SHORT first chunk (the other chunks then start at sh + ls0 + (w - 1) L; without one the
// compile-time form sh + w L keeps the address arithmetic of the common kernels out of the registers)
template <int DIR, int OP, int PEN, int NT, bool SHORT, bool PSHORT>
__device__ __forceinline__ void regular_loop(const Sweep2Args& a, const JKCtx& c) {
  constexpr int H = OpT<OP>::H;
  constexpr int L = ASTR_LMAX;
  constexpr int WN = L + 2 * H;
  const LinePlan& pl = c_plan[OP];
  const FilterCoef& fc = c_fc2;
  const Layout& Lay = a.L;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = min(lane, PEN - 1);                  // lanes past PEN recompute the last pencil, never stored
  const int n = pl.n;
  const int node0 = pl.first_node + (SHORT ? pl.sh : (PSHORT ? pl.sh + pl.ls0 - L + w * L : pl.sh + w * L));
  const int len = SHORT ? pl.ls0 : L;
  const OutRange R(a, n);
  // a chunk whose rows are all written unmodified needs no per-row range checks
  const bool plain = node0 >= a.o_lo && node0 + len - 1 <= a.o_hi;
  const int me = w + 1;
  const unsigned gl8 = (unsigned)((DIR == 1 ? Lay.sj : Lay.sk) * 8);   // bytes per line step
  // tile[(node + 5) * PEN + col] = f(node) of this thread's pencil
  const int woff = (ASTR_HM + node0 - H) * PEN + col;
  // one bundle; PACKED is a compile-time copy of the body so that the tile path keeps its register allocation
  auto body = [&](int bnd, int it, auto packed_tag) {
    constexpr bool PACKED = decltype(packed_tag)::value;
    const int p = (NT == 2) ? (it & 1) : 0, par = it & 1;
    const unsigned ph = (NT == 2) ? ((unsigned)(it >> 1) & 1u) : ((unsigned)it & 1u);
    const LanePos lp = lane_pos<DIR, PEN, PACKED>((unsigned)bnd, lane, a);
    const int valid = lp.valid;
    double* __restrict__ gout = a.out[lp.bz] + lp.off;

    double wv[WN];
    if (!PACKED) {
      mbar_wait(&c.full[p], ph);
      // ---- tile -> registers: the chunk (+ stencil overlap) -------------------------------------
      const unsigned wpa = smem_u32(c.tile[p] + woff);
#pragma unroll
      for (int s = 0; s < WN; ++s) wv[s] = lds_f64(wpa + (unsigned)(s * PEN * 8));
      __syncwarp();
      if (lane == 0) mbar_arrive(&c.empty[p]);      // this warp is done with the tile
    } else {
      // ---- packed bundle: global memory -> registers (every load of the CTA precedes barrier 1, every store follows it)
      const double* gp = a.in[lp.bz] + lp.off + (long long)(node0 - H) * (gl8 / 8);
#pragma unroll
      for (int s = 0; s < WN; ++s) wv[s] = ldg_f64(row_ptr(const_cast<double*>(gp), gl8, s));
    }

    // ---- eliminate, publish S / S' ------------------------------------------------------------
    double* sSl = c.sS[par] + col;
    double* sPl = c.sP[par] + col;
    double e[L];
    {
      double yh, yt;
      if (SHORT) chunk_forward_short<OP>(pl.sreg, fc, wv, e, yh, yt);
      else chunk_forward<OP>(pl.reg, fc, wv, e, yh, yt);
      const double vs = fma_(pl.el[me + ASTR_WPAD].gamma, yh, yt), vp = fma_(pl.el[me + ASTR_WPAD].gammap, yt, yh);
      if (lane < PEN) { sSl[(me + ASTR_WPAD) * PEN] = vs; sPl[(me + ASTR_WPAD) * PEN] = vp; }
    }
    cta_sync(1, c.nthreads);

    // ---- boundary values of this chunk, solution, write-out -----------------------------------
    const ScanOut so = reduced_scan(pl, [&](int idx) { return sSl[idx * PEN]; }, [&](int idx) { return sPl[idx * PEN]; }, me);
    double* po = gout + (long long)node0 * (gl8 / 8);
    if (SHORT) {
      chunk_back_short(pl.sreg, e, so.t_prev, so.h_next, [&](int s, double x) {
        const int node = node0 + s;
        st_if(row_ptr(po, gl8, s), R.value(node, x), valid && s < len && R.writes(node));
      });
    } else if (plain) {
      chunk_back(pl.reg, e, so.t_prev, so.h_next, [&](int s, double x) { st_if(row_ptr(po, gl8, s), x, valid); });
    } else {
      chunk_back(pl.reg, e, so.t_prev, so.h_next, [&](int s, double x) {
        const int node = node0 + s;
        st_if(row_ptr(po, gl8, s), R.value(node, x), valid && R.writes(node));
      });
    }
    // sS/sP are double-buffered by bundle parity: a warp can run at most one barrier ahead of the slowest one
  };
  int it = 0, bnd = blockIdx.x;
  for (; bnd < a.nfull; bnd += gridDim.x, ++it) body(bnd, it, FalseT());
  for (; bnd < c.nbundles; bnd += gridDim.x, ++it) body(bnd, it, TrueT());
}

// HSL: slots of the head block that are processed (8 when the block has at most 8 rows -- the common case, e.g.
// 512-node lines -- else ASTR_HS): the short variant is straight-line code with everything in registers
template <int DIR, int OP, int PEN, int NT, int HSL>
__device__ __forceinline__ void special_loop(const Sweep2Args& a, const CUtensorMap* tm, const JKCtx& c) {
  constexpr int H = OpT<OP>::H, HB = OpT<OP>::HB;
  constexpr int HWN = (HSL + HB + H > OpT<OP>::CR) ? HSL + HB + H : OpT<OP>::CR;
  const LinePlan& pl = c_plan[OP];
  const FilterCoef& fc = c_fc2;
  const Layout& Lay = a.L;
  const int lane = threadIdx.x & 31;
  const int col = min(lane, PEN - 1);
  const int n = pl.n, E = pl.E;
  const bool p0 = (pl.ntype == 1 || pl.ntype == 4), pm = (pl.ntype == 2 || pl.ntype == 4);
  const OutRange R(a, n);
  const unsigned gl8 = (unsigned)((DIR == 1 ? Lay.sj : Lay.sk) * 8);
  const unsigned tile_bytes = (unsigned)(a.rb * a.nbox) * (unsigned)(PEN * 8);
  const int hwlim = max(OpT<OP>::CR, pl.sh + HB + H);  // CR: reach of the closure rows
  const int hnode0 = pl.first_node - HB;              // node of window slot 0
  const int tnode0 = pl.first_node + pl.nrows - pl.st; // node of the first tail row

  // elected thread: fetch one bundle into tile p.  Tensor coordinates (x, y, z, slot) =
  // (i + 16, j + 5, k + 5, field); out-of-range columns of a ragged bundle are zero-filled.
  auto fetch = [&](int bnd, int p) {
    const BundlePos q = locate(bnd, a);
    mbar_expect_tx(&c.full[p], tile_bytes);
    // the boxes of a tile are requested in an order rotated by the CTA index: neighbouring CTAs (neighbouring
    // pencils) otherwise walk the same planes / rows in lockstep and crowd the same DRAM pages
    for (int b0 = 0; b0 < a.nbox; ++b0) {
      const int b = (b0 + (int)blockIdx.x) % a.nbox;
      void* dst = c.tile[p] + (size_t)b * a.rb * PEN;
      if (DIR == 1) tma_load_4d(dst, tm, q.bx * PEN + ASTR_IOFF, b * a.rb, q.by + ASTR_HM, a.slot[q.bz], &c.full[p]);
      else tma_load_4d(dst, tm, q.bx * PEN + ASTR_IOFF, q.by + ASTR_HM, b * a.rb, a.slot[q.bz], &c.full[p]);
    }
  };
  if (lane == 0) {
    if ((int)blockIdx.x < a.nfull) fetch(blockIdx.x, 0);
    if (NT == 2 && (int)(blockIdx.x + gridDim.x) < a.nfull) fetch(blockIdx.x + gridDim.x, 1);
  }
  __syncwarp();

  auto body = [&](int bnd, int it, auto packed_tag) {
    constexpr bool PACKED = decltype(packed_tag)::value;
    const int p = (NT == 2) ? (it & 1) : 0, par = it & 1;
    const unsigned ph = (NT == 2) ? ((unsigned)(it >> 1) & 1u) : ((unsigned)it & 1u);
    const LanePos lp = lane_pos<DIR, PEN, PACKED>((unsigned)bnd, lane, a);
    const int valid = lp.valid;
    double* __restrict__ gout = a.out[lp.bz] + lp.off;

    double dh[HSL], dt[ASTR_TS];
    if (!PACKED) {
      const double* tl = c.tile[p] + ASTR_HM * PEN + col;   // tl[node * PEN] = f(node) of this thread's pencil
      const double* hp = tl + hnode0 * PEN;
      const double* tp = tl + (n - 10) * PEN;
      mbar_wait(&c.full[p], ph);
      // ---- tile -> right-hand sides of the two blocks (the windows die before the tile is released) ----
      {
        double hw[HWN], tw[16];
        load_head_window(hw, hwlim, hnode0, [&](int k) { return hp[k * PEN]; });
#pragma unroll
        for (int k = 0; k < 16; ++k) tw[k] = tp[k * PEN];
        head_rhs_any<OP>(p0, hw, fc, pl.nsf, pl.sh, dh);
        tail_rhs_any<OP>(pm, tw, fc, pl.st - pl.nsl, dt);
      }
      __syncwarp();
      // producer: once every warp has copied its rows into registers the tile is refilled with the bundle NT
      // iterations ahead, so that the fetch runs under the elimination, the scan and the write-out
      if (lane == 0) {
        mbar_arrive(&c.empty[p]);
        if (bnd + NT * (int)gridDim.x < a.nfull) {
          mbar_wait(&c.empty[p], ph);
          fetch(bnd + NT * gridDim.x, p);
        }
      }
      __syncwarp();
    } else {
      // packed bundle: the windows come straight from global memory
      double* gin = const_cast<double*>(a.in[lp.bz]) + lp.off;
      double* hp = gin + (long long)hnode0 * (gl8 / 8);
      double* tp = gin + (long long)(n - 10) * (gl8 / 8);
      double hw[HWN], tw[16];
      load_head_window(hw, hwlim, hnode0, [&](int k) { return ldg_f64(row_ptr(hp, gl8, k)); });
#pragma unroll
      for (int k = 0; k < 16; ++k) tw[k] = ldg_f64(row_ptr(tp, gl8, k));
      head_rhs_any<OP>(p0, hw, fc, pl.nsf, pl.sh, dh);
      tail_rhs_any<OP>(pm, tw, fc, pl.st - pl.nsl, dt);
    }

    // ---- eliminate, publish S / S' -------------------------------------------------------------
    // (in place: the right-hand sides become the eliminated rows)
    double (&he)[HSL] = dh;
    double (&te)[ASTR_TS] = dt;
    double* sSl = c.sS[par] + col;
    double* sPl = c.sP[par] + col;
    {
      double yh, yt;
      spec_forward(pl.head, dh, he, yh, yt);
      const double hs = fma_(pl.el[ASTR_WPAD].gamma, yh, yt), hq = fma_(pl.el[ASTR_WPAD].gammap, yt, yh);
      spec_forward(pl.tail, dt, te, yh, yt);
      const double ts = fma_(pl.el[E - 1 + ASTR_WPAD].gamma, yh, yt), tq = fma_(pl.el[E - 1 + ASTR_WPAD].gammap, yt, yh);
      if (lane < PEN) {
        sSl[ASTR_WPAD * PEN] = hs; sPl[ASTR_WPAD * PEN] = hq;
        sSl[(E - 1 + ASTR_WPAD) * PEN] = ts; sPl[(E - 1 + ASTR_WPAD) * PEN] = tq;
      }
    }
    cta_sync(1, c.nthreads);

    // ---- solution of the two blocks, write-out --------------------------------------------------
    auto GS = [&](int idx) { return sSl[idx * PEN]; };
    auto GP = [&](int idx) { return sPl[idx * PEN]; };
    {
      const ScanOut so = reduced_scan(pl, GS, GP, 0);
      double* po = gout + (long long)pl.first_node * (gl8 / 8);
      const int sh = pl.sh;
      spec_back(pl.head, he, so.t_prev, so.h_next, [&](int s, double x) {
        const int node = pl.first_node + s;
        st_if(row_ptr(po, gl8, s), R.value(node, x), valid && s < sh && R.writes(node));
      });
    }
    {
      const ScanOut so = reduced_scan(pl, GS, GP, E - 1);
      double* po = gout + (long long)tnode0 * (gl8 / 8);
      const int st = pl.st;
      spec_back(pl.tail, te, so.t_prev, so.h_next, [&](int s, double x) {
        const int node = tnode0 + s;
        st_if(row_ptr(po, gl8, s), R.value(node, x), valid && s < st && R.writes(node));
      });
    }
  };
  int it = 0, bnd = blockIdx.x;
  for (; bnd < a.nfull; bnd += gridDim.x, ++it) body(bnd, it, FalseT());
  for (; bnd < c.nbundles; bnd += gridDim.x, ++it) body(bnd, it, TrueT());
}

// SHORT: chunk 0 of the plan is short (pl.ls0 < ASTR_LMAX); HSL: slots of the head block that are processed.  Separate
// kernels, so that the common one (SHORT = false, HSL = 8) keeps the register allocation of its straight-line code
template <int DIR, int OP, int PEN, int NT, bool SHORT, int HSL>
__global__ void __launch_bounds__(512, 1)
sweep2_kernel(const __grid_constant__ Sweep2Args a, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long mbar[4];
  const LinePlan& pl = c_plan[OP];
  const int NW = pl.NW;
  JKCtx c;
  const int tile_doubles = a.rb * a.nbox * PEN;
  c.tile[0] = smem;
  c.tile[1] = smem + (NT == 2 ? tile_doubles : 0);
  double* sbase = smem + NT * tile_doubles;
  c.sS[0] = sbase; c.sP[0] = sbase + ESZ * PEN; c.sS[1] = sbase + 2 * ESZ * PEN; c.sP[1] = sbase + 3 * ESZ * PEN;
  c.full = &mbar[0]; c.empty = &mbar[2];
  c.nthreads = (NW + 1) * 32;
  c.nbundles = a.nbundles;
  for (int i = threadIdx.x; i < 4 * ESZ * PEN; i += blockDim.x) sbase[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int t = 0; t < NT; ++t) { mbar_init(&c.full[t], 1); mbar_init(&c.empty[t], NW + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w < NW) {
    if (SHORT && w == 0) regular_loop<DIR, OP, PEN, NT, true, true>(a, c);
    else regular_loop<DIR, OP, PEN, NT, false, SHORT>(a, c);
  } else special_loop<DIR, OP, PEN, NT, HSL>(a, &tm, c);
}

// =============================================================================================
// i sweeps
// =============================================================================================
// Two line tiles [LINES][sp] (double buffer).  A bundle lives in ONE tile from its fetch to its write-out: the
// chunk windows are read into registers, the solution is staged back IN PLACE (every warp has its window by
// then: barrier 1 lies in between), and one bulk store per line (`cp.async.bulk.global.shared::cta`) moves
// the written node range out.  Meanwhile the other tile receives the next bundle.
// Warp roles: warps 0..NW-1 regular chunks, warp NW the head / tail blocks, warp NW+1 the PRODUCER: it alone
// talks to the bulk engine -- waits until every compute warp has staged its part of a tile (`staged` mbarrier,
// a non-blocking arrive for the compute warps), issues the stores, waits until the engine has READ the tile and
// refills it with the bundle after next.  The compute warps therefore only ever wait for data (`full`) and for
// each other's boundary sums (barrier 1); tools/tma_probe.cu shows that one warp per SM driving 2 x 24 lines
// through the bulk engine this way sustains the full HBM copy bandwidth.
// LINES = 32 when two tiles fit in shared memory, else 24 (lanes 24..31 idle): 512-node lines take 4.2 KB each.
struct ICtx {
  double* tile[2];           // [LINES][sp]: column c of a row = node c - 6
  double* sS[2];             // [parity][ESZ][LINES]: boundary sums of the elements; a parity is reused two bundles
  double* sP[2];             // later, after every reader has signalled `staged` and the tile has been refilled
  unsigned long long* full;  // [2] tile filled (producer -> compute warps)
  unsigned long long* staged;// [2] solution staged (compute warps -> producer)
  int sp, ncompute, nbundles;
};

// regular chunk: x of rows s = 0..L-1 into the tile row (orow = address of row 0's node)
template <int OP, bool SHORT>
__device__ __forceinline__ void stage_chunk(double* orow, const double (&x)[ASTR_LMAX], int len) {
  constexpr int L = ASTR_LMAX;
  // 16-byte aligned when H is even (node0 - H + 6 is even); a SHORT chunk stages its first len (even) rows
  if ((OpT<OP>::H & 1) == 0) {
#pragma unroll
    for (int s = 0; s < L; s += 2)
      if (!SHORT || s < len) *reinterpret_cast<double2*>(orow + s) = make_double2(x[s], x[s + 1]);
  } else {
    orow[0] = x[0];
#pragma unroll
    for (int s = 1; s + 1 < L; s += 2) {
      if (!SHORT || s + 1 < len) *reinterpret_cast<double2*>(orow + s) = make_double2(x[s], x[s + 1]);
      else if (s < len) orow[s] = x[s];                 // the last row of a SHORT chunk (len is even)
    }
    if (!SHORT) orow[L - 1] = x[L - 1];
  }
}
// a compute warp is done with its part of the tile: make the staged values visible to the bulk engine, tell the producer
__device__ __forceinline__ void signal_staged(unsigned long long* staged) {
  fence_async_smem();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(staged);
}

template <int OP, int LINES, bool SHORT, bool PSHORT>
__device__ __forceinline__ void regular_loop_i(const Sweep2Args& a, const ICtx& c) {
  constexpr int H = OpT<OP>::H;
  constexpr int L = ASTR_LMAX;
  constexpr int WN = L + 2 * H;
  const LinePlan& pl = c_plan[OP];
  const FilterCoef& fc = c_fc2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int node0 = pl.first_node + (SHORT ? pl.sh : (PSHORT ? pl.sh + pl.ls0 - L + w * L : pl.sh + w * L));
  const int len = SHORT ? pl.ls0 : L;                      // even (build_line_plan, align_even)
  const int me = w + 1;
  const int row = min(lane, LINES - 1);                    // lanes past LINES recompute the last line, never staged
  const bool owner = lane < LINES;
  // 16-byte loads: node0 - H + 6 is even (build_line_plan, align_even)
  const int woff = row * c.sp + 6 + (node0 - H);
  int it = 0;
  for (int bnd = blockIdx.x; bnd < c.nbundles; bnd += gridDim.x, ++it) {
    const int p = it & 1;
    const unsigned ph = (unsigned)(it >> 1) & 1u;
    double* tile = c.tile[p];
    double* sSl = c.sS[p] + row;
    double* sPl = c.sP[p] + row;
    mbar_wait(&c.full[p], ph);
    double wv[WN];
    {
      const unsigned wpa = smem_u32(tile + woff);
#pragma unroll
      for (int s = 0; s < WN / 2; ++s) { const double2 v = lds_v2f64(wpa + (unsigned)s * 16u); wv[2 * s] = v.x; wv[2 * s + 1] = v.y; }
    }
    double e[L];
    {
      double yh, yt;
      if (SHORT) chunk_forward_short<OP>(pl.sreg, fc, wv, e, yh, yt);
      else chunk_forward<OP>(pl.reg, fc, wv, e, yh, yt);
      const double vs = fma_(pl.el[me + ASTR_WPAD].gamma, yh, yt), vp = fma_(pl.el[me + ASTR_WPAD].gammap, yt, yh);
      if (owner) { sSl[(me + ASTR_WPAD) * LINES] = vs; sPl[(me + ASTR_WPAD) * LINES] = vp; }
    }
    cta_sync(1, c.ncompute);            // every compute warp holds its window; S / S' are published
    const ScanOut so = reduced_scan(pl, [&](int idx) { return sSl[idx * LINES]; }, [&](int idx) { return sPl[idx * LINES]; }, me);
#ifndef ASTR_SKELETON
    {
      double x = so.h_next;
#pragma unroll
      for (int s = L - 1; s >= 0; --s) {
        if (SHORT) x = fma_(-pl.sreg.ac1[s], x, fma_(-pl.sreg.ev[s], so.t_prev, e[s] * pl.sreg.q[s]));
        else x = fma_(-pl.reg.ac1[s], x, fma_(-pl.reg.ev[s], so.t_prev, e[s]));
        e[s] = x;
      }
    }
#endif
    if (owner) stage_chunk<OP, SHORT>(tile + row * c.sp + 6 + node0, e, len);
    signal_staged(&c.staged[p]);
  }
}

template <int OP, int LINES, int HSL>
__device__ __forceinline__ void special_loop_i(const Sweep2Args& a, const ICtx& c) {
  constexpr int H = OpT<OP>::H, HB = OpT<OP>::HB;
  constexpr int HWN = (HSL + HB + H > OpT<OP>::CR) ? HSL + HB + H : OpT<OP>::CR;
  const LinePlan& pl = c_plan[OP];
  const FilterCoef& fc = c_fc2;
  const int lane = threadIdx.x & 31;
  const int n = pl.n, E = pl.E;
  const bool p0 = (pl.ntype == 1 || pl.ntype == 4), pm = (pl.ntype == 2 || pl.ntype == 4);
  const int hwlim = max(OpT<OP>::CR, pl.sh + HB + H);
  const int hnode0 = pl.first_node - HB;
  const int tnode0 = pl.first_node + pl.nrows - pl.st;
  const int row = min(lane, LINES - 1);
  const bool owner = lane < LINES;
  int it = 0;
  for (int bnd = blockIdx.x; bnd < c.nbundles; bnd += gridDim.x, ++it) {
    const int p = it & 1;
    const unsigned ph = (unsigned)(it >> 1) & 1u;
    double* tile = c.tile[p];
    double* sSl = c.sS[p] + row;
    double* sPl = c.sP[p] + row;
    mbar_wait(&c.full[p], ph);
    const double* lrow = tile + row * c.sp + 6;            // lrow[node] = f(node) of this thread's line
    double dh[HSL], dt[ASTR_TS];
    {
      double hw[HWN], tw[16];
      load_head_window(hw, hwlim, hnode0, [&](int k) { return lrow[hnode0 + k]; });
#pragma unroll
      for (int k = 0; k < 16; ++k) tw[k] = lrow[n - 10 + k];
      head_rhs_any<OP>(p0, hw, fc, pl.nsf, pl.sh, dh);
      tail_rhs_any<OP>(pm, tw, fc, pl.st - pl.nsl, dt);
    }
    double (&he)[HSL] = dh;                        // in place: the right-hand sides become the eliminated rows
    double (&te)[ASTR_TS] = dt;
    {
      double yh, yt;
      spec_forward(pl.head, dh, he, yh, yt);
      const double hs = fma_(pl.el[ASTR_WPAD].gamma, yh, yt), hp = fma_(pl.el[ASTR_WPAD].gammap, yt, yh);
      spec_forward(pl.tail, dt, te, yh, yt);
      const double ts = fma_(pl.el[E - 1 + ASTR_WPAD].gamma, yh, yt), tp = fma_(pl.el[E - 1 + ASTR_WPAD].gammap, yt, yh);
      if (owner) {
        sSl[ASTR_WPAD * LINES] = hs; sPl[ASTR_WPAD * LINES] = hp;
        sSl[(E - 1 + ASTR_WPAD) * LINES] = ts; sPl[(E - 1 + ASTR_WPAD) * LINES] = tp;
      }
    }
    cta_sync(1, c.ncompute);
    auto GS = [&](int idx) { return sSl[idx * LINES]; };
    auto GP = [&](int idx) { return sPl[idx * LINES]; };
    {
      const ScanOut s0 = reduced_scan(pl, GS, GP, 0);
      spec_back(pl.head, he, s0.t_prev, s0.h_next, [&](int s, double x) { he[s] = x; });
      const ScanOut s1 = reduced_scan(pl, GS, GP, E - 1);
      spec_back(pl.tail, te, s1.t_prev, s1.h_next, [&](int s, double x) { te[s] = x; });
    }
    if (owner) {
      double* orow = tile + row * c.sp + 6;                // orow[node]
#pragma unroll
      for (int g = 0; g < HSL / 8; ++g)
        if (g * 8 < pl.sh) {
#pragma unroll
          for (int s = g * 8; s < g * 8 + 8; ++s) if (s < pl.sh) orow[pl.first_node + s] = he[s];
        }
#pragma unroll
      for (int s = 0; s < ASTR_TS; ++s) if (s < pl.st) orow[tnode0 + s] = te[s];
    }
    signal_staged(&c.staged[p]);
  }
}

// The producer warp: lane l moves line l of a bundle.
template <int LINES>
__device__ __forceinline__ void producer_loop_i(const Sweep2Args& a, const ICtx& c, int n) {
  const Layout& Lay = a.L;
  const int lane = threadIdx.x & 31;
  const OutRange R(a, n);
  const unsigned line_bytes = (unsigned)(((n + 13) & ~1) * 8);   // nodes -6..n+5 (+1 when n is odd)
  // written nodes: one bulk store covers the node pairs that are written unmodified, [nf0, nf1); the (at most
  // a few) nodes of [w_lo, w_hi] outside go out as single stores
  const int lo = max(R.w_lo, R.storez ? R.o_lo : R.w_lo), hi = min(R.w_hi, R.storez ? R.o_hi : R.w_hi);
  const int nf0 = (lo + 1) & ~1;
  const int nf1 = max((hi + 1) & ~1, nf0);
  const int e0 = min(nf0, R.w_hi + 1), e1 = max(nf1, R.w_lo);
  // bundle -> field and first line; this lane's line as an offset (node 0) into the field
  struct LinePos { int bz, nvalid; long long off; };
  auto line_of = [&](int bnd) {
    LinePos r;
    const unsigned z = fdiv((unsigned)bnd, a.dx);             // dx: bundles per field
    const int l0 = (int)((unsigned)bnd - z * a.dx.d) * LINES;
    r.bz = (int)z;
    r.nvalid = min(LINES, a.nlines - l0);
    const unsigned l = (unsigned)min(l0 + lane, a.nlines - 1);
    const unsigned k = fdiv(l, a.dline);
    r.off = Lay.idx(0, (int)(l - k * a.dline.d), (int)k);
    return r;
  };
  auto fetch = [&](int bnd, int p) {
    const LinePos q = line_of(bnd);
    if (lane == 0) mbar_expect_tx(&c.full[p], line_bytes * (unsigned)q.nvalid);
    __syncwarp();
    if (lane < q.nvalid) bulk_load(c.tile[p] + lane * c.sp, a.in[q.bz] + q.off - 6, line_bytes, &c.full[p]);
  };
  if ((int)blockIdx.x < c.nbundles) fetch(blockIdx.x, 0);
  if ((int)(blockIdx.x + gridDim.x) < c.nbundles) fetch(blockIdx.x + gridDim.x, 1);
  int it = 0;
  for (int bnd = blockIdx.x; bnd < c.nbundles; bnd += gridDim.x, ++it) {
    const int p = it & 1;
    const unsigned ph = (unsigned)(it >> 1) & 1u;
    const LinePos bp = line_of(bnd);
    const int nl = bp.nvalid;                              // valid lines of the bundle
    mbar_wait(&c.staged[p], ph);                           // every compute warp has staged its rows of tile p
    if (lane < nl) {
      const double* srow = c.tile[p] + lane * c.sp + 6;    // srow[node]
      double* grow = a.out[bp.bz] + bp.off;                // grow[node]
      if (nf1 > nf0) bulk_store(grow + nf0, srow + nf0, (unsigned)(nf1 - nf0) * 8u);
      for (int node = R.w_lo; node < e0; ++node) grow[node] = R.value(node, srow[node]);
      for (int node = e1; node <= R.w_hi; ++node) grow[node] = R.value(node, srow[node]);
    }
    bulk_commit();
    bulk_wait_read();                                      // the engine has read the tile (the data may still be on its way)
    __syncwarp();
    if (bnd + 2 * (int)gridDim.x < c.nbundles) fetch(bnd + 2 * gridDim.x, p);
  }
  bulk_wait_all();
}

template <int OP, int LINES, bool SHORT, int HSL>
__global__ void __launch_bounds__(544, 1) sweep2i_kernel(const __grid_constant__ Sweep2Args a) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long mbar[4];
  const LinePlan& pl = c_plan[OP];
  ICtx c;
  c.sp = a.sp;
  c.tile[0] = smem;
  c.tile[1] = smem + LINES * c.sp;
  double* sbase = smem + 2 * LINES * c.sp;
  c.sS[0] = sbase; c.sP[0] = sbase + ESZ * LINES; c.sS[1] = sbase + 2 * ESZ * LINES; c.sP[1] = sbase + 3 * ESZ * LINES;
  c.full = &mbar[0]; c.staged = &mbar[2];
  const int NW = pl.NW;
  c.ncompute = (NW + 1) * 32;
  c.nbundles = a.nbundles;
  for (int i = threadIdx.x; i < 4 * ESZ * LINES; i += blockDim.x) sbase[i] = 0.0;
  if (threadIdx.x == 0) {
    mbar_init(&c.full[0], 1); mbar_init(&c.full[1], 1);
    mbar_init(&c.staged[0], NW + 1); mbar_init(&c.staged[1], NW + 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w < NW) {
    if (SHORT && w == 0) regular_loop_i<OP, LINES, true, true>(a, c);
    else regular_loop_i<OP, LINES, false, SHORT>(a, c);
  } else if (w == NW) special_loop_i<OP, LINES, HSL>(a, c);
  else producer_loop_i<LINES>(a, c, pl.n);
}

int g_sms = 0;

struct LaunchCache { int dev = -1, threads = 0, occ = 0; size_t smem = 0; };

template <class K>
int prepare(K kern, LaunchCache& lc, int threads, size_t smem) {
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  if (lc.dev == dev && lc.threads == threads && lc.smem == smem) return 0;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lc.occ, kern, threads, smem));
  if (lc.occ < 1) return astr_fail_msg("sweep2: kernel does not fit on an SM");
  CUDA_OK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  lc.dev = dev; lc.threads = threads; lc.smem = smem;
  return 0;
}

template <int DIR, int OP, int PEN, int NT, bool SHORT, int HSL>
int launch2_var(Sweep2Args& a, const CUtensorMap& tm, int NW, size_t smem, cudaStream_t st) {
  auto kern = sweep2_kernel<DIR, OP, PEN, NT, SHORT, HSL>;
  const int threads = (NW + 1) * 32;
  static LaunchCache lc;
  const int rc = prepare(kern, lc, threads, smem);
  if (rc) return rc;
  const Layout& L = a.L;
  long long nbx = (L.im + PEN) / PEN;
  const long long nby = (DIR == 2 ? L.jm : L.km) + 1;
  // left-over pencils of a row: packed rag_G rows to a bundle when at least two rows fit
  const int rag = (L.im + 1) % PEN;
  long long npk = 0;
  a.rag_r = 1; a.rag_G = 0; a.rag_i0 = 0;
  if (rag > 0 && 2 * rag <= 32 && nbx > 1) {
    nbx -= 1;
    a.rag_r = rag; a.rag_G = 32 / rag; a.rag_i0 = (int)nbx * PEN;
    npk = (nby + a.rag_G - 1) / a.rag_G;
  }
  a.nfull = (int)(nbx * nby * a.nf);
  const long long nbundles = a.nfull + npk * a.nf;
  a.dx = make_fastdiv((unsigned)nbx); a.dxy = make_fastdiv((unsigned)(nbx * nby)); a.nby = (int)nby;
  a.dpk = make_fastdiv((unsigned)(npk > 0 ? npk : 1)); a.drr = make_fastdiv((unsigned)a.rag_r);
  a.dpair = make_fastdiv(1); a.dline = make_fastdiv(1); a.nlines = 0;
  a.nbundles = (int)nbundles;
  long long grid = (long long)g_sms * lc.occ;
  if (grid > nbundles) grid = nbundles;
  kern<<<(unsigned)grid, threads, smem, st>>>(a, tm);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}

// kernel variant of a plan: short head block (<= 8 rows) with or without a SHORT first chunk, or the long head block
template <int DIR, int OP, int PEN, int NT>
int launch2_cfg(Sweep2Args& a, const CUtensorMap& tm, const LinePlan& plan, size_t smem, cudaStream_t st) {
  if (plan.sh > 8) return launch2_var<DIR, OP, PEN, NT, false, ASTR_HS>(a, tm, plan.NW, smem, st);
  if (plan.ls0 != ASTR_LMAX) return launch2_var<DIR, OP, PEN, NT, true, 8>(a, tm, plan.NW, smem, st);
  return launch2_var<DIR, OP, PEN, NT, false, 8>(a, tm, plan.NW, smem, st);
}

// two 32-pencil tiles when they fit (lines up to ~400 nodes), else one.  24-pencil double buffering, which is what
// the i kernel does with whole lines, measured SLOWER here (2.44 / 2.53 ms against 2.19 / 2.32 ms for the five-field
// derivative sweeps at 512^3): the 192-byte row segments cost more DRAM efficiency than the overlap returns
template <int DIR, int OP>
int launch2(Sweep2Args& a, const CUtensorMap& tm, const LinePlan& plan, cudaStream_t st) {
  const size_t rows = (size_t)a.rb * a.nbox;
  const size_t cap = 227 * 1024 - 64;
  const size_t s232 = (2 * rows * 32 + 4 * ESZ * 32) * sizeof(double);
  const size_t s132 = (rows * 32 + 4 * ESZ * 32) * sizeof(double);
  if (s232 <= cap) return launch2_cfg<DIR, OP, 32, 2>(a, tm, plan, s232, st);
  if (s132 <= cap) return launch2_cfg<DIR, OP, 32, 1>(a, tm, plan, s132, st);
  return -1;
}

template <int OP, int LINES, bool SHORT, int HSL>
int launch2i_var(Sweep2Args& a, const LinePlan& plan, size_t smem, cudaStream_t st) {
  auto kern = sweep2i_kernel<OP, LINES, SHORT, HSL>;
  const int threads = (plan.NW + 2) * 32;       // regular chunks + head/tail warp + producer warp
  static LaunchCache lc;
  const int rc = prepare(kern, lc, threads, smem);
  if (rc) return rc;
  const Layout& L = a.L;
  const long long nlines = (long long)(L.jm + 1) * (L.km + 1);
  const long long nbx = (nlines + LINES - 1) / LINES;       // bundles per field
  const long long nbundles = nbx * a.nf;
  a.dx = make_fastdiv((unsigned)nbx); a.dxy = make_fastdiv(1); a.nby = 1;
  a.dline = make_fastdiv((unsigned)(L.jm + 1)); a.nlines = (int)nlines;
  a.nfull = (int)nbundles; a.rag_r = 1; a.rag_G = 0; a.rag_i0 = 0; a.dpk = make_fastdiv(1); a.drr = make_fastdiv(1);
  a.dpair = make_fastdiv(1); a.nbundles = (int)nbundles;
  long long grid = (long long)g_sms * lc.occ;
  if (grid > nbundles) grid = nbundles;
  kern<<<(unsigned)grid, threads, smem, st>>>(a);
  astr_count_launch();
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <int OP, int LINES>
int launch2i_lines(Sweep2Args& a, const LinePlan& plan, size_t smem, cudaStream_t st) {
  if (plan.sh > 8) return launch2i_var<OP, LINES, false, ASTR_HS>(a, plan, smem, st);
  if (plan.ls0 != ASTR_LMAX) return launch2i_var<OP, LINES, true, 8>(a, plan, smem, st);
  return launch2i_var<OP, LINES, false, 8>(a, plan, smem, st);
}

template <int OP>
int launch2i(Sweep2Args& a, const LinePlan& plan, cudaStream_t st) {
  const int n = plan.n;
  // the chunk windows must start on a 16-byte boundary of the shared-memory line (build_line_plan, align_even)
  if (((plan.first_node + plan.sh - OpT<OP>::H + 6) & 1) != 0) return -1;
  // row pitch of the line tiles: sp / 2 odd, so that the lane-per-line 16-byte accesses of a quarter warp hit
  // 8 distinct 16-byte bank groups
  int sp = n + 12 + (n & 1);
  while ((sp & 15) != 2) sp += 2;
  a.sp = sp;
  const size_t cap = 227 * 1024 - 64;
  const size_t s32 = ((size_t)2 * 32 * sp + 4 * ESZ * 32) * sizeof(double);
  const size_t s24 = ((size_t)2 * 24 * sp + 4 * ESZ * 24) * sizeof(double);
  if (s32 <= cap) return launch2i_lines<OP, 32>(a, plan, s32, st);
  if (s24 <= cap) return launch2i_lines<OP, 24>(a, plan, s24, st);
  return -1;
}


int set_plan_local(int optype, const LinePlan& plan, const FilterCoef& fc) {
  CUDA_OK(cudaMemcpyToSymbol(c_plan, &plan, sizeof(LinePlan), sizeof(LinePlan) * (size_t)optype));
  CUDA_OK(cudaMemcpyToSymbol(c_fc2, &fc, sizeof(FilterCoef)));
  return 0;
}

}  // namespace
