// astr_b200/csrc/common.cuh -- shared declarations of libastr_gpu.so (sm_100a only).
//
// Device data layout.  Every 3-D field, halo'd or not in the
// reference (src/commarray.F90:63-106), lives in ONE padded box on the device:
//
//   index(i,j,k) = org + i + pitch*(j + njt*k),   i in [-32, pitch-32), j in [-5, jm+5], ...
//
// * node i=0 sits at element 32 of each row, rows are `pitch` doubles long with pitch a
//   multiple of 32, and the pool is 256-byte aligned: node 0 of every row is 256-byte
//   aligned, so the 32-pencil bundles of the j / k line solves move exactly one aligned
//   256-byte block per line position.  The L2 slice hash works on 256-byte blocks: segments
//   that start on a 128-byte boundary (round 1: node 0 at element 16) straddle two blocks, and
//   tools/tma_probe_jk.cu measures 4.6-5.0 TB/s for the tile copy of such a layout against
//   5.5 TB/s for the aligned one (profiles/r02_tma_probe_jk.txt).  Node -6 is 16-byte aligned
//   (bulk copies of i lines) and rows never straddle sectors.
// * all fields share the same strides, so one index serves every array of a kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "linecore.h"

#define ASTR_HM 5
#define ASTR_IOFF 32        // element offset of node i=0 inside a row
#define ASTR_MAXF 16        // max fields per batched launch
#define ASTR_BW 16          // pencils per bundle (one 128-byte line of doubles)

struct Layout {
  int im, jm, km;           // nodes 0..im, 0..jm, 0..km
  int pitch, njt, nkt;      // row pitch (doubles), jm+11, km+11
  long long sj, sk;         // strides of j and k in doubles
  long long org;            // offset of node (0,0,0)
  long long fstride;        // doubles per field (multiple of 32)
  __host__ __device__ inline long long idx(int i, int j, int k) const {
    return org + i + sj * (long long)j + sk * (long long)k;
  }
};

// One pre-factored tridiagonal line operator (reference `type compact_scheme`,
// src/commtype.F90:13-23) plus the chunk-propagation tables of the partitioned solve.
struct LineOp {
  const double* ac1;        // src/commfunc.F90:752-774, row r = node - first_node
  const double* ac2;        // ac2[0] := 1, ac3[0] := 0 (row 0 is never scaled, :796)
  const double* ac3;
  const double* pf;         // pf[r] = prod_{m=cs(c(r))..r} (-ac3[m])
  const double* qb;         // qb[r] = prod_{m=r..ce(c(r))} (-ac1[m])
  int first_node, nrows, ntype, n, C;
  int nsf, nsl;             // closure rows at the first / last end
};

// Row range of chunk c of the partitioned solve: [chunk_start(c), chunk_start(c+1)).  Odd chunks
// start on an odd row and even chunks on an even row, so the two chunks that share a half-warp
// in the i-direction mapping always sit an odd number of elements apart (bank-conflict free).
__host__ __device__ inline int chunk_start(int c, int nrows, int C) {
  if (c <= 0) return 0;
  if (c >= C) return nrows;
  const int s = (c * nrows) / C;
  return (c & 1) ? (s | 1) : (s & ~1);
}

// OP_FLUXP / OP_FLUXM: compact 5th-order upwind interface flux, flux_compact with flux_uw / flux_dw
// (src/flux.F90:125-266); the solution at row `node` is the interface value fh(node), node -1..n
enum { OP_DERIV = 0, OP_FILTER = 1, OP_FLUXP = 2, OP_FLUXM = 3 };
enum { EPI_STORE = 0, EPI_STOREZ = 1, EPI_ADD = 2 };

struct SweepArgs {
  Layout L;
  LineOp op;
  const double* in[ASTR_MAXF];
  double* out[ASTR_MAXF];
  int nf;
  int epi;
  int o_lo, o_hi;           // nodes of the line that are written
  int sp;                   // i-direction: smem row pitch (doubles, == 2 mod 16)
};

#define CUDA_OK(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return astr_fail(#call, e_, __FILE__, __LINE__);   \
  } while (0)

int astr_fail(const char* what, cudaError_t e, const char* file, int line);
int astr_fail_msg(const char* msg);
void astr_count_launch(int n = 1);

// sweep.cu
int astr_set_filter_coef(const FilterCoef& fc);
int astr_set_flux_coef(double bfacmpld);
size_t astr_sweep_smem_bytes(int dir, int n, int C, int NG, int* sp_out);
int astr_launch_sweep(int dir, int optype, const SweepArgs& a, cudaStream_t st);
int astr_sweep_max_chunks(int nrows);
// sweep2.cu
int astr_sweep2_set_plan(int dir, int optype, const LinePlan& plan, const FilterCoef& fc);
int astr_sweep2_register_pool(int which, const double* base, int nslots, const Layout& L);
int astr_launch_sweep2(int dir, int optype, const LinePlan& plan, const SweepArgs& s, cudaStream_t st);
