// astr_b200/csrc/sweep2.cu -- host side of the register line-solve engine: tensor maps of the field pools and
// the dispatch to the per-direction translation units (sweep2_i.cu, sweep2_j.cu, sweep2_k.cu; kernels in sweep2_impl.cuh).
#include "common.cuh"
#include <cuda.h>
#include <cstdio>
#include "sweep2_args.cuh"

int astr_sweep2_set_plan_i(int optype, const LinePlan& plan, const FilterCoef& fc);
int astr_sweep2_set_plan_j(int optype, const LinePlan& plan, const FilterCoef& fc);
int astr_sweep2_set_plan_k(int optype, const LinePlan& plan, const FilterCoef& fc);
int astr_sweep2_launch_i(int optype, Sweep2Args& a, const LinePlan& plan, cudaStream_t st);
int astr_sweep2_launch_j(int optype, Sweep2Args& a, const CUtensorMap& tm, const LinePlan& plan, cudaStream_t st);
int astr_sweep2_launch_k(int optype, Sweep2Args& a, const CUtensorMap& tm, const LinePlan& plan, cudaStream_t st);

namespace {

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

struct PoolMap {
  const double* base = nullptr;
  int nslots = 0;
  CUtensorMap tm[3][2];     // [sweep direction][0: 32-pencil boxes, 1: 24-pencil boxes]; direction 0 unused
  int rb[3] = {0, 0, 0}, nbox[3] = {0, 0, 0};
};
constexpr int NPOOLS = 3;
PoolMap g_maps[NPOOLS];

int make_map(PoolMap& pm, const Layout& L, int dir, int pv) {
  const int pen = pv ? 24 : 32;
  const int rows = (dir == 1 ? L.njt : L.nkt);
  const int nbox = (rows + 255) / 256;
  const int rb = (((rows + nbox - 1) / nbox) + 1) & ~1;     // even: 192-byte rows x rb stays a multiple of 128 bytes (TMA destination alignment)
  pm.rb[dir] = rb; pm.nbox[dir] = nbox;
  const cuuint64_t gdim[4] = {(cuuint64_t)L.pitch, (cuuint64_t)L.njt, (cuuint64_t)L.nkt, (cuuint64_t)pm.nslots};
  const cuuint64_t gstr[3] = {(cuuint64_t)L.sj * 8, (cuuint64_t)L.sk * 8, (cuuint64_t)L.fstride * 8};
  const cuuint32_t box[4] = {(cuuint32_t)pen, (cuuint32_t)(dir == 1 ? rb : 1), (cuuint32_t)(dir == 2 ? rb : 1), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  // L2 promotion no wider than the alignment of the box rows (256-byte rows start on a 128-byte boundary, 192-byte
  // rows on a 64-byte boundary): the 256-byte promotion made every row fetch two extra half lines (ncu: 17 % more
  // L2 read sectors than TMA bytes)
  const CUresult r = g_encode(&pm.tm[dir][pv], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(pm.base), gdim, gstr,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              pv ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "sweep2: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return astr_fail_msg(buf);
  }
  return 0;
}


}  // namespace


// each direction's translation unit keeps its own constant-bank copy of the operator tables
int astr_sweep2_set_plan(int dir, int optype, const LinePlan& plan, const FilterCoef& fc) {
  if (dir == 0) return astr_sweep2_set_plan_i(optype, plan, fc);
  if (dir == 1) return astr_sweep2_set_plan_j(optype, plan, fc);
  return astr_sweep2_set_plan_k(optype, plan, fc);
}

// Registers an allocation of `nslots` fields (stride L.fstride) as TMA source `which` (0..2).
int astr_sweep2_register_pool(int which, const double* base, int nslots, const Layout& L) {
  if (which < 0 || which >= NPOOLS) return astr_fail_msg("sweep2: bad pool index");
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
  for (int d = 1; d <= 2; ++d)
    for (int pv = 0; pv < 2; ++pv) {
      const int rc = make_map(pm, L, d, pv);
      if (rc) return rc;
    }
  return 0;
}

// Returns -1 when the request is outside what these kernels cover (the caller then uses the
// shared-memory engine of sweep.cu), 0 on success, >0 on error.
int astr_launch_sweep2(int dir, int optype, const LinePlan& plan, const SweepArgs& s, cudaStream_t st) {
  if (!plan.ok || s.epi == EPI_ADD) return -1;
  if (plan.NW < 1 || plan.NW > ASTR_NWMAX) return -1;
  // in place or disjoint only: a bundle is read completely before any of its rows is written
  Sweep2Args a;
  a.L = s.L;
  for (int i = 0; i < ASTR_MAXF; ++i) { a.in[i] = s.in[i]; a.out[i] = s.out[i]; a.slot[i] = 0; }
  a.nf = s.nf; a.epi = s.epi; a.o_lo = s.o_lo; a.o_hi = s.o_hi;
  a.rb = a.nbox = a.sp = 0;
  if (dir == 0) return astr_sweep2_launch_i(optype, a, plan, st);
  // every input field must be a slot of one registered allocation
  const PoolMap* pm = nullptr;
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
  a.rb = pm->rb[dir]; a.nbox = pm->nbox[dir];
  if (dir == 1) return astr_sweep2_launch_j(optype, a, pm->tm[1][0], plan, st);
  return astr_sweep2_launch_k(optype, a, pm->tm[2][0], plan, st);
}
