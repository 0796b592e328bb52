// tools/tma_probe_jk.cu -- micro-benchmark of the j / k line-bundle data path (B200, sm_100a).
// Copies 5 fields of a padded 513^3 box (pitch 544, 523 x 523 rows: the layout of the 512^3 bench) src -> dst as the
// j / k line-solve kernels would move them: tiles of [ROWS line nodes][PEN pencils] through shared memory, TMA
// tensor loads in, TMA tensor stores out, a ring of NS tiles per CTA driven by one warp.  No arithmetic: this is the
// ceiling of the access pattern (PEN x 8-byte row segments, stride = one row for j, one plane for k) as a function
// of the segment width, the bytes in flight and the CTAs per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe_jk tma_probe_jk.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src)) : "memory");
}

struct Geo { int nbx, nby, nrb, rows, pen, nf, dir, ns, ioff; unsigned tile_bytes; };

// one warp per CTA; lane 0 drives the ring.  STG = 0: TMA store; STG = 1: all lanes copy the tile out with LDS/STG
template <int STG>
__global__ void k_ring(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, Geo g,
                       double* __restrict__ dst, long long sj, long long sk, long long fstride) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) unsigned long long full[40];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) for (int s = 0; s < g.ns; ++s) mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long nbundles = (long long)g.nbx * g.nby * g.nf;
  long long nmine = 0;
  for (long long b = blockIdx.x; b < nbundles; b += gridDim.x) ++nmine;
  const long long nitems = nmine * g.nrb;
  auto coords = [&](long long q, int& c0, int& c1, int& c2, int& c3) {
    const long long b = blockIdx.x + (q / g.nrb) * gridDim.x;
    const int rb = (int)(q % g.nrb);
    const int bx = (int)(b % g.nbx);
    const int by = (int)((b / g.nbx) % g.nby);
    const int f = (int)(b / ((long long)g.nbx * g.nby));
    c0 = g.ioff + bx * g.pen; c3 = f;
    if (g.dir == 1) { c1 = rb * g.rows; c2 = by + 5; } else { c1 = by + 5; c2 = rb * g.rows; }
  };
  auto load = [&](long long q, int s) {
    int c0, c1, c2, c3;
    coords(q, c0, c1, c2, c3);
    mbar_expect_tx(&full[s], g.tile_bytes);
    tma_load_4d(smem + (size_t)s * g.tile_bytes, &tin, c0, c1, c2, c3, &full[s]);
  };
  if (threadIdx.x == 0) for (int k = 0; k < g.ns && k < nitems; ++k) load(k, k);
  int s = 0;
  unsigned phase = 0;
  for (long long q = 0; q < nitems; ++q) {
    mbar_wait(&full[s], phase);
    int c0, c1, c2, c3;
    coords(q, c0, c1, c2, c3);
    if (STG == 0) {
      if (threadIdx.x == 0) {
        tma_store_4d(&tout, c0, c1, c2, c3, smem + (size_t)s * g.tile_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (q + g.ns < nitems) load(q + g.ns, s);
      }
    } else {
      // rows of PEN doubles, one row per (PEN / 2)-lane group with 16-byte accesses
      const int lpr = g.pen / 2;                      // lanes per row
      const int rpi = blockDim.x / lpr;               // rows per iteration
      const int rl = threadIdx.x / lpr, cl = threadIdx.x % lpr;
      const double2* sp = reinterpret_cast<const double2*>(smem + (size_t)s * g.tile_bytes);
      const int total = g.dir == 1 ? 523 : 523;
      const int r0 = (g.dir == 1 ? c1 : c2);
      double* base = dst + (long long)c3 * fstride + c0 + (g.dir == 1 ? (long long)c2 * sk : (long long)c1 * sj);
      const long long rs = g.dir == 1 ? sj : sk;
      for (int r = rl; r < g.rows && r0 + r < total; r += rpi)
        *reinterpret_cast<double2*>(base + (long long)(r0 + r) * rs + 2 * cl) = sp[r * lpr + cl];
      __syncthreads();
      if (threadIdx.x == 0 && q + g.ns < nitems) load(q + g.ns, s);
    }
    if (++s == g.ns) { s = 0; phase ^= 1; }
  }
  if (STG == 0 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  (void)lane;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
float time_ms(F f, int iters = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / iters;
}

int run(int pitch, int njt, bool full_list, int ioff = 16) {
  const int nkt = 523, nf = 5, nn = 513;
  const long long sj = pitch, sk = (long long)pitch * njt, fstride = sk * nkt;
  const size_t bytes = (size_t)fstride * nf * 8;
  double *src, *dst;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes));
  CK(cudaMemset(src, 1, bytes)); CK(cudaMemset(dst, 0, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  CK(cudaFuncSetAttribute(k_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  CK(cudaFuncSetAttribute(k_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  // algorithmic bytes: 513 x 523 lines x 513 pencils, read + written
  const double gb = 2.0 * 8.0 * nf * (double)nn * 523.0 * nn / 1e9;

  struct Cfg { int dir, pen, rows, ns, ctas, stg, threads; };
  std::vector<Cfg> cfgs;
  for (int dir = 1; dir <= 2; ++dir) {
    cfgs.push_back({dir, 32, 175, 3, 1, 1, 512});     // today's structure without the arithmetic: one 3-box tile, LDS/STG
    cfgs.push_back({dir, 32, 88, 9, 1, 0, 32});
    if (!full_list) continue;
    cfgs.push_back({dir, 32, 175, 3, 1, 0, 32});      // the same tile as a ring of its 3 boxes, TMA stores
    cfgs.push_back({dir, 32, 131, 6, 1, 0, 32});
    cfgs.push_back({dir, 32, 88, 4, 2, 0, 32});
    cfgs.push_back({dir, 32, 44, 18, 1, 0, 32});
    cfgs.push_back({dir, 32, 22, 36, 1, 0, 32});
    cfgs.push_back({dir, 16, 175, 6, 1, 0, 32});      // 128-byte segments
    cfgs.push_back({dir, 16, 175, 3, 2, 0, 32});
    cfgs.push_back({dir, 16, 175, 3, 2, 1, 256});
    cfgs.push_back({dir, 64, 88, 4, 1, 0, 32});       // 512-byte segments
    cfgs.push_back({dir, 64, 131, 3, 1, 0, 32});
    cfgs.push_back({dir, 64, 44, 9, 1, 0, 32});
    cfgs.push_back({dir, 8, 176, 6, 2, 0, 32});       // 64-byte segments
  }
  for (const Cfg& c : cfgs) {
    CUtensorMap tin, tout;
    const cuuint64_t gdim[4] = {(cuuint64_t)pitch, (cuuint64_t)njt, (cuuint64_t)nkt, (cuuint64_t)nf};
    const cuuint64_t gstr[3] = {(cuuint64_t)sj * 8, (cuuint64_t)sk * 8, (cuuint64_t)fstride * 8};
    const cuuint32_t box[4] = {(cuuint32_t)c.pen, (cuuint32_t)(c.dir == 1 ? c.rows : 1), (cuuint32_t)(c.dir == 2 ? c.rows : 1), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion prom = c.pen >= 16 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    CUresult r1 = encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, dst, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); continue; }
    Geo g;
    g.ioff = ioff; g.dir = c.dir; g.pen = c.pen; g.rows = c.rows; g.ns = c.ns; g.nf = nf;
    g.nbx = (nn + c.pen - 1) / c.pen;                 // 513 pencils: the ragged bundle is copied too (as the kernels fetch it)
    g.nby = nn; g.nrb = (523 + c.rows - 1) / c.rows;
    g.tile_bytes = (unsigned)c.rows * c.pen * 8;
    const size_t smem = (size_t)g.tile_bytes * c.ns;
    if (smem * c.ctas > 226 * 1024) { printf("skip (smem)\n"); continue; }
    float ms;
    if (c.stg == 0) ms = time_ms([&] { k_ring<0><<<sms * c.ctas, c.threads, smem>>>(tin, tout, g, dst, sj, sk, fstride); });
    else ms = time_ms([&] { k_ring<1><<<sms * c.ctas, c.threads, smem>>>(tin, tout, g, dst, sj, sk, fstride); });
    printf("ioff=%d pitch=%d njt=%d dir=%c pen=%2d (%3d B) rows/tile=%3d stages=%2d CTAs/SM=%d %s (%3zu KB/SM): %.3f ms  %.0f GB/s algorithmic\n",
           ioff, pitch, njt, "ijk"[c.dir], c.pen, c.pen * 8, c.rows, c.ns, c.ctas, c.stg ? "LDS/STG  " : "TMA store", smem * c.ctas / 1024, ms,
           gb / ms * 1e3);
    fflush(stdout);
  }
  // verify the last configuration's copy on a sample
  std::vector<char> h(1 << 16);
  CK(cudaMemcpy(h.data(), (char*)dst + (size_t)(5 * sk + 5 * sj + 16) * 8, 4096, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 4096; ++i) bad += h[i] != 1;
  if (bad) printf("VERIFY FAILED\n");
  CK(cudaFree(src)); CK(cudaFree(dst));
  return bad;
}

int main(int argc, char** argv) {
  if (argc > 2) {            // alignment of the row segments: element offset of node 0 inside a row
    for (int p : {544, 576}) for (int io : {0, 16, 32, 8}) run(p, 523, false, io);
    return 0;
  }
  if (argc > 1) {            // padding scan: row pitch (j stride) and rows per plane (k stride)
    const int pitches[] = {528, 544, 560, 576, 592, 608};
    const int njts[] = {523, 524, 525, 526, 527, 528, 531, 536};
    for (int p : pitches) run(p, 523, false);
    for (int n : njts) if (n != 523) run(544, n, false);
    return 0;
  }
  return run(544, 523, true);
}
