// tools/tma_probe.cu -- micro-benchmark of the data paths the line-solve kernels are built from (B200, sm_100a).
// Copies a large array src -> dst in 4 KB lines through shared memory with (a) bulk-async loads + bulk-async
// stores in a ring of stages, (b) bulk loads + LDS/STG write-out, (c) plain LDG/STG, and prints GB/s (read +
// write).  Answers: how much of the HBM copy bandwidth can one CTA per SM sustain through the bulk engine, and
// with how many bytes in flight.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
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
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// (a) one warp per CTA drives a ring of NS stages of LPS lines each: load stage -> wait -> store stage -> wait read.
// Lines of a stage are issued by the 32 lanes in parallel.
template <int NS>
__global__ void k_bulk_ring(const char* __restrict__ src, char* __restrict__ dst, long long nlines, int lps, int line_bytes) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) unsigned long long full[NS];
  const int lane = threadIdx.x;
  if (lane == 0) for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long long nchunks = (nlines + lps - 1) / lps;
  const size_t stage_bytes = (size_t)lps * line_bytes;
  auto load = [&](long long ch, int s) {
    const long long l0 = ch * lps;
    const int nl = (int)min((long long)lps, nlines - l0);
    if (lane == 0) mbar_expect_tx(&full[s], (unsigned)nl * line_bytes);
    __syncwarp();
    for (int l = lane; l < nl; l += 32)
      bulk_load(smem + s * stage_bytes + (size_t)l * line_bytes, src + (l0 + l) * line_bytes, line_bytes, &full[s]);
  };
  // prologue: fill the ring
  long long ch = blockIdx.x;
  int k = 0;
  for (; k < NS && ch + (long long)k * gridDim.x < nchunks; ++k) load(ch + (long long)k * gridDim.x, k);
  unsigned phase[NS];
  for (int s = 0; s < NS; ++s) phase[s] = 0;
  int s = 0;
  for (long long c = ch; c < nchunks; c += gridDim.x) {
    mbar_wait(&full[s], phase[s]);
    phase[s] ^= 1;
    const long long l0 = c * lps;
    const int nl = (int)min((long long)lps, nlines - l0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int l = lane; l < nl; l += 32)
      bulk_store(dst + (l0 + l) * line_bytes, smem + s * stage_bytes + (size_t)l * line_bytes, line_bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    const long long nxt = c + (long long)NS * gridDim.x;
    if (nxt < nchunks) load(nxt, s);
    s = (s + 1 == NS) ? 0 : s + 1;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (b) bulk loads, write-out by all threads with LDS.128 + STG.128 (NT threads), ring of NS stages
template <int NS>
__global__ void k_bulk_ldsstg(const char* __restrict__ src, char* __restrict__ dst, long long nlines, int lps, int line_bytes) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) unsigned long long full[NS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long nchunks = (nlines + lps - 1) / lps;
  const size_t stage_bytes = (size_t)lps * line_bytes;
  auto load = [&](long long ch, int s) {   // warp 0
    const long long l0 = ch * lps;
    const int nl = (int)min((long long)lps, nlines - l0);
    if (lane == 0) mbar_expect_tx(&full[s], (unsigned)nl * line_bytes);
    __syncwarp();
    for (int l = lane; l < nl; l += 32)
      bulk_load(smem + s * stage_bytes + (size_t)l * line_bytes, src + (l0 + l) * line_bytes, line_bytes, &full[s]);
  };
  long long ch = blockIdx.x;
  if (w == 0) for (int k = 0; k < NS && ch + (long long)k * gridDim.x < nchunks; ++k) load(ch + (long long)k * gridDim.x, k);
  unsigned phase[NS];
  for (int s = 0; s < NS; ++s) phase[s] = 0;
  int s = 0;
  for (long long c = ch; c < nchunks; c += gridDim.x) {
    mbar_wait(&full[s], phase[s]);
    phase[s] ^= 1;
    const long long l0 = c * lps;
    const int nl = (int)min((long long)lps, nlines - l0);
    const size_t nb = (size_t)nl * line_bytes;
    const int4* sp = reinterpret_cast<const int4*>(smem + s * stage_bytes);
    int4* gp = reinterpret_cast<int4*>(dst + l0 * line_bytes);
    for (size_t i = tid; i < nb / 16; i += blockDim.x) gp[i] = sp[i];
    __syncthreads();
    const long long nxt = c + (long long)NS * gridDim.x;
    if (w == 0 && nxt < nchunks) load(nxt, s);
    s = (s + 1 == NS) ? 0 : s + 1;
  }
}

// (c) plain copy
__global__ void k_copy(const int4* __restrict__ src, int4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    int4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i + u * stride < n) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i + u * stride < n) dst[i + u * stride] = v[u];
  }
}

template <class F>
float time_ms(F f, int iters = 5) {
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

int main() {
  const int line_bytes = 4096;
  const long long nlines = 1320000;               // 5.4 GB, as one 5-field sweep
  const size_t bytes = (size_t)nlines * line_bytes;
  char *src, *dst;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes));
  CK(cudaMemset(src, 1, bytes)); CK(cudaMemset(dst, 0, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double gb = 2.0 * bytes / 1e9;
  {
    float ms = time_ms([&] { k_copy<<<sms * 8, 512>>>((const int4*)src, (int4*)dst, bytes / 16); });
    printf("plain LDG/STG copy                          : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  }
  CK(cudaFuncSetAttribute(k_bulk_ring<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ring<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ring<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ring<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ldsstg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(k_bulk_ldsstg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct Cfg { int ns, lps, ctas; };
  const Cfg cfgs[] = {{2, 24, 1}, {2, 12, 1}, {3, 16, 1}, {4, 12, 1}, {6, 8, 1}, {2, 12, 2}, {2, 6, 4}, {4, 6, 2}, {4, 3, 4}};
  for (const Cfg& c : cfgs) {
    const size_t smem = (size_t)c.ns * c.lps * line_bytes;
    float ms = 0;
    auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms * c.ctas, 32, smem>>>(src, dst, nlines, c.lps, line_bytes); }); };
    if (c.ns == 2) run(k_bulk_ring<2>); else if (c.ns == 3) run(k_bulk_ring<3>); else if (c.ns == 4) run(k_bulk_ring<4>); else run(k_bulk_ring<6>);
    printf("bulk load + bulk store  stages=%d lines/stage=%2d CTAs/SM=%d (%3zu KB/SM in ring): %.3f ms  %.0f GB/s\n", c.ns, c.lps,
           c.ctas, smem * c.ctas / 1024, ms, gb / ms * 1e3);
  }
  const Cfg cfg2[] = {{2, 24, 1}, {4, 12, 1}, {2, 12, 2}};
  for (const Cfg& c : cfg2) {
    const size_t smem = (size_t)c.ns * c.lps * line_bytes;
    float ms = 0;
    auto run = [&](auto kern) { ms = time_ms([&] { kern<<<sms * c.ctas, 512 / c.ctas, smem>>>(src, dst, nlines, c.lps, line_bytes); }); };
    if (c.ns == 2) run(k_bulk_ldsstg<2>); else run(k_bulk_ldsstg<4>);
    printf("bulk load + LDS/STG     stages=%d lines/stage=%2d CTAs/SM=%d: %.3f ms  %.0f GB/s\n", c.ns, c.lps, c.ctas, ms, gb / ms * 1e3);
  }
  // verify
  std::vector<char> h(1 << 20);
  CK(cudaMemcpy(h.data(), dst + bytes - h.size(), h.size(), cudaMemcpyDeviceToHost));
  for (char v : h) if (v != 1) { printf("VERIFY FAILED\n"); return 1; }
  printf("verify ok\n");
  return 0;
}
