// Micro-benchmark: how fast can CTAs stream global memory through a TMA-bulk-copy ring?
// variants: chunk bytes, stages, consumer warps, CTAs/SM (via smem padding), plain LDG streaming for reference.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok; const uint32_t a = s32(b);
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(par) : "memory"); } while (!ok);
}
__device__ __forceinline__ void tma1d(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
// each CTA streams `nchunk` chunks starting at its own offset; consumers sum the data (touch every 16 B)
__global__ void ring_kernel(const uint8_t* g, size_t cta_stride, int nchunk, int chunk, int nst, int ncons, double* out, int touch) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncons); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* src = g + (size_t)blockIdx.x * cta_stride;
  double acc = 0;
  if (warp == ncons) {
    for (int i = 0; i < nchunk; ++i) {
      const int st = i % nst, use = i / nst;
      if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
      if (lane == 0) { mbar_expect(&full[st], chunk); tma1d(sm + (size_t)st * chunk, src + (size_t)i * chunk, chunk, &full[st]); }
      __syncwarp();
    }
  } else {
    for (int i = 0; i < nchunk; ++i) {
      const int st = i % nst, use = i / nst;
      mbar_wait(&full[st], use & 1);
      if (touch) {
        const double* p = (const double*)(sm + (size_t)st * chunk);
        for (int k = warp * 32 + lane; k < chunk / 8; k += ncons * 32) acc += p[k];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }
  if (acc == 1.2345) out[0] = acc;
}
// reference: plain coalesced LDG streaming, same footprint per CTA
__global__ void ldg_kernel(const uint8_t* g, size_t cta_stride, size_t bytes, double* out) {
  const double2* p = (const double2*)(g + (size_t)blockIdx.x * cta_stride);
  double acc = 0;
  for (size_t k = threadIdx.x; k < bytes / 16; k += blockDim.x) { double2 v = __ldg(p + k); acc += v.x + v.y; }
  if (acc == 1.2345) out[0] = acc;
}
int main() {
  const size_t total = (size_t)3 << 30;
  uint8_t* g; cudaMalloc(&g, total); cudaMemset(g, 0, total);
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int ncta = getenv("NCTA") ? atoi(getenv("NCTA")) : 1024;
  const size_t stride = (total / ncta) & ~(size_t)16383;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int chunk, nst, ncons, smem_pad_kb, touch, cluster; };
  Cfg cfgs[] = {{16384, 3, 8, 64, 0, 4}, {16384, 6, 8, 64, 0, 4}, {16384, 12, 8, 16, 0, 4}, {16384, 3, 8, 64, 1, 4}, {16384, 12, 8, 16, 1, 4},
                 {32768, 6, 8, 16, 0, 4}, {8192, 12, 8, 64, 0, 4}};
  for (auto& c : cfgs) {
    const int nchunk = (int)(stride / c.chunk);
    const size_t smem = (size_t)c.chunk * c.nst + (size_t)c.smem_pad_kb * 1024;
    const int threads = (c.ncons + 1) * 32;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel, threads, smem);
    auto launch = [&]() {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, ring_kernel, (const uint8_t*)g, stride, nchunk, c.chunk, c.nst, c.ncons, out, c.touch);
    };
    launch();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cluster %d ring chunk %5d stages %2d cons %d smem %3zu KB occ %d touch %d : %.3f ms  %.0f GB/s  err=%s\n", c.cluster, c.chunk, c.nst, c.ncons, smem / 1024, occ, c.touch, ms,
           (double)nchunk * c.chunk * ncta / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  for (int threads : {256, 512, 1024}) {
    ldg_kernel<<<ncta, threads>>>(g, stride, stride, out);
    cudaEventRecord(a);
    ldg_kernel<<<ncta, threads>>>(g, stride, stride, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("ldg threads %d : %.3f ms  %.0f GB/s\n", threads, ms, (double)stride * ncta / ms / 1e6);
  }
  return 0;
}
