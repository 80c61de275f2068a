// Micro-benchmark: bandwidth ONE SM draws through a TMA bulk-copy ring when few / all SMs stream (1 CTA per SM),
// with and without consumers touching the data -- the per-SM ceiling of the stream kernel's shards.
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
int main() {
  const size_t total = (size_t)4 << 30;
  uint8_t* g; cudaMalloc(&g, total); cudaMemset(g, 0, total);
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int chunk = 16384;
  for (int per_cta_mb : {3, 24})
    for (int touch : {0, 1})
      for (int nst : {4, 9, 12})
        for (int ncta : {8, 32, 64, 128, 148}) {
          const size_t stride = (size_t)per_cta_mb << 20;
          const int nchunk = (int)(stride / chunk);
          const size_t smem = (size_t)200 * 1024;   // forces one CTA per SM
          const int ncons = 16, threads = (ncons + 1) * 32;
          auto launch = [&]() { ring_kernel<<<ncta, threads, smem>>>(g, stride, nchunk, chunk, nst, ncons, out, touch); };
          launch(); launch();
          cudaEventRecord(a);
          for (int r = 0; r < 5; ++r) launch();
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
          printf("%2d MB/CTA touch %d stages %2d CTAs %3d: %7.1f us  total %6.0f GB/s  per SM %5.1f GB/s  %s\n", per_cta_mb, touch, nst, ncta, ms * 1e3,
                 (double)stride * ncta / ms / 1e6, (double)stride / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
