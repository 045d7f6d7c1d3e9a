// Micro-benchmark: issue rate of tcgen05.mma (M=128, K=16, bf16, A and B from shared memory) for
// different N and different numbers of interleaved accumulators.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_bench tools/umma_bench.cu && /tmp/umma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int N, int ways, int iters, int a_stages, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (a_stages * 16384 + 32768) / 4; i += 128) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + a_stages * 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t aoff = (uint32_t)((i % a_stages) * 16384);
      for (int k = 0; k < 4; ++k)
        for (int w = 0; w < ways; ++w)
          umma(tm + (uint32_t)(w * N), make_desc(a0 + aoff) + 2 * k, make_desc(b0) + 2 * k, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int N : {64, 128, 256}) {
    for (int ways : {1, 2, 4, 8}) {
      if (ways * N > 512) continue;
      for (int a_stages : {1, 4}) {
        bench<<<148, 128, (a_stages * 16 + 32 + 2) * 1024>>>(N, ways, iters, a_stages, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc = (double)h[0] / (iters * 4.0 * ways);
        printf("N=%3d ways=%d a_stages=%d : %.1f cycles/MMA  -> %.0f MAC/cycle/SM (peak 4096)\n", N, ways, a_stages, cyc,
               128.0 * N * 16 / cyc);
      }
    }
  }
  return 0;
}
