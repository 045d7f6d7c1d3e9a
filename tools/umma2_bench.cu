// Micro-benchmark: issue rate of tcgen05.mma.cta_group::2 (M=256 over a CTA pair, K=16, bf16, A and B from
// shared memory; each CTA holds its 128 rows of A and HALF of the N rows of B) next to the cta_group::1
// numbers of tools/umma_bench.cu.  Evidence for the next round (2-CTA conv tiles), not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma2_bench tools/umma2_bench.cu && /tmp/umma2_bench
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
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench2(int N, int ways, int iters, int a_stages, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (a_stages * 16384 + 16384) / 4; i += 128) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    if (rank == 0) {
      // M = 256 (both CTAs' 128 rows), N columns, both operands K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + a_stages * 16384);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t aoff = (uint32_t)((i % a_stages) * 16384);
        for (int k = 0; k < 4; ++k)
          for (int w = 0; w < ways; ++w)
            umma2(tm + (uint32_t)(w * N), make_desc(a0 + aoff) + 2 * k, make_desc(b0) + 2 * k, idesc, 1u);
      }
      asm volatile(
          "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
              smem_u32(&bar)),
          "h"((uint16_t)3)
          : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if (rank == 0) out[blockIdx.x / 2] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int N : {64, 128, 256}) {
    for (int ways : {1, 2, 4}) {
      if (ways * N > 512) continue;
      for (int a_stages : {1, 4}) {
        bench2<<<148, 128, (a_stages * 16 + 16 + 2) * 1024>>>(N, ways, iters, a_stages, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[74];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc = (double)h[0] / (iters * 4.0 * ways);
        printf("cta_group::2 M=256 N=%3d ways=%d a_stages=%d : %.1f cycles/MMA  -> %.0f MAC/cycle/SM (peak 4096)\n", N, ways,
               a_stages, cyc, 128.0 * N * 16 / cyc);
      }
    }
  }
  return 0;
}
