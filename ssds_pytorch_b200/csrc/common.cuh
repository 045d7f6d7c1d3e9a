// Shared host/device helpers for libssdsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/ssdsb200.h"

namespace ssdsb {

// ---- host: error plumbing (never throw across the C ABI) --------------------------------------
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define SSDSB_REQUIRE(cond, ...)                                         \
  do {                                                                   \
    if (!(cond)) return ::ssdsb::fail(SSDSB_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define SSDSB_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::ssdsb::fail(SSDSB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

#define SSDSB_LAUNCH_CHECK(name)                                                           \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::ssdsb::fail(SSDSB_ERR_CUDA, "launch of %s failed: %s", name,                \
                           cudaGetErrorString(_e));                                        \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- device: order-preserving float <-> uint32 map --------------------------------------------
// Larger float  <=> larger unsigned key, for every non-NaN float (negative values included).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
// 64-bit selection key: descending key order == descending score, ascending index among equals.
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t idx) {
  return ((unsigned long long)float_to_ordered(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
  return ordered_to_float((uint32_t)(k >> 32));
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
  return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}

// ---- device: block-wide bitonic sort of 64-bit keys in shared memory, DESCENDING --------------
// n must be a power of two; all threads of the block must call; ends with a barrier.
template <int NT>
__device__ __forceinline__ void block_bitonic_sort_desc(unsigned long long* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += NT) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
        int p = i | j;
        unsigned long long a = s[i], b = s[p];
        bool desc = ((i & k) == 0);
        bool swap = desc ? (a < b) : (a > b);
        if (swap) {
          s[i] = b;
          s[p] = a;
        }
      }
      __syncthreads();
    }
  }
}

// ---- device: running top-K accumulator held in shared memory ----------------------------------
// buf has `cap` slots (power of two).  Callers append keys that exceed *thr (an exclusive lower
// bound) in tiles; whenever fewer than `tile` free slots remain they call topk_prune, which keeps
// the K largest (sorted, descending) and raises *thr to the K-th key.  Keys are unique (they embed
// the index), so "> thr" never drops a needed element.
template <int NT>
__device__ __forceinline__ void topk_prune(unsigned long long* buf, int* s_cnt,
                                           unsigned long long* s_thr, int K) {
  // precondition: barrier passed, *s_cnt stable
  int n = *s_cnt;
  int P = 32;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += NT) buf[i] = 0ull;
  __syncthreads();
  block_bitonic_sort_desc<NT>(buf, P);
  if (threadIdx.x == 0) {
    if (n > K) {
      *s_cnt = K;
      if (buf[K - 1] > *s_thr) *s_thr = buf[K - 1];
    }
  }
  __syncthreads();
}

// warp-aggregated append of up to E candidate keys per thread.  Returns the buffer fill level right
// after this warp's reservation (0 if the warp appended nothing); the maximum over all warps is
// the final fill level, so `__syncthreads_or(ret > limit)` is an exact, race-free overflow test.
template <int E>
__device__ __forceinline__ int topk_append(unsigned long long* buf, int* s_cnt,
                                            const unsigned long long (&keys)[E],
                                            const bool (&take)[E]) {
  int mine = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) mine += take[e] ? 1 : 0;
  const unsigned lane = threadIdx.x & 31;
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (unsigned)d) incl += v;
  }
  int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (total > 0) {
    if (lane == 31) base = atomicAdd(s_cnt, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int pos = base + incl - mine;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (take[e]) buf[pos++] = keys[e];
    return base + total;
  }
  return 0;
}

// Cheaper prune for the streaming phase: block-wide MSB-first radix SELECT of the K-th largest key
// (8-bit digits, early exit once the candidate set is a single key) followed by an unordered
// compaction of the K survivors through `buf2` (K slots) back to buf[0..K).  ~5x fewer instructions
// than sorting the buffer and no per-thread key arrays (registers stay free for the streaming loop);
// the final, ordered result is produced once by topk_prune.  `scratch` needs 260 ints of smem.
// Every thread must call; n = *s_cnt must be stable (a barrier has passed); ends with a barrier.
template <int NT>
__device__ __noinline__ void topk_prune_select(unsigned long long* buf, unsigned long long* buf2,
                                               int* s_cnt, unsigned long long* s_thr, int K,
                                               int* scratch, unsigned long long* s_kth) {
  const int n = *s_cnt;
  if (n <= K) return;  // block-uniform
  const int tid = threadIdx.x;
  unsigned long long prefix = 0ull, mask = 0ull;
  int kk = K;
  bool found = false;
  for (int byte = 7; byte >= 0; --byte) {
    for (int k = tid; k < 256; k += NT) scratch[k] = 0;
    __syncthreads();
    const int shift = byte * 8;
    for (int i = tid; i < n; i += NT) {
      const unsigned long long key = buf[i];
      if ((key & mask) == prefix) atomicAdd(&scratch[(int)((key >> shift) & 255ull)], 1);
    }
    __syncthreads();
    if (tid < 32) {  // lane l owns bins 255-8l .. 248-8l (descending)
      int c[8], sum = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        c[e] = scratch[255 - (tid * 8 + e)];
        sum += c[e];
      }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (tid >= d) incl += v;
      }
      const int excl = incl - sum;
      if (excl < kk && kk <= incl) {
        int run = excl;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (run < kk && kk <= run + c[e]) {
            scratch[256] = 255 - (tid * 8 + e);
            scratch[257] = kk - run;
            scratch[258] = c[e];
          }
          run += c[e];
        }
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)scratch[256] << shift;
    mask |= 255ull << shift;
    kk = scratch[257];
    if (scratch[258] == 1) {  // a single key carries this prefix: it is the K-th largest
      found = true;
      break;
    }
  }
  if (found) {
    for (int i = tid; i < n; i += NT) {
      const unsigned long long key = buf[i];
      if ((key & mask) == prefix) *s_kth = key;
    }
  } else if (tid == 0) {
    *s_kth = prefix;  // all 8 digits fixed
  }
  if (tid == 0) scratch[259] = 0;
  __syncthreads();
  const unsigned long long kth = *s_kth;
  for (int i = tid; i < n; i += NT) {
    const unsigned long long key = buf[i];
    if (key >= kth) buf2[atomicAdd(&scratch[259], 1)] = key;   // exactly K survivors
  }
  __syncthreads();
  for (int i = tid; i < K; i += NT) buf[i] = buf2[i];
  if (tid == 0) {
    *s_cnt = K;
    if (kth > *s_thr) *s_thr = kth;
  }
  __syncthreads();
}

}  // namespace ssdsb
