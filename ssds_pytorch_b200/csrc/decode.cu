// decode: threshold + per-image/level top-k + index->(a,c,y,x) + loc gather + delta2box + clamp +
// centerness rescore, for ALL levels of ALL images in two launches.
//
// reference: ssds/modeling/layers/box.py:408-477 (per level, python loop over the batch) and
// ssds/modeling/layers/decoder.py:36-48 (loop over levels + torch.cat along dim 1).
//
// B200 design (HBM-bound: every score is read exactly once, 128-bit loads):
//   kernel 1  decode_select   grid = (slices per image summed over levels, B).  A CTA streams one
//             slice (<= 64 Ki scores) of one (image, level) score map, keeps a running top-K of
//             64-bit keys (ordered score << 32 | ~flat_index) in shared memory (bitonic prune) and
//             shares its K-th key with the other slices of the same map through a global
//             atomicMax, so late slices reject almost everything with one compare.  Survivors are
//             packed into a per-(image, level) candidate list.
//   kernel 2  decode_finalize grid = (L, B).  Merges the candidate list to the exact top-K (same
//             accumulator), then K threads turn keys into boxes: gather the 4 deltas, add the grid
//             anchor, delta2box (box.py:74-87), clamp, centerness rescore (box.py:464-471), and
//             write straight into the concatenated [B, L*K] outputs (zero padded).
// Order: descending score, ascending flat index among equal scores (torch.topk leaves it open).
#include <stdlib.h>

#include "common.cuh"
#include "decode_emit.cuh"
#include "decode_large.h"

namespace ssdsb {
namespace {

constexpr int DEC_NT = 256;
constexpr int DEC_EPT = 8;                 // two float4 per thread per tile
constexpr int DEC_TILE = DEC_NT * DEC_EPT; // 2048 scores
constexpr int DEC_SLICE = 64 * 1024;       // scores per CTA
constexpr int DEC_MAX_K = 1024;

struct DecodeParams {
  ssdsb_level lv[SSDSB_MAX_LEVELS];
  int slice_begin[SSDSB_MAX_LEVELS + 1];  // prefix sum of slices per image
  int cand_off[SSDSB_MAX_LEVELS + 1];     // prefix sum of candidate capacity per image
  int n_levels, B, K, cap;
  int K_total;   // slots per (image, level) in the outputs (= top_n)
  int out_off;   // first slot written by this round (top_n > 1024 runs several rounds of <= 1024)
  float threshold;
  int rescore;
};

// workspace layout: [gthr: B*L u64][gcnt: B*L i32 (padded)][upper: B*L u64][cand: B*cand_total u64]
// `upper` = exclusive upper key bound of the current round per (image, level): ~0 in round 0, then the
// smallest key emitted by the previous round, 0 once a map is exhausted.
__device__ __forceinline__ unsigned long long* ws_gthr(void* ws) {
  return reinterpret_cast<unsigned long long*>(ws);
}
__device__ __forceinline__ int* ws_gcnt(void* ws, int B, int L) {
  return reinterpret_cast<int*>(reinterpret_cast<unsigned long long*>(ws) + (size_t)B * L);
}
__device__ __host__ __forceinline__ size_t ws_head_bytes(int B, int L) {
  return (size_t)B * L * 8 + (((size_t)B * L * 4 + 7) / 8) * 8;
}
__device__ __forceinline__ unsigned long long* ws_upper(void* ws, int B, int L) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(ws) + ws_head_bytes(B, L));
}
__device__ __forceinline__ unsigned long long* ws_cand(void* ws, int B, int L) {
  return ws_upper(ws, B, L) + (size_t)B * L;
}

__global__ void __launch_bounds__(DEC_NT)
decode_select(const __grid_constant__ DecodeParams p, void* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned long long buf[];  // [p.cap + p.K] (buffer + compaction scratch)
  __shared__ int s_cnt;
  __shared__ unsigned long long s_thr, s_kth;
  __shared__ int s_scratch[260];

  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.slice_begin[l + 1]) ++l;
  const int slice = blockIdx.x - p.slice_begin[l];
  const ssdsb_level& lv = p.lv[l];
  const int n = lv.A * lv.C * lv.H * lv.W;
  const int begin = slice * DEC_SLICE;
  const int end = min(n, begin + DEC_SLICE);
  const float* src = lv.conf + (size_t)b * n;
  const int tid = threadIdx.x;
  const int L = p.n_levels;
  unsigned long long* gthr = ws_gthr(ws) + (size_t)b * L + l;
  const unsigned long long upper = ws_upper(ws, p.B, L)[(size_t)b * L + l];
  if (upper == 0ull) return;                 // this map was exhausted by an earlier round (uniform)
  const float upperF = (upper == ~0ull) ? INFINITY : key_score(upper);

  if (tid == 0) {
    s_cnt = 0;
    s_thr = 0ull;
  }
  __syncthreads();

  const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(lv.conf) & 15) == 0);
  const float thr = p.threshold;
  const int limit = p.cap - DEC_TILE;
  int since_sync = 0;
  if (tid == 0) {   // adopt whatever bound earlier slices of this map already published
    unsigned long long g0 = *reinterpret_cast<volatile unsigned long long*>(gthr);
    if (g0 > s_thr) s_thr = g0;
  }
  __syncthreads();

  // Scan order: the slice is cut into 32 equal runs of 128-byte lines; in every iteration each group
  // of 8 lanes reads the next line of "its" run.  Loads stay fully coalesced (whole lines), but each
  // 1024-element tile is a sample spread over the whole slice (all its channels), so the running
  // K-th score converges after the first tiles instead of being reset by every "record" channel.
  const int len = end - begin;
  const int lines = (len + 31) >> 5;                   // 32 floats per 128-byte line
  const int run = (lines + 31) >> 5;                   // lines per run (32 runs)
  // one iteration = two consecutive lines of this 8-lane group's run (2 x float4 per thread)
  const int iters = vec_ok ? (run + 1) / 2 : (len + DEC_TILE - 1) / DEC_TILE;
  auto load_tile = [&](int it, float (&v)[DEC_EPT], int (&i0)[2]) {
    if (vec_ok) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int lr = 2 * it + h;                     // line inside the run
        const int line = (tid >> 3) * run + lr;
        i0[h] = begin + line * 32 + (tid & 7) * 4;
        if (it < iters && lr < run && line < lines && i0[h] + 4 <= end) {
          float4 q = __ldcs(reinterpret_cast<const float4*>(src + i0[h]));  // streamed once
          v[h * 4 + 0] = q.x; v[h * 4 + 1] = q.y; v[h * 4 + 2] = q.z; v[h * 4 + 3] = q.w;
        } else {
          i0[h] = end;                                 // nothing for this lane in this half
#pragma unroll
          for (int e = 0; e < 4; ++e) v[h * 4 + e] = -INFINITY;
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        i0[h] = begin + it * DEC_TILE + h * (DEC_TILE / 2) + tid * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[h * 4 + e] = (it < iters && i0[h] + e < end) ? __ldcs(src + i0[h] + e) : -INFINITY;
      }
    }
  };
  float vn[DEC_EPT];
  int i0n[2];
  load_tile(0, vn, i0n);
  for (int it = 0; it < iters; ++it) {
    float v[DEC_EPT];
    const int i0[2] = {i0n[0], i0n[1]};
#pragma unroll
    for (int e = 0; e < DEC_EPT; ++e) v[e] = vn[e];
    load_tile(it + 1, vn, i0n);                        // prefetch: four loads in flight per thread
    // fast reject in the float domain; the exact 64-bit key test only runs in warps that have a
    // candidate (rare once the running K-th score has converged)
    const unsigned long long cur = s_thr;
    const float curF = (cur == 0ull) ? -INFINITY : key_score(cur);
    const float lim = fmaxf(thr, curF);
    bool pre = false;
#pragma unroll
    for (int e = 0; e < DEC_EPT; ++e) pre |= (v[e] >= lim) & (v[e] <= upperF);
    int fill = 0;
    if (__any_sync(0xffffffffu, pre)) {
      unsigned long long k[DEC_EPT];
      bool take[DEC_EPT];
#pragma unroll
      for (int e = 0; e < DEC_EPT; ++e) {
        const int idx = i0[e >> 2] + (e & 3);
        k[e] = make_key(v[e], (uint32_t)idx);
        take[e] = (idx < end) && (v[e] >= thr) && (k[e] > cur) && (k[e] < upper);
      }
      fill = topk_append<DEC_EPT>(buf, &s_cnt, k, take);
    }
    if (__syncthreads_or(fill > limit)) {
      topk_prune_select<DEC_NT>(buf, buf + p.cap, &s_cnt, &s_thr, p.K, s_scratch, &s_kth);
      if (tid == 0 && s_cnt >= p.K) {  // publish our K-th key, adopt the best one seen anywhere
        unsigned long long old = atomicMax(gthr, s_thr);
        if (old > s_thr) s_thr = old;
      }
      __syncthreads();
      since_sync = 0;
    } else if (++since_sync == 8) {  // cheap periodic refresh of the shared bound
      if (tid == 0) {
        unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(gthr);
        if (g > s_thr) s_thr = g;
      }
      __syncthreads();
      since_sync = 0;
    }
  }
  __syncthreads();
  topk_prune_select<DEC_NT>(buf, buf + p.cap, &s_cnt, &s_thr, p.K, s_scratch, &s_kth);
  if (tid == 0) {
    if (s_cnt >= p.K) atomicMax(gthr, s_thr);
    s_thr = *reinterpret_cast<volatile unsigned long long*>(gthr);
  }
  __syncthreads();
  // emit (unordered) every key >= the shared bound; the bound itself is somebody's K-th key and must
  // survive.  decode_finalize does the one and only sort.
  const unsigned long long g = s_thr;
  const int cnt = s_cnt;
  __shared__ int s_emit, s_base;
  if (tid == 0) s_emit = 0;
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < cnt; i += DEC_NT) mine += (buf[i] >= g) ? 1 : 0;
  if (mine) atomicAdd(&s_emit, mine);
  __syncthreads();
  const int ne = s_emit;
  if (tid == 0 && ne > 0) s_base = atomicAdd(ws_gcnt(ws, p.B, L) + (size_t)b * L + l, ne);
  if (tid == 0) s_emit = 0;
  __syncthreads();
  if (ne > 0) {
    const int cand_total = p.cand_off[L];
    unsigned long long* dst =
        ws_cand(ws, p.B, L) + (size_t)b * cand_total + p.cand_off[l] + s_base;
    for (int i = tid; i < cnt; i += DEC_NT) {
      const unsigned long long key = buf[i];
      if (key >= g) dst[atomicAdd(&s_emit, 1)] = key;
    }
  }
}

__global__ void __launch_bounds__(DEC_NT)
decode_finalize(const __grid_constant__ DecodeParams p, void* __restrict__ ws,
                float* __restrict__ out_scores, float* __restrict__ out_boxes,
                float* __restrict__ out_classes, int32_t* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned long long buf[];  // [p.cap + p.K]
  __shared__ int s_cnt;
  __shared__ unsigned long long s_thr, s_kth;
  __shared__ int s_scratch[260];

  const int l = blockIdx.x, b = blockIdx.y;
  const int L = p.n_levels, K = p.K;
  const int tid = threadIdx.x;
  const ssdsb_level& lv = p.lv[l];
  const int cand_total = p.cand_off[L];
  const unsigned long long* cand =
      ws_cand(ws, p.B, L) + (size_t)b * cand_total + p.cand_off[l];
  const int ncand = ws_gcnt(ws, p.B, L)[(size_t)b * L + l];

  if (tid == 0) {
    s_cnt = 0;
    s_thr = 0ull;
  }
  __syncthreads();
  const int limit = p.cap - DEC_TILE;
  for (int base = 0; base < ncand; base += DEC_TILE) {
    unsigned long long k[DEC_EPT];
    bool take[DEC_EPT];
    const unsigned long long cur = s_thr;
#pragma unroll
    for (int e = 0; e < DEC_EPT; ++e) {
      int i = base + e * DEC_NT + tid;
      k[e] = (i < ncand) ? cand[i] : 0ull;
      take[e] = (i < ncand) && (k[e] > cur);
    }
    const int fill = topk_append<DEC_EPT>(buf, &s_cnt, k, take);
    if (__syncthreads_or(fill > limit))
      topk_prune_select<DEC_NT>(buf, buf + p.cap, &s_cnt, &s_thr, K, s_scratch, &s_kth);
  }
  __syncthreads();
  topk_prune_select<DEC_NT>(buf, buf + p.cap, &s_cnt, &s_thr, K, s_scratch, &s_kth);  // -> <= K keys
  topk_prune<DEC_NT>(buf, &s_cnt, &s_thr, K);                                       // one small sort
  const int nout = min(s_cnt, K);

  const size_t row = (size_t)b * L * p.K_total + (size_t)l * p.K_total + p.out_off;
  if (tid == 0)   // bound for the next round: the smallest key emitted now, or "exhausted"
    ws_upper(ws, p.B, L)[(size_t)b * L + l] = (nout == K) ? buf[K - 1] : 0ull;

  for (int t = tid; t < K; t += DEC_NT)
    emit_detection(lv, b, t < nout, t < nout ? buf[t] : 0ull, p.rescore, row + t, out_scores, out_boxes,
                   out_classes, out_index);
}

int fill_params(DecodeParams& p, const ssdsb_level* levels, int n_levels, int B, int top_n_total) {
  const int top_n = top_n_total < DEC_MAX_K ? top_n_total : DEC_MAX_K;   // per round
  p.n_levels = n_levels;
  p.B = B;
  p.K = top_n;
  p.K_total = top_n_total;
  p.out_off = 0;
  p.cap = next_pow2(top_n + DEC_TILE + 1);   // prune when fewer than one tile of slots is left
  if (p.cap < 2 * DEC_TILE) p.cap = 2 * DEC_TILE;
  p.slice_begin[0] = 0;
  p.cand_off[0] = 0;
  for (int l = 0; l < n_levels; ++l) {
    p.lv[l] = levels[l];
    long long n = (long long)levels[l].A * levels[l].C * levels[l].H * levels[l].W;
    int slices = (int)((n + DEC_SLICE - 1) / DEC_SLICE);
    if (slices < 1) slices = 1;
    p.slice_begin[l + 1] = p.slice_begin[l] + slices;
    p.cand_off[l + 1] = p.cand_off[l] + slices * top_n;
  }
  return 0;
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

static int validate_levels(const ssdsb_level* levels, int n_levels, int B, int top_n,
                           bool need_ptrs) {
  SSDSB_REQUIRE(levels != nullptr, "decode: levels is NULL");
  SSDSB_REQUIRE(n_levels >= 1 && n_levels <= SSDSB_MAX_LEVELS, "decode: n_levels=%d outside [1,%d]",
                n_levels, SSDSB_MAX_LEVELS);
  SSDSB_REQUIRE(B >= 0, "decode: negative batch");
  SSDSB_REQUIRE(top_n >= 1, "decode: top_n=%d must be >= 1", top_n);
  if (top_n > 64 * DEC_MAX_K)
    return fail(SSDSB_ERR_UNSUPPORTED, "decode: top_n=%d > %d not implemented", top_n, 64 * DEC_MAX_K);
  for (int l = 0; l < n_levels; ++l) {
    const ssdsb_level& v = levels[l];
    SSDSB_REQUIRE(v.A >= 1 && v.C >= 1 && v.H >= 1 && v.W >= 1 && v.stride >= 1,
                  "decode: level %d has a non-positive dimension", l);
    SSDSB_REQUIRE((long long)v.A * v.C * v.H * v.W < (1ll << 31),
                  "decode: level %d has too many scores per image", l);
    if (need_ptrs) {
      SSDSB_REQUIRE(v.conf && v.loc && v.anchors, "decode: level %d has a NULL pointer", l);
      SSDSB_REQUIRE(((uintptr_t)v.anchors & 15) == 0, "decode: anchors must be 16-byte aligned");
    }
  }
  return SSDSB_OK;
}

extern "C" size_t ssdsb_decode_workspace_bytes(const ssdsb_level* levels, int n_levels, int B,
                                               int top_n) {
  if (validate_levels(levels, n_levels, B, top_n, false) != SSDSB_OK) return 0;
  if (top_n > DEC_MAX_K && top_n <= decode_large_max_k()) return decode_large_workspace_bytes(levels, n_levels, B, top_n);
  DecodeParams p;
  fill_params(p, levels, n_levels, B, top_n);
  return ws_head_bytes(B, n_levels) + (size_t)B * n_levels * 8 + (size_t)B * p.cand_off[n_levels] * 8 + 16;
}

extern "C" int ssdsb_decode(const ssdsb_level* levels, int n_levels, int B, float threshold,
                            int top_n, int rescore, float* d_scores, float* d_boxes,
                            float* d_classes, int32_t* d_index, void* d_workspace,
                            size_t workspace_bytes, void* stream) {
  int rc = validate_levels(levels, n_levels, B, top_n, true);
  if (rc != SSDSB_OK) return rc;
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_scores && d_boxes && d_classes, "decode: NULL output");
  SSDSB_REQUIRE(((uintptr_t)d_boxes & 15) == 0, "decode: boxes output must be 16-byte aligned");
  const size_t need = ssdsb_decode_workspace_bytes(levels, n_levels, B, top_n);
  if (!d_workspace || workspace_bytes < need || ((uintptr_t)d_workspace & 15) != 0)
    return fail(SSDSB_ERR_WORKSPACE, "decode: workspace %zu B given, %zu B (16-byte aligned) needed",
                workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  // 1024 < top_n <= 24576: three streaming passes + one shared-memory sort per (image, level) (decode_large.cu)
  // instead of one pass over the score maps per 1024 results.  SSDSB_DECODE_ROUNDS=1 keeps the rounds (A/B runs).
  if (top_n > DEC_MAX_K && top_n <= decode_large_max_k() && !getenv("SSDSB_DECODE_ROUNDS"))
    return decode_large(levels, n_levels, B, threshold, top_n, rescore, d_scores, d_boxes, d_classes, d_index,
                        d_workspace, workspace_bytes, st);
  DecodeParams p;
  fill_params(p, levels, n_levels, B, top_n);
  p.threshold = threshold;
  p.rescore = rescore;
  const size_t head = ws_head_bytes(B, n_levels);
  // upper bounds start at "none"
  SSDSB_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned char*>(d_workspace) + head, 0xff,
                             (size_t)B * n_levels * 8, st));
  const size_t smem = (size_t)(p.cap + p.K) * sizeof(unsigned long long);
  SSDSB_CUDA(cudaFuncSetAttribute(decode_select, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  SSDSB_CUDA(cudaFuncSetAttribute(decode_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  dim3 g1(p.slice_begin[n_levels], B);
  dim3 g2(n_levels, B);
  // top_n <= 1024: one round.  Larger top_n (the 20 000-per-level NMS stress of SURVEY 8d cfg 5):
  // successive rounds each extract the next <= 1024 keys below the previous round's smallest key.
  const int per_round = p.K;
  for (int off = 0; off < top_n; off += per_round) {
    p.K = (top_n - off) < per_round ? (top_n - off) : per_round;
    p.out_off = off;
    SSDSB_CUDA(cudaMemsetAsync(d_workspace, 0, head, st));
    decode_select<<<g1, DEC_NT, smem, st>>>(p, d_workspace);
    SSDSB_LAUNCH_CHECK("decode_select");
    decode_finalize<<<g2, DEC_NT, smem, st>>>(p, d_workspace, d_scores, d_boxes, d_classes, d_index);
    SSDSB_LAUNCH_CHECK("decode_finalize");
  }
  return SSDSB_OK;
}
