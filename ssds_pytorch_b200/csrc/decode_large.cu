// Exact top-K of long score rows with THREE streaming passes and a parallel sort — used by
//   * decode for LARGE top_n (1024 < top_n <= 65536 per level; the 20 000-per-level NMS stress of SURVEY 8d cfg 5),
//     instead of one pass over the score maps per 1024 results, and
//   * nms for long candidate rows (N > 8192): the top-2048 keys of each image in sorted order, instead of one CTA
//     streaming and pruning all N scores.
//
// reference: ssds/modeling/layers/box.py:408-477 (keep = conf >= thr; topk(min(top_n, |keep|)) sorted descending;
// index -> (a, c, y, x); delta2box; rescore) and :496-505 (drop score <= 0, sort descending) — same outputs, same
// order (descending score, ascending flat index among equal scores) as decode.cu / nms.cu.
//
// Keys are ordered-float(score) (32 bits) + flat index (unique).  A 3-level MSD radix select over the score bits,
// fused with emission so that a row is never read more than three times:
//   pass 1  histogram of bits [31:21] of every score >= thr                        -> b1 = bin of the K-th key
//   pass 2  keys with bin1 > b1 -> "sure" list; histogram of bits [20:10] inside bin b1 -> b2
//   pass 3  inside bin b1: bin2 > b2 -> "sure" list; bin2 == b2 -> "maybe" list
//           (sure has < K keys; sure + maybe >= K; the maybe keys share their top 22 score bits: a handful)
//   dl_degenerate   only when more than DL_MCAP keys share those 22 bits (constant score maps): an exact but slow
//                   single-CTA re-scan of that one row selects on the last 10 bits, then the lowest flat indices
//   dl_sort_chunks  every 2048-key chunk of [sure | maybe] is sorted by its own CTA (bitonic, shared memory)
//   dl_merge        rank of a key = its rank in its chunk + sum over the other chunks of the keys above it
//                   (binary search, L2); rank < K -> the key is result number `rank`: decode emits the detection
//                   row directly (box.py:443-471), nms writes the sorted key list.
// The scan after each histogram is done by the LAST CTA of the row (threadfence + ticket), emitted keys are staged in
// shared memory and flushed with one global reservation per CTA, so the whole selection is memset + 6 launches, no
// host sync, CUDA-graph capturable.  Histogram updates are warp-aggregated with match.any: sigmoid scores of
// neighbouring pixels share their exponent and top mantissa bits, which would serialise plain shared-memory atomics.
#include "decode_emit.cuh"
#include "decode_large.h"

namespace ssdsb {
namespace {

constexpr int DL_NT = 256;
constexpr int DL_SLICE = 16 * 1024;      // scores per CTA in the streaming passes
constexpr int DL_BINS = 2048;
constexpr int DL_MCAP = 8192;            // maybe-list capacity per row
constexpr int DL_CHUNK = 2048;           // keys sorted per CTA
constexpr int DL_SORT_NT = 1024;
constexpr int DL_STAGE = 3072;           // keys staged per CTA before a flush (>= one iteration's worth + slack)

struct DlSeg {      // per row (image, level), in the workspace
  int b1, above1;   // bin of the K-th key among bits [31:21]; keys in higher bins
  int b2, above2;   // same inside bin b1 for bits [20:10]
  int n_sure, n_maybe;
  int ticket1, ticket2;
  int pad[8];
};

struct DlRow {
  const float* scores;      // row 0 of this "level"; image b's row starts at scores + b * n
  int n;
};

struct DlParams {
  DlRow row[SSDSB_MAX_LEVELS];
  ssdsb_level lv[SSDSB_MAX_LEVELS];        // decode mode only (emission)
  int slice_begin[SSDSB_MAX_LEVELS + 1];
  int n_levels, B, K, n_chunks;
  float threshold;
  int rescore;
  DlSeg* seg;                 // [B*L]
  int* hist1;                 // [B*L][DL_BINS]
  int* hist2;                 // [B*L][DL_BINS]
  unsigned long long* cand;   // [B*L][K + DL_MCAP]: sure keys at [0, n_sure), maybe keys at [K, K + n_maybe)
  unsigned long long* sorted; // [B*L][n_chunks * DL_CHUNK]
};

__device__ __forceinline__ void hist_add_warp(int* s_hist, int bin, bool valid) {
  // warp-aggregated: lanes with the same bin elect a leader that adds their count
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned peers = __match_any_sync(act, bin);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_hist[bin], __popc(peers));
  }
}

// the last CTA of a row: find the bin d (from the top) where the running count reaches `want`;
// (d, keys in bins above d) -> s_out[0], s_out[1]; d = -1 when the row holds fewer than `want` keys.
__device__ __forceinline__ void scan_bins(const int* __restrict__ ghist, int want, int* s_out) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int rem = want, d = 0, above = 0;
    bool found = false;
    for (int base = DL_BINS - 32; base >= 0 && !found; base -= 32) {
      const int cnt = __ldcg(ghist + base + 31 - lane);        // lane 0 = highest bin of the chunk
      int inc = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, inc >= rem);
      if (hit) {
        const int hl = __ffs(hit) - 1;
        const int before = __shfl_sync(0xffffffffu, inc, hl) - __shfl_sync(0xffffffffu, cnt, hl);
        d = base + 31 - hl;
        above += before;
        found = true;
      } else {
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        rem -= tot;
        above += tot;
      }
    }
    if (lane == 0) {
      s_out[0] = found ? d : -1;
      s_out[1] = above;
    }
  }
  __syncthreads();
}

// PASS: 1 = hist1; 2 = emit(bin1 > b1) + hist2 inside b1; 3 = emit inside b1 (bin2 > b2 -> sure, == b2 -> maybe)
template <int PASS>
__global__ void __launch_bounds__(DL_NT)
dl_pass(const __grid_constant__ DlParams p) {
  __shared__ int s_hist[PASS <= 2 ? DL_BINS : 1];
  __shared__ unsigned long long s_keys[PASS >= 2 ? DL_STAGE : 1];
  __shared__ int s_out[2];
  __shared__ int s_flag, s_nk, s_base;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.slice_begin[l + 1]) ++l;
  const int slice = blockIdx.x - p.slice_begin[l];
  const int L = p.n_levels;
  const int sg = b * L + l;
  DlSeg* seg = p.seg + sg;
  const int n = p.row[l].n;
  const int begin = slice * DL_SLICE;
  const int end = min(n, begin + DL_SLICE);
  const float* src = p.row[l].scores + (size_t)b * n;
  const float thr = p.threshold;
  int b1 = 0, b2 = 0;
  if (PASS >= 2) {
    b1 = __ldcg(&seg->b1);
    if (PASS == 2 && b1 < 0) return;          // fewer than K keys pass: pass 3 takes them all, no hist2 needed
    if (PASS == 3) b2 = __ldcg(&seg->b2);
  }
  if (PASS <= 2) {
    for (int k = tid; k < DL_BINS; k += DL_NT) s_hist[k] = 0;
  }
  if (tid == 0) s_nk = 0;
  __syncthreads();
  unsigned long long* sure = p.cand + (size_t)sg * (p.K + DL_MCAP);
  unsigned long long* maybe = sure + p.K;
  const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.row[l].scores) & 15) == 0);

  // sure keys are staged in shared memory; one global reservation per flush
  auto flush = [&]() {          // all threads; s_nk stable (a barrier has passed)
    const int cnt = s_nk;
    if (cnt == 0) return;       // block-uniform
    if (tid == 0) s_base = atomicAdd(&seg->n_sure, cnt);
    __syncthreads();
    const int base = s_base;
    for (int i = tid; i < cnt; i += DL_NT) sure[base + i] = s_keys[i];
    __syncthreads();
    if (tid == 0) s_nk = 0;
    __syncthreads();
  };
  auto stage = [&](bool take, unsigned long long key) {      // full-warp call
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_nk, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) s_keys[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
  };
  auto handle = [&](float v, int idx, bool inb) {
    const bool ok = inb && (v >= thr);
    const uint32_t ord = float_to_ordered(v);
    const int bin1 = (int)(ord >> 21);
    if (PASS == 1) {
      hist_add_warp(s_hist, bin1, ok);
    } else if (PASS == 2) {
      hist_add_warp(s_hist, (int)((ord >> 10) & 2047u), ok && bin1 == b1);
      stage(ok && bin1 > b1, make_key(v, (uint32_t)idx));
    } else {
      const int bin2 = (int)((ord >> 10) & 2047u);
      // b1 < 0: fewer than K keys pass the threshold -> all of them are results
      stage(ok && (b1 < 0 || (bin1 == b1 && bin2 > b2)), make_key(v, (uint32_t)idx));
      const bool m_ = ok && b1 >= 0 && bin1 == b1 && bin2 == b2;
      const unsigned mm = __ballot_sync(0xffffffffu, m_);
      if (mm) {                                             // rare: a handful of keys per row
        int base = 0;
        if (lane == 0) base = atomicAdd(&seg->n_maybe, __popc(mm));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int o = base + __popc(mm & ((1u << lane) - 1u));
        if (m_ && o < DL_MCAP) maybe[o] = make_key(v, (uint32_t)idx);      // overflow: dl_degenerate re-scans
      }
    }
  };

  // warp-uniform trip counts (the handlers use full-mask ballots); 4 x 16-byte loads per thread in flight
  if (vec_ok) {
    const int iters = (end - begin + DL_NT * 16 - 1) / (DL_NT * 16);
    for (int it = 0; it < iters; ++it) {
      float4 q[4];
      int i0[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        i0[u] = begin + (it * 4 + u) * DL_NT * 4 + tid * 4;
        in[u] = i0[u] + 4 <= end;
        q[u] = in[u] ? __ldcs(reinterpret_cast<const float4*>(src + i0[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        handle(q[u].x, i0[u] + 0, in[u]); handle(q[u].y, i0[u] + 1, in[u]);
        handle(q[u].z, i0[u] + 2, in[u]); handle(q[u].w, i0[u] + 3, in[u]);
        if (PASS >= 2 && (u & 1)) {                           // <= 2048 keys staged since the last check
          __syncthreads();
          if (s_nk > DL_STAGE - 2048) flush();                // block-uniform
        }
      }
    }
  } else {
    const int iters = (end - begin + DL_NT - 1) / DL_NT;
    for (int it = 0; it < iters; ++it) {
      const int i0 = begin + it * DL_NT + tid;
      const bool inb = i0 < end;
      handle(inb ? __ldcs(src + i0) : 0.0f, i0, inb);
      if (PASS >= 2 && (it & 7) == 7) {
        __syncthreads();
        if (s_nk > DL_STAGE - 2048) flush();
      }
    }
  }
  __syncthreads();
  if (PASS >= 2) flush();
  if (PASS == 3) return;

  // flush the CTA histogram, then the last CTA of the row scans it
  int* ghist = (PASS == 1 ? p.hist1 : p.hist2) + (size_t)sg * DL_BINS;
  for (int k = tid; k < DL_BINS; k += DL_NT) {
    const int c = s_hist[k];
    if (c) atomicAdd(ghist + k, c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int slices = p.slice_begin[l + 1] - p.slice_begin[l];
    s_flag = (atomicAdd(PASS == 1 ? &seg->ticket1 : &seg->ticket2, 1) == slices - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (PASS == 1) {
    scan_bins(ghist, p.K, s_out);
    if (tid == 0) {
      seg->b1 = s_out[0];
      seg->above1 = s_out[0] >= 0 ? s_out[1] : 0;
    }
  } else {
    const int want = p.K - __ldcg(&seg->above1);
    scan_bins(ghist, want, s_out);
    if (tid == 0) {
      seg->b2 = s_out[0];                // >= 0: bin b1 holds at least `want` keys
      seg->above2 = s_out[1];
    }
  }
}

// Degenerate rows only: more than DL_MCAP keys share the top 22 score bits with the K-th key.  One CTA re-scans the
// row: selects on the last 10 score bits, then takes the lowest flat indices among the keys equal to the boundary
// score, and appends exactly the missing keys to the sure list (which then holds K keys).
__global__ void __launch_bounds__(DL_SORT_NT, 1)
dl_degenerate(const __grid_constant__ DlParams p) {
  __shared__ int s_hist[1024];
  __shared__ int s_cnt, s_run, s_b3, s_r2;
  __shared__ int s_warp[DL_SORT_NT / 32];
  const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int sg = b * p.n_levels + l;
  DlSeg* seg = p.seg + sg;
  const int b1 = __ldcg(&seg->b1);
  if (b1 < 0 || __ldcg(&seg->n_maybe) <= DL_MCAP) return;
  const int K = p.K;
  const int n_sure = __ldcg(&seg->n_sure);
  const int need = K - n_sure;                           // > 0 keys still missing, all inside (b1, b2)
  const uint32_t prefix = ((uint32_t)b1 << 11) | (uint32_t)__ldcg(&seg->b2);
  const int nn = p.row[l].n;
  const float* src = p.row[l].scores + (size_t)b * nn;
  unsigned long long* sure = p.cand + (size_t)sg * (K + DL_MCAP);
  for (int k = tid; k < 1024; k += DL_SORT_NT) s_hist[k] = 0;
  __syncthreads();
  for (int i = tid; i < nn; i += DL_SORT_NT) {
    const float v = __ldg(src + i);
    const uint32_t ord = float_to_ordered(v);
    if (v >= p.threshold && (ord >> 10) == prefix) atomicAdd(&s_hist[ord & 1023u], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int rem = need, d = 1023;
    for (; d > 0; --d) {
      if (s_hist[d] >= rem) break;
      rem -= s_hist[d];
    }
    s_b3 = d;
    s_r2 = rem;                 // keys equal to the boundary score to take, lowest flat index first
    s_cnt = n_sure;
    s_run = 0;
  }
  __syncthreads();
  const uint32_t b3 = (uint32_t)s_b3;
  const int r2 = s_r2;
  for (int base = 0; base < nn; base += DL_SORT_NT) {
    const int i = base + tid;
    const float v = (i < nn) ? __ldg(src + i) : -1.0f;
    const uint32_t ord = float_to_ordered(v);
    const bool inp = (i < nn) && v >= p.threshold && (ord >> 10) == prefix;
    const bool gt = inp && (ord & 1023u) > b3;
    const bool eq = inp && (ord & 1023u) == b3;
    const unsigned me = __ballot_sync(0xffffffffu, eq);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) s_warp[wid] = __popc(me);
    __syncthreads();
    int before = s_run;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int rank = before + __popc(me & ((1u << lane) - 1u));       // 0-based among equals, index order
    if (gt || (eq && rank < r2)) sure[atomicAdd(&s_cnt, 1)] = make_key(v, (uint32_t)i);
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < DL_SORT_NT / 32; ++w) t += s_warp[w];
      s_run += t;
    }
    __syncthreads();
  }
  if (tid == 0) {
    seg->n_sure = s_cnt;         // == K
    seg->n_maybe = 0;
  }
}

// logical candidate list of a row = [sure | maybe]; sorts chunk blockIdx.x of it (descending) into `sorted`
__global__ void __launch_bounds__(DL_SORT_NT, 1)
dl_sort_chunks(const __grid_constant__ DlParams p) {
  __shared__ unsigned long long s[DL_CHUNK];
  const int c = blockIdx.x, l = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int sg = b * p.n_levels + l;
  const DlSeg* seg = p.seg + sg;
  const int K = p.K;
  const int n_sure = min(__ldcg(&seg->n_sure), K);
  const int n_maybe = (__ldcg(&seg->b1) >= 0) ? min(__ldcg(&seg->n_maybe), DL_MCAP) : 0;
  const int n = n_sure + n_maybe;
  const int lo = c * DL_CHUNK;
  if (lo >= n) return;
  const unsigned long long* cand = p.cand + (size_t)sg * (K + DL_MCAP);
  for (int i = tid; i < DL_CHUNK; i += DL_SORT_NT) {
    const int q = lo + i;
    s[i] = (q < n) ? __ldcg(cand + (q < n_sure ? q : K + (q - n_sure))) : 0ull;   // 0 < every real key
  }
  __syncthreads();
  // bitonic network, descending; one compare-exchange per thread and step (DL_CHUNK == 2 * DL_SORT_NT)
  for (int k = 2; k <= DL_CHUNK; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
      const int q = i | j;
      const unsigned long long a = s[i], d = s[q];
      const bool desc = ((i & k) == 0);
      if (desc ? (a < d) : (a > d)) {
        s[i] = d;
        s[q] = a;
      }
      __syncthreads();
    }
  }
  unsigned long long* out = p.sorted + ((size_t)sg * p.n_chunks + c) * DL_CHUNK;
  for (int i = tid; i < DL_CHUNK; i += DL_SORT_NT) out[i] = s[i];
}

// number of keys in the descending array a[0..m) that are greater than x
__device__ __forceinline__ int count_greater(const unsigned long long* __restrict__ a, int m, unsigned long long x) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldcg(a + mid) > x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// MODE 0: decode — result number `rank` is turned into a detection row; MODE 1: the sorted key list itself
template <int MODE>
__global__ void __launch_bounds__(DL_SORT_NT, 1)
dl_merge(const __grid_constant__ DlParams p, float* __restrict__ out_scores, float* __restrict__ out_boxes,
         float* __restrict__ out_classes, int32_t* __restrict__ out_index, unsigned long long* __restrict__ out_keys,
         int* __restrict__ out_count) {
  const int c = blockIdx.x, l = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int L = p.n_levels, K = p.K;
  const int sg = b * L + l;
  const DlSeg* seg = p.seg + sg;
  const int n_sure = min(__ldcg(&seg->n_sure), K);
  const int n_maybe = (__ldcg(&seg->b1) >= 0) ? min(__ldcg(&seg->n_maybe), DL_MCAP) : 0;
  const int n = n_sure + n_maybe;
  const int nout = min(n, K);
  const int nch = (n + DL_CHUNK - 1) / DL_CHUNK;
  const unsigned long long* sorted = p.sorted + (size_t)sg * p.n_chunks * DL_CHUNK;
  const size_t row = (size_t)sg * K;
  if (MODE == 1 && c == 0 && tid == 0) out_count[sg] = nout;
  // zero padding of the result slots nobody owns
  for (int t = c * DL_CHUNK + tid; t < min(K, (c + 1) * DL_CHUNK); t += DL_SORT_NT) {
    if (t >= nout) {
      if (MODE == 0) emit_detection(p.lv[l], b, false, 0ull, p.rescore, row + t, out_scores, out_boxes, out_classes,
                                    out_index);
      else out_keys[row + t] = 0ull;
    }
  }
  if (c >= nch) return;
  const int mine = min(DL_CHUNK, n - c * DL_CHUNK);
  for (int i = tid; i < mine; i += DL_SORT_NT) {
    const unsigned long long x = __ldcg(sorted + (size_t)c * DL_CHUNK + i);
    int rank = i;
    for (int o = 0; o < nch; ++o) {
      if (o == c) continue;
      rank += count_greater(sorted + (size_t)o * DL_CHUNK, min(DL_CHUNK, n - o * DL_CHUNK), x);
    }
    if (rank < K) {
      if (MODE == 0) emit_detection(p.lv[l], b, true, x, p.rescore, row + rank, out_scores, out_boxes, out_classes,
                                    out_index);
      else out_keys[row + rank] = x;
    }
  }
}

struct DlLayout {
  size_t seg, hist1, hist2, zero_bytes, cand, sorted, total;
  int n_chunks;
};

DlLayout dl_layout(int B, int L, int K) {
  DlLayout w;
  const size_t segs = (size_t)B * L;
  w.n_chunks = (K + DL_MCAP + DL_CHUNK - 1) / DL_CHUNK;
  size_t o = 0;
  w.seg = o; o += align_up(segs * sizeof(DlSeg), 256);
  w.hist1 = o; o += align_up(segs * DL_BINS * 4, 256);
  w.hist2 = o; o += align_up(segs * DL_BINS * 4, 256);
  w.zero_bytes = o;               // everything up to here is zeroed per call
  w.cand = o; o += align_up(segs * (size_t)(K + DL_MCAP) * 8, 256);
  w.sorted = o; o += align_up(segs * (size_t)w.n_chunks * DL_CHUNK * 8, 256);
  w.total = o;
  return w;
}

int dl_select(DlParams& p, const DlLayout& w, unsigned char* ws, cudaStream_t st) {
  p.seg = reinterpret_cast<DlSeg*>(ws + w.seg);
  p.hist1 = reinterpret_cast<int*>(ws + w.hist1);
  p.hist2 = reinterpret_cast<int*>(ws + w.hist2);
  p.cand = reinterpret_cast<unsigned long long*>(ws + w.cand);
  p.sorted = reinterpret_cast<unsigned long long*>(ws + w.sorted);
  p.n_chunks = w.n_chunks;
  p.slice_begin[0] = 0;
  for (int l = 0; l < p.n_levels; ++l) {
    int slices = (p.row[l].n + DL_SLICE - 1) / DL_SLICE;
    if (slices < 1) slices = 1;
    p.slice_begin[l + 1] = p.slice_begin[l] + slices;
  }
  SSDSB_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
  const dim3 g(p.slice_begin[p.n_levels], p.B);
  dl_pass<1><<<g, DL_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_pass<1>");
  dl_pass<2><<<g, DL_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_pass<2>");
  dl_pass<3><<<g, DL_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_pass<3>");
  dl_degenerate<<<dim3(p.n_levels, p.B), DL_SORT_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_degenerate");
  dl_sort_chunks<<<dim3(p.n_chunks, p.n_levels, p.B), DL_SORT_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_sort_chunks");
  return SSDSB_OK;
}

}  // namespace

int decode_large_max_k() { return 64 * 1024; }

size_t decode_large_workspace_bytes(int n_levels, int B, int top_n) { return dl_layout(B, n_levels, top_n).total + 256; }

int decode_large(const ssdsb_level* levels, int n_levels, int B, float threshold, int top_n, int rescore,
                 float* d_scores, float* d_boxes, float* d_classes, int32_t* d_index, void* d_workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  const DlLayout w = dl_layout(B, n_levels, top_n);
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  if (!d_workspace || workspace_bytes < w.total + 256)
    return fail(SSDSB_ERR_WORKSPACE, "decode: workspace %zu B given, %zu B needed", workspace_bytes, w.total + 256);
  DlParams p;
  p.n_levels = n_levels; p.B = B; p.K = top_n;
  p.threshold = threshold; p.rescore = rescore;
  for (int l = 0; l < n_levels; ++l) {
    p.lv[l] = levels[l];
    p.row[l].scores = levels[l].conf;
    p.row[l].n = levels[l].A * levels[l].C * levels[l].H * levels[l].W;
  }
  int rc = dl_select(p, w, ws, st);
  if (rc != SSDSB_OK) return rc;
  dl_merge<0><<<dim3(p.n_chunks, n_levels, B), DL_SORT_NT, 0, st>>>(p, d_scores, d_boxes, d_classes, d_index, nullptr,
                                                                   nullptr);
  SSDSB_LAUNCH_CHECK("dl_merge<0>");
  return SSDSB_OK;
}

size_t topk_rows_workspace_bytes(int B, int K) { return dl_layout(B, 1, K).total + 256; }

int topk_rows(const float* d_scores, int B, int N, float min_score, int K, unsigned long long* d_keys, int* d_count,
              void* d_workspace, size_t workspace_bytes, cudaStream_t st) {
  const DlLayout w = dl_layout(B, 1, K);
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  if (!d_workspace || workspace_bytes < w.total + 256)
    return fail(SSDSB_ERR_WORKSPACE, "topk_rows: workspace %zu B given, %zu B needed", workspace_bytes, w.total + 256);
  DlParams p;
  p.n_levels = 1; p.B = B; p.K = K;
  p.threshold = min_score; p.rescore = 0;
  p.row[0].scores = d_scores;
  p.row[0].n = N;
  int rc = dl_select(p, w, ws, st);
  if (rc != SSDSB_OK) return rc;
  dl_merge<1><<<dim3(p.n_chunks, 1, B), DL_SORT_NT, 0, st>>>(p, nullptr, nullptr, nullptr, nullptr, d_keys, d_count);
  SSDSB_LAUNCH_CHECK("dl_merge<1>");
  return SSDSB_OK;
}

}  // namespace ssdsb
