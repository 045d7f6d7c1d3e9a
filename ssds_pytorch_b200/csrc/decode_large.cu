// Exact top-K of long score rows with FIVE cheap streaming passes and a parallel sort — used by
//   * decode for LARGE top_n (1024 < top_n <= 65536 per level; the 20 000-per-level NMS stress of SURVEY 8d cfg 5),
//     instead of one pass over the score maps per 1024 results, and
//   * nms for long candidate rows (N > 8192): the top-2048 keys of each image in sorted order, instead of one CTA
//     streaming and pruning all N scores.
//
// reference: ssds/modeling/layers/box.py:408-477 (keep = conf >= thr; topk(min(top_n, |keep|)) sorted descending;
// index -> (a, c, y, x); delta2box; rescore) and :496-505 (drop score <= 0, sort descending) — same outputs, same
// order (descending score, ascending flat index among equal scores) as decode.cu / nms.cu.
//
// Keys are ordered-float(score) (32 bits) + flat index (unique).  A 3-level MSD radix select over ALL 32 score bits
// finds the K-th score exactly, fused with emission; the tie group at the K-th score is cut by flat index:
//   pass 1  histogram of score bits [31:21] of every score >= thr                       -> b1
//   pass 2  keys above bin b1 -> result list; histogram of bits [20:10] inside b1       -> b2
//   pass 3  keys of b1 above b2 -> result list; histogram of bits [9:0] inside (b1,b2)  -> b3: the K-th score itself
//   pass 4  keys of (b1,b2) above b3 -> result list; per-slice count of keys EQUAL to the K-th score
//   pass 5  of those equal keys, the (K - #above) with the lowest flat indices -> result list (slices are in index
//           order, so a slice only needs the counts of the slices before it; most CTAs exit at once)
//   dl_sort_chunks  every 2048-key chunk of the result list is sorted by its own CTA (bitonic, shared memory)
//   dl_merge        rank of a key = its rank in its chunk + sum over the other chunks of the keys above it
//                   (binary search, L2): the key is result number `rank` — decode emits the detection row directly
//                   (box.py:443-471), nms writes the sorted key list.
// No distribution-dependent slow path: a random-init head puts ALL scores within 1 % of each other (and a blank
// image makes most of them exactly equal) — the radix levels resolve down to the last bit and the index cut is a
// parallel prefix over slices.  The scan after each histogram is done by the LAST CTA of the row (threadfence +
// ticket); result keys are staged in shared memory and flushed with one global reservation per CTA; the first
// histogram (a few hot bins: neighbouring scores share exponent and top mantissa bits) goes through a per-thread
// run-length cache.  memset + 7 launches, no host sync, CUDA-graph capturable.
#include "decode_emit.cuh"
#include "decode_large.h"

namespace ssdsb {
namespace {

constexpr int DL_NT = 256;
constexpr int DL_SLICE_MAX = 16 * 1024;  // scores per CTA in the streaming passes (short rows use smaller slices)
constexpr int DL_BINS = 2048;
constexpr int DL_CHUNK = 2048;           // keys sorted per CTA
constexpr int DL_SORT_NT = 1024;
constexpr int DL_STAGE = 3072;           // keys staged per CTA before a flush (>= 2048 + the flush threshold)
constexpr int DL_GROUP = DL_NT * 4;      // elements one CTA handles per 16-byte load round

struct DlSeg {      // per row (image, level), in the workspace
  int b1, above1;   // bin of the K-th key among score bits [31:21]; keys in higher bins   (b1 = -1: < K keys pass)
  int b2, above2;   // same inside bin b1 for bits [20:10]
  int b3, above3;   // same inside (b1, b2) for bits [9:0]: (b1, b2, b3) is the K-th key's score, exactly
  int n_sure;
  int ticket1, ticket2, ticket3;
  int pad[6];
};

struct DlRow {
  const float* scores;      // image 0's row of this "level"; image b's row starts at scores + b * n
  int n;
};

struct DlParams {
  DlRow row[SSDSB_MAX_LEVELS];
  ssdsb_level lv[SSDSB_MAX_LEVELS];        // decode mode only (emission)
  int slice_begin[SSDSB_MAX_LEVELS + 1];
  int n_levels, B, K, n_chunks;
  int slice;                  // scores per CTA: a multiple of 4 * DL_GROUP
  float threshold;
  int rescore;
  DlSeg* seg;                 // [B*L]
  int* hist;                  // [3][B*L][DL_BINS]
  int* slice_eq;              // [B][total slices]: keys equal to the K-th score per slice (pass 4 -> pass 5)
  unsigned long long* cand;   // [B*L][K]: the selected keys, unordered
  unsigned long long* sorted; // [B*L][n_chunks * DL_CHUNK]
};

// PASS 1 histogram (score bits [31:21]): sigmoid scores cluster in a handful of bins (same exponent, same top
// mantissa bits), so per-element shared-memory atomics would serialise on a few addresses.  Each thread keeps a
// one-entry run-length cache (bin, count) across its whole slice and only touches shared memory when the bin changes.
struct BinRun {
  int bin, cnt;
};
__device__ __forceinline__ void run_add(int* s_hist, BinRun& r, int bin, bool valid) {
  if (!valid) return;
  if (bin == r.bin) {
    ++r.cnt;
  } else {
    if (r.cnt) atomicAdd(&s_hist[r.bin], r.cnt);
    r.bin = bin;
    r.cnt = 1;
  }
}

// the last CTA of a row: find the bin d (from the top) where the running count reaches `want`;
// (d, keys in bins above d) -> s_out[0], s_out[1]; d = -1 when the row holds fewer than `want` keys.
// The whole histogram is first pulled into shared memory by all threads (ONE L2 latency; a serial walk over global
// memory cost ~1 us per 32 bins and dominated the histogram passes).  Called by every thread of the CTA.
__device__ __forceinline__ void scan_bins(const int* __restrict__ ghist, int want, int* s_bins, int* s_out) {
  for (int k = threadIdx.x; k < DL_BINS; k += DL_NT) s_bins[k] = __ldcg(ghist + k);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int rem = want, d = 0, above = 0;
    bool found = false;
    for (int base = DL_BINS - 32; base >= 0 && !found; base -= 32) {
      const int cnt = s_bins[base + 31 - lane];                // lane 0 = highest bin of the chunk
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

template <int PASS>
__global__ void __launch_bounds__(DL_NT)
dl_pass(const __grid_constant__ DlParams p) {
  __shared__ int s_hist[PASS <= 3 ? DL_BINS : 1];
  __shared__ unsigned long long s_keys[PASS >= 2 ? DL_STAGE : 1];
  __shared__ int s_out[2];
  __shared__ int s_flag, s_nk, s_base, s_eq, s_run;
  __shared__ int s_warp[DL_NT / 32];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.slice_begin[l + 1]) ++l;
  const int slice = blockIdx.x - p.slice_begin[l];
  const int L = p.n_levels;
  const int sg = b * L + l;
  DlSeg* seg = p.seg + sg;
  const int n = p.row[l].n;
  const int begin = slice * p.slice;
  const int end = min(n, begin + p.slice);
  const float* src = p.row[l].scores + (size_t)b * n;
  const float thr = p.threshold;
  const int total_slices = p.slice_begin[L];
  int* my_eq = p.slice_eq + (size_t)b * total_slices + blockIdx.x;
  int b1 = 0;
  uint32_t pre22 = 0u, vstar = 0u;     // (b1, b2) as the top 22 ordered-score bits; the full ordered K-th score
  int r_eq = 0, eq_before = 0;
  if (PASS >= 2) {
    b1 = __ldcg(&seg->b1);
    if (PASS >= 3 && b1 < 0) return;   // fewer than K keys pass: pass 2 took them all
    if (PASS >= 3) pre22 = ((uint32_t)b1 << 11) | (uint32_t)__ldcg(&seg->b2);
    if (PASS >= 4) vstar = (pre22 << 10) | (uint32_t)__ldcg(&seg->b3);
    if (PASS == 5) {
      r_eq = p.K - __ldcg(&seg->above1) - __ldcg(&seg->above2) - __ldcg(&seg->above3);
      // keys equal to the K-th score in the earlier slices of this row (index order = slice order)
      int part = 0;
      for (int s2 = tid; s2 < slice; s2 += DL_NT)
        part += __ldcg(p.slice_eq + (size_t)b * total_slices + p.slice_begin[l] + s2);
      for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
      if (lane == 0) s_warp[wid] = part;
      __syncthreads();
      for (int w = 0; w < DL_NT / 32; ++w) eq_before += s_warp[w];
      if (eq_before >= r_eq || __ldcg(my_eq) == 0) return;      // nothing to take from this slice (block-uniform)
      __syncthreads();
    }
  }
  if (PASS <= 3) {
    for (int k = tid; k < DL_BINS; k += DL_NT) s_hist[k] = 0;
  }
  if (tid == 0) {
    s_nk = 0;
    s_eq = 0;
    s_run = eq_before;
  }
  __syncthreads();
  unsigned long long* sure = p.cand + (size_t)sg * p.K;
  const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.row[l].scores) & 15) == 0);

  // result keys are staged in shared memory; one global reservation per flush
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
  BinRun run = {0, 0};
  // one group = DL_GROUP consecutive elements, 4 per lane in index order (lane-major): v[e] at index i0 + e
  auto group = [&](const float (&v)[4], int i0, bool inb) {
    bool ok[4];
    uint32_t ord[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ok[e] = inb && (v[e] >= thr);
      ord[e] = float_to_ordered(v[e]);
    }
    if (PASS == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) run_add(s_hist, run, (int)(ord[e] >> 21), ok[e]);
    } else if (PASS == 2) {
      if (b1 < 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) stage(ok[e], make_key(v[e], (uint32_t)(i0 + e)));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int bin1 = (int)(ord[e] >> 21);
          // the sub-bins of b1 are spread (next 11 score bits): plain shared-memory atomics, few collisions
          if (ok[e] && bin1 == b1) atomicAdd(&s_hist[(ord[e] >> 10) & 2047u], 1);
          stage(ok[e] && bin1 > b1, make_key(v[e], (uint32_t)(i0 + e)));
        }
      }
    } else if (PASS == 3) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t top22 = ord[e] >> 10;
        if (ok[e] && top22 == pre22) atomicAdd(&s_hist[ord[e] & 1023u], 1);
        stage(ok[e] && (top22 >> 11) == (uint32_t)b1 && top22 > pre22, make_key(v[e], (uint32_t)(i0 + e)));
      }
    } else if (PASS == 4) {
      int neq = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        stage(ok[e] && (ord[e] >> 10) == pre22 && ord[e] > vstar, make_key(v[e], (uint32_t)(i0 + e)));
        neq += (ok[e] && ord[e] == vstar) ? 1 : 0;
      }
      neq = __reduce_add_sync(0xffffffffu, neq);
      if (lane == 0 && neq) atomicAdd(&s_eq, neq);
    } else {
      // ordered ranks of the keys equal to the K-th score: lane-major inside the group == flat index order
      int c = 0;
      bool eq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        eq[e] = ok[e] && ord[e] == vstar;
        c += eq[e] ? 1 : 0;
      }
      int inc = c;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_warp[wid] = inc;
      __syncthreads();
      int rank = s_run + inc - c;
      for (int w = 0; w < wid; ++w) rank += s_warp[w];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool take = eq[e] && rank < r_eq;
        stage(take, make_key(v[e], (uint32_t)(i0 + e)));
        rank += eq[e] ? 1 : 0;
      }
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < DL_NT / 32; ++w) t += s_warp[w];
        s_run += t;
      }
      __syncthreads();
    }
  };

  // block-uniform trip counts (the handlers use full-mask collectives / barriers); 4 x 16-byte loads in flight
  if (vec_ok) {
    const int iters = (end - begin + 4 * DL_GROUP - 1) / (4 * DL_GROUP);
    for (int it = 0; it < iters; ++it) {
      float4 q[4];
      int i0[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        i0[u] = begin + (it * 4 + u) * DL_GROUP + tid * 4;
        in[u] = i0[u] + 4 <= end;
        q[u] = in[u] ? __ldcs(reinterpret_cast<const float4*>(src + i0[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
        group(v, i0[u], in[u]);
        if (PASS >= 2 && (u & 1)) {                           // <= 2048 keys staged since the last check
          __syncthreads();
          if (s_nk > DL_STAGE - 2048) flush();                // block-uniform
        }
      }
    }
  } else {
    // unaligned / odd-length rows: the same groups with scalar loads (element i0 + e of lane tid)
    const int iters = (end - begin + DL_GROUP - 1) / DL_GROUP;
    for (int it = 0; it < iters; ++it) {
      const int i0 = begin + it * DL_GROUP + tid * 4;
      float v[4];
      bool any_in = false;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool in_e = i0 + e < end;
        v[e] = in_e ? __ldcs(src + i0 + e) : -INFINITY;     // -inf < thr for every finite thr: never a key
        any_in = any_in || in_e;
      }
      group(v, i0, any_in);
      if (PASS >= 2 && (it & 1)) {
        __syncthreads();
        if (s_nk > DL_STAGE - 2048) flush();
      }
    }
  }
  if (PASS == 1 && run.cnt) atomicAdd(&s_hist[run.bin], run.cnt);
  __syncthreads();
  if (PASS >= 2) flush();
  if (PASS == 4 && tid == 0) *my_eq = s_eq;
  if (PASS >= 4) return;

  // flush the CTA histogram, then the last CTA of the row scans it
  int* ghist = p.hist + ((size_t)(PASS - 1) * p.B * L + sg) * DL_BINS;
  for (int k = tid; k < DL_BINS; k += DL_NT) {
    const int c = s_hist[k];
    if (c) atomicAdd(ghist + k, c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int slices = p.slice_begin[l + 1] - p.slice_begin[l];
    int* ticket = PASS == 1 ? &seg->ticket1 : (PASS == 2 ? &seg->ticket2 : &seg->ticket3);
    s_flag = (atomicAdd(ticket, 1) == slices - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (PASS == 1) {
    scan_bins(ghist, p.K, s_hist, s_out);
    if (tid == 0) {
      seg->b1 = s_out[0];
      seg->above1 = s_out[0] >= 0 ? s_out[1] : 0;
    }
  } else if (PASS == 2) {
    if (b1 >= 0) {                                             // (b1 < 0: no histogram was built)
      scan_bins(ghist, p.K - __ldcg(&seg->above1), s_hist, s_out);
      if (tid == 0) {
        seg->b2 = s_out[0];
        seg->above2 = s_out[1];
      }
    }
  } else {
    scan_bins(ghist, p.K - __ldcg(&seg->above1) - __ldcg(&seg->above2), s_hist, s_out);
    if (tid == 0) {
      seg->b3 = s_out[0];
      seg->above3 = s_out[1];
    }
  }
}

// sorts chunk blockIdx.x of a row's result list (descending) into `sorted`
__global__ void __launch_bounds__(DL_SORT_NT, 1)
dl_sort_chunks(const __grid_constant__ DlParams p) {
  __shared__ unsigned long long s[DL_CHUNK];
  const int c = blockIdx.x, l = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int sg = b * p.n_levels + l;
  const DlSeg* seg = p.seg + sg;
  const int K = p.K;
  const int n = min(__ldcg(&seg->n_sure), K);
  const int lo = c * DL_CHUNK;
  if (lo >= n) return;
  const unsigned long long* cand = p.cand + (size_t)sg * K;
  for (int i = tid; i < DL_CHUNK; i += DL_SORT_NT) {
    const int q = lo + i;
    s[i] = (q < n) ? __ldcg(cand + q) : 0ull;   // 0 < every real key
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
  const int n = min(__ldcg(&seg->n_sure), K);
  const int nch = (n + DL_CHUNK - 1) / DL_CHUNK;
  const unsigned long long* sorted = p.sorted + (size_t)sg * p.n_chunks * DL_CHUNK;
  const size_t row = (size_t)sg * K;
  if (MODE == 1 && c == 0 && tid == 0) out_count[sg] = n;
  // zero padding of the result slots nobody owns
  for (int t = c * DL_CHUNK + tid; t < min(K, (c + 1) * DL_CHUNK); t += DL_SORT_NT) {
    if (t >= n) {
      if (MODE == 0) emit_detection(p.lv[l], b, false, 0ull, p.rescore, row + t, out_scores, out_boxes, out_classes,
                                    out_index);
      else out_keys[row + t] = 0ull;
    }
  }
  if (c >= nch) return;                      // block-uniform
  // every other chunk is staged through shared memory once and binary-searched there by all threads
  __shared__ unsigned long long s_other[DL_CHUNK];
  const int mine = min(DL_CHUNK, n - c * DL_CHUNK);
  unsigned long long x[DL_CHUNK / DL_SORT_NT];
  int rank[DL_CHUNK / DL_SORT_NT];
#pragma unroll
  for (int u = 0; u < DL_CHUNK / DL_SORT_NT; ++u) {
    const int i = tid + u * DL_SORT_NT;
    x[u] = (i < mine) ? __ldcg(sorted + (size_t)c * DL_CHUNK + i) : 0ull;
    rank[u] = i;
  }
  for (int o = 0; o < nch; ++o) {
    if (o == c) continue;
    const int m = min(DL_CHUNK, n - o * DL_CHUNK);
    __syncthreads();
    for (int i = tid; i < m; i += DL_SORT_NT) s_other[i] = __ldcg(sorted + (size_t)o * DL_CHUNK + i);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < DL_CHUNK / DL_SORT_NT; ++u) {
      int lo = 0, hi = m;                    // keys of s_other[0..m) (descending) greater than x[u]
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_other[mid] > x[u]) lo = mid + 1;
        else hi = mid;
      }
      rank[u] += lo;
    }
  }
#pragma unroll
  for (int u = 0; u < DL_CHUNK / DL_SORT_NT; ++u) {
    if (tid + u * DL_SORT_NT < mine) {
      if (MODE == 0) emit_detection(p.lv[l], b, true, x[u], p.rescore, row + rank[u], out_scores, out_boxes,
                                    out_classes, out_index);
      else out_keys[row + rank[u]] = x[u];
    }
  }
}

struct DlLayout {
  size_t seg, hist, zero_bytes, slice_eq, cand, sorted, total;
  int n_chunks;
};

DlLayout dl_layout(int B, int L, int K, long long total_slices) {
  DlLayout w;
  const size_t segs = (size_t)B * L;
  w.n_chunks = (K + DL_CHUNK - 1) / DL_CHUNK;
  size_t o = 0;
  w.seg = o; o += align_up(segs * sizeof(DlSeg), 256);
  w.hist = o; o += align_up(3 * segs * DL_BINS * 4, 256);
  w.zero_bytes = o;               // everything up to here is zeroed per call
  w.slice_eq = o; o += align_up((size_t)B * (size_t)total_slices * 4, 256);
  w.cand = o; o += align_up(segs * (size_t)K * 8, 256);
  w.sorted = o; o += align_up(segs * (size_t)w.n_chunks * DL_CHUNK * 8, 256);
  w.total = o;
  return w;
}

// scores per CTA: as large as DL_SLICE_MAX for long rows, smaller when that would leave most SMs idle
int dl_slice_size(const int* n_per_level, int L, int B) {
  long long tot = 0;
  for (int l = 0; l < L; ++l) tot += n_per_level[l];
  int slice = DL_SLICE_MAX;
  while (slice > 4 * DL_GROUP && tot * B / slice < 2 * 148) slice >>= 1;
  return slice;
}

long long dl_total_slices(const int* n_per_level, int L, int slice) {
  long long t = 0;
  for (int l = 0; l < L; ++l) {
    long long sl = ((long long)n_per_level[l] + slice - 1) / slice;
    t += sl < 1 ? 1 : sl;
  }
  return t;
}

int dl_select(DlParams& p, const DlLayout& w, unsigned char* ws, cudaStream_t st) {
  p.seg = reinterpret_cast<DlSeg*>(ws + w.seg);
  p.hist = reinterpret_cast<int*>(ws + w.hist);
  p.slice_eq = reinterpret_cast<int*>(ws + w.slice_eq);
  p.cand = reinterpret_cast<unsigned long long*>(ws + w.cand);
  p.sorted = reinterpret_cast<unsigned long long*>(ws + w.sorted);
  p.n_chunks = w.n_chunks;
  p.slice_begin[0] = 0;
  for (int l = 0; l < p.n_levels; ++l) {
    int slices = (p.row[l].n + p.slice - 1) / p.slice;
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
  dl_pass<4><<<g, DL_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_pass<4>");
  dl_pass<5><<<g, DL_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_pass<5>");
  dl_sort_chunks<<<dim3(p.n_chunks, p.n_levels, p.B), DL_SORT_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("dl_sort_chunks");
  return SSDSB_OK;
}

DlLayout layout_for_levels(const ssdsb_level* levels, int n_levels, int B, int top_n, int* slice) {
  int nn[SSDSB_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) nn[l] = levels[l].A * levels[l].C * levels[l].H * levels[l].W;
  *slice = dl_slice_size(nn, n_levels, B);
  return dl_layout(B, n_levels, top_n, dl_total_slices(nn, n_levels, *slice));
}

}  // namespace

int decode_large_max_k() { return 64 * 1024; }

size_t decode_large_workspace_bytes(const ssdsb_level* levels, int n_levels, int B, int top_n) {
  int slice;
  return layout_for_levels(levels, n_levels, B, top_n, &slice).total + 256;
}

int decode_large(const ssdsb_level* levels, int n_levels, int B, float threshold, int top_n, int rescore,
                 float* d_scores, float* d_boxes, float* d_classes, int32_t* d_index, void* d_workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  int slice;
  const DlLayout w = layout_for_levels(levels, n_levels, B, top_n, &slice);
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  if (!d_workspace || workspace_bytes < w.total + 256)
    return fail(SSDSB_ERR_WORKSPACE, "decode: workspace %zu B given, %zu B needed", workspace_bytes, w.total + 256);
  DlParams p;
  p.slice = slice;
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

size_t topk_rows_workspace_bytes(int B, int N, int K) {
  const int slice = dl_slice_size(&N, 1, B);
  return dl_layout(B, 1, K, dl_total_slices(&N, 1, slice)).total + 256;
}

int topk_rows(const float* d_scores, int B, int N, float min_score, int K, unsigned long long* d_keys, int* d_count,
              void* d_workspace, size_t workspace_bytes, cudaStream_t st) {
  const int slice = dl_slice_size(&N, 1, B);
  const DlLayout w = dl_layout(B, 1, K, dl_total_slices(&N, 1, slice));
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
  if (!d_workspace || workspace_bytes < w.total + 256)
    return fail(SSDSB_ERR_WORKSPACE, "topk_rows: workspace %zu B given, %zu B needed", workspace_bytes, w.total + 256);
  DlParams p;
  p.slice = slice;
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
