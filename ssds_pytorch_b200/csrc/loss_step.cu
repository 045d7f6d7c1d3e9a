// The loss half of one training step in ONE launch: anchor matching + (MultiBoxLoss hard-negative mining |
// FocalLoss) + (SmoothL1 | IoU-family) localisation loss + the caller's masking / per-level normalisation,
// over ALL levels and ALL images of the batch.
//
// reference: ssds/pipeline/pipeline_anchor_basic.py:62-97 (same body pipeline_anchor_apex.py:37-72)
//     for each level:  extract_targets (box.py:362-405 -> snap_to_anchors_by_iou :116-226)
//                      cls_criterion (criterion.py:43-71 MultiBoxLoss | :95-108 FocalLoss) * (depth >= 0), sum
//                      loc_criterion (criterion.py:138-151 | :175-239) * (depth > 0), sum
//                      fg_targets += max(#(depth > 0), 1)
//     cls_loss = sum / fg_targets, loc_loss = sum / fg_targets
// Hard negatives are mined per (image, level) — MultiBoxLoss is called per level on that level's anchors and
// ranks per image (criterion.py:57-68; per-image is the intended meaning of the B==1-only expand_as).
//
// B200 design (one pass over the logits, HBM-bound: 4*C bytes per anchor read once):
//   * grid = every (level, image, 512-anchor tile); a thread owns one anchor (x fastest => each class-plane
//     access of a warp is one coalesced 128-byte line); the image's targets are staged once per CTA in shared
//     memory; matching uses match_core.cuh (bit-identical depth to ssdsb_match_iou); box targets exist only in
//     registers, for positives; no one-hot target, no depth / box_target tensors unless the caller asks.
//   * BCE: softplus(x) = max(x,0) + log1p(exp(-|x|)) with ex2.approx and a degree-8 polynomial for log1p on
//     (0,1] (max rel. error 2.3e-7) — the precise expf/log1pf pair costs ~3x the instructions and would make
//     the pass issue-bound instead of HBM-bound.  8 independent plane loads per thread in flight.
//   * per anchor: (max CE as ordered u32, sum CE) -> workspace (8 B/anchor, L2-resident); positives, focal
//     terms and loc losses are block-reduced in a fixed order into per-CTA partials.
//   * the LAST CTA of an (image, level) pair (threadfence + atomic ticket) does that pair's 3-pass 11-bit
//     radix select of the num_neg-th largest max-CE, the index cut among equals (== stable descending sort),
//     and the deterministic double-precision sums; the last pair to finish reduces the final scalars.
//     Big levels are scheduled first, so their selects overlap the streaming pass of the smaller levels.
// No host sync, no data-dependent launch shape: CUDA-graph capturable.  Deterministic (fixed reduction orders).
#include "loss_math.cuh"
#include "match_core.cuh"

namespace ssdsb {
namespace {

constexpr int LS_NT = 512;
constexpr int LS_WARPS = LS_NT / 32;

struct LsLevel {
  const float* conf;
  const float* loc;
  const float4* anchors;
  float* depth;
  float* box_target;
  int A, C, H, W, stride;
  int HW, N;        // H*W, A*H*W
  int tiles;        // ceil(N / LS_NT)
  int cta0;         // first CTA of this level (level-major, then image, then tile)
  long long off;    // offset of this level's [B][N] block in the mce / sce workspace arrays
};

struct LsParams {
  LsLevel lv[SSDSB_MAX_LEVELS];
  int L, B, T;
  const float* targets;
  float match_thr, unmatch_thr;
  int cls_kind, negpos_ratio;
  float alpha, gamma;
  int loc_kind;
  float beta;
  // workspace
  uint32_t* mce;
  float* sce;
  double* part_cls;   // [ctas] positives' CE sum (MultiBox) / focal sum over depth >= 0 (Focal)
  double* part_loc;   // [ctas]
  int* npos;          // [L*B]
  int* done;          // [L*B] CTAs finished per pair
  int* pairs_done;    // [1]
  double* pair_cls;   // [L*B]
  double* pair_loc;   // [L*B]
  // outputs
  float* out_scalars;   // [3] cls_loss, loc_loss, fg_targets
  float* out_cls_sum;   // [L*B] or NULL
  float* out_loc_sum;   // [L*B] or NULL
  float* out_num_pos;   // [L*B] or NULL
};

__device__ __forceinline__ float softplus_fast(float x) {
  // max(x,0) + log1p(exp(-|x|));  log1p(e) = e*q(e), q = degree-8 minimax-like fit on [0,1]
  const float e = __expf(-fabsf(x));
  float q = 0.005126102361828089f;
  q = fmaf(q, e, -0.029074065387248993f);
  q = fmaf(q, e, 0.0775160863995552f);
  q = fmaf(q, e, -0.13602247834205627f);
  q = fmaf(q, e, 0.19076880812644958f);
  q = fmaf(q, e, -0.2483539879322052f);
  q = fmaf(q, e, 0.3331812024116516f);
  q = fmaf(q, e, -0.4999944567680359f);
  q = fmaf(q, e, 0.9999999403953552f);
  return fmaf(q, e, fmaxf(x, 0.0f));
}

// block reduction of one double per thread in a fixed order; result valid on thread 0
__device__ __forceinline__ double block_sum(double v, double* s_w) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < LS_WARPS; ++w) t += s_w[w];
  return t;
}

__global__ void __launch_bounds__(LS_NT, 2)
loss_step_kernel(const __grid_constant__ LsParams p) {
  __shared__ Tgt s_t[MATCH_TCHUNK];
  __shared__ int s_n;
  __shared__ double s_w[LS_WARPS];
  __shared__ int s_flag;
  __shared__ int s_hist[2048];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_running, s_cut, s_tie;
  __shared__ int s_warp[LS_WARPS];

  // ---- which (level, image, tile) ----
  int l = 0;
#pragma unroll 1
  for (int k = 1; k < p.L; ++k)
    if ((int)blockIdx.x >= p.lv[k].cta0) l = k;
  const LsLevel& lv = p.lv[l];
  const int rel = (int)blockIdx.x - lv.cta0;
  const int b = rel / lv.tiles;
  const int tile = rel - b * lv.tiles;
  const int pair = l * p.B + b;
  const int HW = lv.HW, N = lv.N, A = lv.A, C = lv.C;
  const int tid = threadIdx.x;
  const int i = tile * LS_NT + tid;                 // (a, y, x), x fastest
  const bool active = i < N;
  const int a = active ? i / HW : 0;
  const int yx = active ? i - a * HW : 0;
  const int y = yx / lv.W, x = yx - y * lv.W;

  // ---- match (box.py:116-226) ----
  const float4 ba = __ldg(lv.anchors + a);
  const float fx = (float)(x * lv.stride), fy = (float)(y * lv.stride);
  const float ax1 = fx + ba.x, ay1 = fy + ba.y, ax2 = fx + ba.z, ay2 = fy + ba.w;
  const float aarea = (ax2 - ax1 + 1.0f) * (ay2 - ay1 + 1.0f);
  MatchState m = match_init();
  const float* tg = p.targets + (size_t)b * p.T * 5;
  for (int t0 = 0; t0 < p.T; t0 += MATCH_TCHUNK) {
    __syncthreads();
    if (tid < 32) stage_targets_warp(tg, p.T, t0, 0.0f, s_t, &s_n, tid);
    __syncthreads();
    match_fold(m, s_t, s_n, ax1, ay1, ax2, ay2, aarea, 0.0f, 0.0f, 0.0f);
  }
  int ci = -1;
  const float depth = active ? match_depth(m, p.match_thr, p.unmatch_thr, 0.0f, &ci) : -1.0f;
  const bool is_pos = active && depth > 0.0f;
  const size_t ba_off = (size_t)b * A + a;
  if (active && lv.depth) lv.depth[ba_off * HW + yx] = depth;

  // ---- localisation loss on positives (criterion.py:138-151 | :175-239), box target in registers ----
  double loc_v = 0.0;
  if (p.loc_kind >= 0 || lv.box_target) {
    float dt[4] = {0.f, 0.f, 0.f, 0.f};
    if (active && m.any && (is_pos || lv.box_target)) match_delta(m, ax1, ay1, ax2, ay2, dt);
    if (active && lv.box_target) {
      float* bt = lv.box_target + ba_off * 4 * HW + yx;
      bt[0] = dt[0]; bt[(size_t)HW] = dt[1]; bt[(size_t)2 * HW] = dt[2]; bt[(size_t)3 * HW] = dt[3];
    }
    if (is_pos && p.loc_kind >= 0) {
      const float* lp = lv.loc + ba_off * 4 * HW + yx;
      float pr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pr[k] = __ldcs(lp + (size_t)k * HW);
      if (p.loc_kind == LOC_SMOOTHL1) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += smooth_l1(pr[k], dt[k], p.beta);
        loc_v = (double)s;
      } else {
        loc_v = (double)iou_family(pr, dt, p.loc_kind);
      }
    }
  }

  // ---- classification pass: C logits of this anchor, 8 plane loads in flight ----
  double cls_v = 0.0;      // what goes to the per-CTA partial
  if (active && depth >= 0.0f) {
    const float* lg = lv.conf + ba_off * C * HW + yx;
    const int cpos = is_pos ? (int)depth - 1 : -1;
    if (p.cls_kind == 0) {
      float mx = 0.0f, sum = 0.0f;      // softplus > 0, so 0 is a valid identity for the max
      // every class is reduced as a negative, BCE(x, 0) = softplus(x); the one positive class of a positive anchor is
      // corrected after the loop (BCE(x, 1) = softplus(x) - x), and a positive's max is never used (criterion.py:59)
      auto take = [&](float x, int) {
        const float ce = softplus_fast(x);
        mx = fmaxf(mx, ce);
        sum += ce;
      };
      // software pipeline: the 8 plane loads of the NEXT group are issued before the current group is reduced, so
      // 16 independent loads per thread are in flight (the pass is bound by load latency x occupancy otherwise)
      int c = 0;
      float v[8], nx[8];
      const int groups = C >> 3;
      if (groups > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(lg + (size_t)k * HW);
      }
      for (int gq = 0; gq < groups; ++gq, c += 8) {
        if (gq + 1 < groups) {
#pragma unroll
          for (int k = 0; k < 8; ++k) nx[k] = __ldcs(lg + (size_t)(c + 8 + k) * HW);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) take(v[k], c + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = nx[k];
      }
      for (; c < C; ++c) take(__ldcs(lg + (size_t)c * HW), c);
      if (cpos >= 0 && cpos < C) sum -= __ldg(lg + (size_t)cpos * HW);   // (a label >= C matches no class plane)
      if (is_pos) {                                     // always kept; max_ce := 0 (criterion.py:59)
        cls_v = (double)sum;
        p.mce[lv.off + (size_t)b * N + i] = float_to_ordered(0.0f);
        p.sce[lv.off + (size_t)b * N + i] = 0.0f;
      } else {                                          // hard-negative candidate
        p.mce[lv.off + (size_t)b * N + i] = float_to_ordered(mx);
        p.sce[lv.off + (size_t)b * N + i] = sum;
      }
    } else if (p.gamma == 2.0f) {
      // FocalLoss (criterion.py:95-108) with the default gamma = 2, same fast transcendental path as above:
      // e = exp(-|x|), softplus from the polynomial, sigmoid from one reciprocal.  Every class is reduced as a
      // negative, (1-alpha) * p^2 * softplus(x); the positive class of a positive anchor is corrected afterwards.
      const float an = 1.0f - p.alpha;
      float sum = 0.0f;
      auto neg_term = [&](float x) {
        const float e = __expf(-fabsf(x));
        const float inv = __frcp_rn(1.0f + e);
        const float pr = (x >= 0.0f) ? inv : e * inv;               // sigmoid(x)
        return an * (pr * pr) * softplus_fast(x);
      };
      int c = 0;
      float v[8], nx[8];
      const int groups = C >> 3;
      if (groups > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(lg + (size_t)k * HW);
      }
      for (int gq = 0; gq < groups; ++gq, c += 8) {
        if (gq + 1 < groups) {
#pragma unroll
          for (int k = 0; k < 8; ++k) nx[k] = __ldcs(lg + (size_t)(c + 8 + k) * HW);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += neg_term(v[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = nx[k];
      }
      for (; c < C; ++c) sum += neg_term(__ldcs(lg + (size_t)c * HW));
      if (cpos >= 0 && cpos < C) {
        const float x = __ldg(lg + (size_t)cpos * HW);
        const float e = __expf(-fabsf(x));
        const float inv = __frcp_rn(1.0f + e);
        const float q = (x >= 0.0f) ? e * inv : inv;                // 1 - sigmoid(x)
        sum += p.alpha * (q * q) * (softplus_fast(x) - x) - neg_term(x);
      }
      cls_v = (double)sum;
    } else {
      float sum = 0.0f;
      for (int c = 0; c < C; ++c)
        sum += focal_term(__ldcs(lg + (size_t)c * HW), (c == cpos) ? 1.0f : 0.0f, p.alpha, p.gamma);
      cls_v = (double)sum;
    }
  } else if (active && p.cls_kind == 0) {               // ignored anchor: rank value 0, contributes nothing
    p.mce[lv.off + (size_t)b * N + i] = float_to_ordered(0.0f);
    p.sce[lv.off + (size_t)b * N + i] = 0.0f;
  }

  // ---- per-CTA partials (fixed order) + positives count ----
  const double cs = block_sum(cls_v, s_w);
  const double lsum = block_sum(loc_v, s_w);
  const int npos_cta = __syncthreads_count(is_pos ? 1 : 0);
  if (tid == 0) {
    p.part_cls[blockIdx.x] = cs;
    p.part_loc[blockIdx.x] = lsum;
    if (npos_cta) atomicAdd(p.npos + pair, npos_cta);
    __threadfence();
    s_flag = (atomicAdd(p.done + pair, 1) == lv.tiles - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;

  // =============== last CTA of this (image, level) pair ===============
  __threadfence();
  const int npos = __ldcg(p.npos + pair);
  double neg_sum = 0.0;
  if (p.cls_kind == 0) {
    long long want = (long long)p.negpos_ratio * npos;
    if (want > N - 1) want = N - 1;                     // criterion.py:65
    if (want > 0) {
      const uint32_t* u = p.mce + lv.off + (size_t)b * N;
      if (tid == 0) {
        s_prefix = 0u;
        s_remaining = (int)want;
      }
      // 3-pass radix select (11 + 11 + 10 bits) of the want-th largest key
      const int shifts[3] = {21, 10, 0};
      const int bits[3] = {11, 11, 10};
      uint32_t mask = 0u;
      for (int pass = 0; pass < 3; ++pass) {
        const int nb = 1 << bits[pass];
        for (int k = tid; k < nb; k += LS_NT) s_hist[k] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const int shift = shifts[pass];
        // 8 independent L2 loads per thread in flight (this tail is latency-bound: one CTA, N up to 76 800 keys).
        // max-CE values of one image cluster in a handful of bins (same exponent, same top mantissa bits), which
        // would serialise per-element shared-memory atomics: each thread keeps a one-entry (bin, count) run cache.
        int run_bin = 0, run_cnt = 0;
        for (int base = 0; base < N; base += LS_NT * 8) {
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = base + k * LS_NT + tid;
            v[k] = (j < N) ? __ldcg(u + j) : 0u;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = base + k * LS_NT + tid;
            if (j < N && (v[k] & mask) == prefix) {
              const int bin = (int)((v[k] >> shift) & (uint32_t)(nb - 1));
              if (bin == run_bin) {
                ++run_cnt;
              } else {
                if (run_cnt) atomicAdd(&s_hist[run_bin], run_cnt);
                run_bin = bin;
                run_cnt = 1;
              }
            }
          }
        }
        if (run_cnt) atomicAdd(&s_hist[run_bin], run_cnt);
        __syncthreads();
        if (tid < 32) {
          // warp 0: find the bin d (from the top) where the running count reaches `rem`
          int rem = s_remaining;
          int d = nb - 1;
          bool found = false;
          for (int base = nb - 32; base >= 0 && !found; base -= 32) {
            const int cnt = s_hist[base + 31 - tid];          // lane 0 = highest bin of the chunk
            int inc = cnt;
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, inc, o);
              if (tid >= o) inc += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, inc >= rem);
            if (hit) {
              const int lane = __ffs(hit) - 1;
              const int before = __shfl_sync(0xffffffffu, inc, lane) - __shfl_sync(0xffffffffu, cnt, lane);
              d = base + 31 - lane;
              rem -= before;
              found = true;
            } else {
              rem -= __shfl_sync(0xffffffffu, inc, 31);
            }
          }
          if (!found) d = 0;
          if (tid == 0) {
            s_remaining = rem;
            s_prefix = prefix | ((uint32_t)d << shift);
            s_tie = s_hist[d];           // after the last pass: how many keys carry exactly the cut value
          }
        }
        mask |= (uint32_t)(nb - 1) << shift;
        __syncthreads();
      }
      const uint32_t vstar = s_prefix;
      const int r = s_remaining;       // take the first r (index order) of the keys equal to vstar
      if (tid == 0) {
        s_running = 0;
        s_cut = (s_tie == r) ? N : -1;   // every key equal to the cut value is taken: no index cut needed (the
      }                                  // usual case: distinct floats, tie group of one)
      __syncthreads();
      for (int base = 0; s_tie != r && base < N; base += LS_NT) {
        const int j = base + tid;
        const bool eq = (j < N) && (__ldcg(u + j) == vstar);
        const unsigned mm = __ballot_sync(0xffffffffu, eq);
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) s_warp[wid] = __popc(mm);
        __syncthreads();
        int before = s_running;
        for (int w = 0; w < wid; ++w) before += s_warp[w];
        const int rank = before + __popc(mm & ((1u << lane) - 1u)) + 1;   // 1-based among equals
        if (eq && rank == r) s_cut = j;
        __syncthreads();
        if (tid == 0) {
          int t = 0;
          for (int w = 0; w < LS_WARPS; ++w) t += s_warp[w];
          s_running += t;
        }
        __syncthreads();
        if (s_running >= r) break;
      }
      const uint32_t cut = (uint32_t)s_cut;
      const float* sc = p.sce + lv.off + (size_t)b * N;
      double acc = 0.0;
      for (int base = 0; base < N; base += LS_NT * 8) {             // same fixed order as before, 16 loads in flight
        uint32_t v[8];
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int j = base + k * LS_NT + tid;
          v[k] = (j < N) ? __ldcg(u + j) : 0u;
          f[k] = (j < N) ? __ldcg(sc + j) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int j = base + k * LS_NT + tid;
          if (j < N && (v[k] > vstar || (v[k] == vstar && (uint32_t)j <= cut))) acc += (double)f[k];
        }
      }
      neg_sum = block_sum(acc, s_w);
    }
  }
  if (tid == 0) {
    double pc = neg_sum, pl = 0.0;
    const int c0 = lv.cta0 + b * lv.tiles;
    for (int t = 0; t < lv.tiles; ++t) {
      pc += __ldcg(p.part_cls + c0 + t);
      pl += __ldcg(p.part_loc + c0 + t);
    }
    p.pair_cls[pair] = pc;
    p.pair_loc[pair] = pl;
    if (p.out_cls_sum) p.out_cls_sum[pair] = (float)pc;
    if (p.out_loc_sum) p.out_loc_sum[pair] = (float)pl;
    if (p.out_num_pos) p.out_num_pos[pair] = (float)npos;
    __threadfence();
    if (atomicAdd(p.pairs_done, 1) == p.L * p.B - 1) {
      // the last pair: final scalars (pipeline_anchor_basic.py:76,92-94), fixed order
      __threadfence();
      double tc = 0.0, tl = 0.0, fg = 0.0;
      for (int k = 0; k < p.L; ++k) {
        int np = 0;
        for (int bb = 0; bb < p.B; ++bb) {
          tc += __ldcg(p.pair_cls + k * p.B + bb);
          tl += __ldcg(p.pair_loc + k * p.B + bb);
          np += __ldcg(p.npos + k * p.B + bb);
        }
        fg += (double)(np > 1 ? np : 1);
      }
      p.out_scalars[0] = (float)(tc / fg);
      p.out_scalars[1] = (float)(tl / fg);
      p.out_scalars[2] = (float)fg;
    }
  }
}

struct LsLayout {
  size_t counters, part_cls, part_loc, pair_cls, pair_loc, mce, sce, total;
};

LsLayout ls_layout(int pairs, int ctas, long long anchors_total) {
  LsLayout w;
  size_t o = 0;
  w.counters = o; o += align_up((size_t)(2 * pairs + 1) * 4, 256);       // npos, done, pairs_done (zeroed per call)
  w.part_cls = o; o += align_up((size_t)ctas * 8, 256);
  w.part_loc = o; o += align_up((size_t)ctas * 8, 256);
  w.pair_cls = o; o += align_up((size_t)pairs * 8, 256);
  w.pair_loc = o; o += align_up((size_t)pairs * 8, 256);
  w.mce = o; o += align_up((size_t)anchors_total * 4, 256);
  w.sce = o; o += align_up((size_t)anchors_total * 4, 256);
  w.total = o;
  return w;
}

int ls_plan(const ssdsb_loss_level* levels, int L, int B, LsParams* p, int* ctas, long long* anchors_total) {
  SSDSB_REQUIRE(levels && L >= 1 && L <= SSDSB_MAX_LEVELS, "detection_loss: %d levels outside [1,%d]", L,
                SSDSB_MAX_LEVELS);
  SSDSB_REQUIRE(B >= 1 && B <= 65535, "detection_loss: B=%d outside [1,65535]", B);
  // CTAs are laid out level-major in the caller's order (finest level first in every reference config): the
  // big levels' end-of-pair selects then overlap the streaming pass of the smaller ones
  int cta = 0;
  long long off = 0;
  for (int i = 0; i < L; ++i) {
    const ssdsb_loss_level& s = levels[i];
    SSDSB_REQUIRE(s.A >= 1 && s.C >= 1 && s.H >= 1 && s.W >= 1 && s.stride >= 1,
                  "detection_loss: level %d has a non-positive dimension", i);
    SSDSB_REQUIRE((long long)s.A * s.H * s.W < (1ll << 31), "detection_loss: level %d has too many anchors", i);
    SSDSB_REQUIRE(s.conf && s.anchors && ((uintptr_t)s.anchors & 15) == 0,
                  "detection_loss: level %d: NULL conf/anchors or unaligned anchors", i);
  }
  for (int i = 0; i < L; ++i) {
    const ssdsb_loss_level& s = levels[i];
    LsLevel& d = p->lv[i];
    d.conf = s.conf; d.loc = s.loc; d.anchors = reinterpret_cast<const float4*>(s.anchors);
    d.depth = s.depth; d.box_target = s.box_target;
    d.A = s.A; d.C = s.C; d.H = s.H; d.W = s.W; d.stride = s.stride;
    d.HW = s.H * s.W; d.N = s.A * d.HW;
    d.tiles = (d.N + LS_NT - 1) / LS_NT;
    d.cta0 = cta;
    d.off = off;
    cta += d.tiles * B;
    off += (long long)B * d.N;
  }
  *ctas = cta;
  *anchors_total = off;
  return SSDSB_OK;
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" size_t ssdsb_detection_loss_workspace_bytes(const ssdsb_loss_level* levels, int L, int B) {
  LsParams p;
  int ctas = 0;
  long long tot = 0;
  if (ls_plan(levels, L, B, &p, &ctas, &tot) != SSDSB_OK) return 0;
  return ls_layout(L * B, ctas, tot).total;
}

extern "C" int ssdsb_detection_loss(const ssdsb_loss_level* levels, int L, int B, const float* d_targets,
                                    int T, float match_threshold, float unmatch_threshold, int cls_kind,
                                    int negpos_ratio, float alpha, float gamma, int loc_kind, float beta,
                                    float* d_out_scalars, float* d_out_cls_sum, float* d_out_loc_sum,
                                    float* d_out_num_pos, void* d_workspace, size_t workspace_bytes,
                                    void* stream) {
  LsParams p;
  int ctas = 0;
  long long tot = 0;
  int rc = ls_plan(levels, L, B, &p, &ctas, &tot);
  if (rc != SSDSB_OK) return rc;
  SSDSB_REQUIRE(T >= 0 && (T == 0 || d_targets), "detection_loss: NULL targets");
  SSDSB_REQUIRE(cls_kind == SSDSB_CLS_MULTIBOX || cls_kind == SSDSB_CLS_FOCAL, "detection_loss: cls_kind=%d", cls_kind);
  SSDSB_REQUIRE(loc_kind >= -1 && loc_kind <= LOC_CIOU, "detection_loss: loc_kind=%d", loc_kind);
  SSDSB_REQUIRE(negpos_ratio >= 0, "detection_loss: negpos_ratio=%d", negpos_ratio);
  SSDSB_REQUIRE(d_out_scalars, "detection_loss: NULL output");
  if (loc_kind >= 0)
    for (int i = 0; i < L; ++i) SSDSB_REQUIRE(levels[i].loc, "detection_loss: level %d: loc is NULL", i);
  const LsLayout w = ls_layout(L * B, ctas, tot);
  if (!d_workspace || workspace_bytes < w.total || ((uintptr_t)d_workspace & 255) != 0)
    return fail(SSDSB_ERR_WORKSPACE, "detection_loss: workspace %zu B given, %zu B (256-byte aligned) needed",
                workspace_bytes, w.total);
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  p.L = L; p.B = B; p.T = T;
  p.targets = d_targets;
  p.match_thr = match_threshold; p.unmatch_thr = unmatch_threshold;
  p.cls_kind = cls_kind; p.negpos_ratio = negpos_ratio;
  p.alpha = alpha; p.gamma = gamma;
  p.loc_kind = loc_kind; p.beta = beta;
  p.npos = reinterpret_cast<int*>(ws + w.counters);
  p.done = p.npos + L * B;
  p.pairs_done = p.done + L * B;
  p.part_cls = reinterpret_cast<double*>(ws + w.part_cls);
  p.part_loc = reinterpret_cast<double*>(ws + w.part_loc);
  p.pair_cls = reinterpret_cast<double*>(ws + w.pair_cls);
  p.pair_loc = reinterpret_cast<double*>(ws + w.pair_loc);
  p.mce = reinterpret_cast<uint32_t*>(ws + w.mce);
  p.sce = reinterpret_cast<float*>(ws + w.sce);
  p.out_scalars = d_out_scalars;
  p.out_cls_sum = d_out_cls_sum; p.out_loc_sum = d_out_loc_sum; p.out_num_pos = d_out_num_pos;
  cudaStream_t st = (cudaStream_t)stream;
  SSDSB_CUDA(cudaMemsetAsync(ws + w.counters, 0, (size_t)(2 * L * B + 1) * 4, st));
  loss_step_kernel<<<ctas, LS_NT, 0, st>>>(p);
  SSDSB_LAUNCH_CHECK("loss_step_kernel");
  return SSDSB_OK;
}
