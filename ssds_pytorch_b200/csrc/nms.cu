// Batched class-aware (D)IoU NMS — one CTA per image.
//
// Semantics follow reference ssds/modeling/layers/box.py:480-546 exactly:
//   * candidates with score <= 0 (or NaN) are dropped (:496);
//   * the rest are visited in descending score order (:505; ties -> ascending input position);
//   * a later candidate j is removed by a kept pivot i iff  class[j] == class[i]  and
//     NOT (iou(i,j) <= thr)  (:532), iou = inter / (area_j + area_i - inter + 1e-7) with +1 pixel
//     widths (:507,:520-521); with DIoU  iou = clamp(iou - d2/c2, -1, 1), d2 measured between the
//     TOP-LEFT corners (:527), c2 = outer diagonal^2 + 1e-7 (:528-529);
//   * at most `ndetections` pivots are taken (:512); rows are zero padded (:489-491).
// All arithmetic is fp32 in the reference's operation order (the TU is built with -fmad=false), so
// keep/suppress decisions are bit-identical to the torch CPU path for finite inputs.
//
// Algorithm (B200): keys = (ordered score << 32 | ~index) are streamed once from HBM/L2 into a
// shared-memory running top-SEL buffer (block bitonic prune), sorted, and consumed in chunks of
// CH=128 candidates: (1) every chunk candidate is tested against the pivots kept so far,
// (2) a CHxCH suppression bit-matrix is built by all warps, (3) warp 0 scans it sequentially.
// The loop stops as soon as `ndetections` pivots exist, so the common case touches one chunk.
// If a round's SEL candidates are exhausted before D pivots are found, another selection round
// runs over keys below the last one processed (rare; bounded by N/SEL rounds).
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "decode_large.h"

namespace ssdsb {
namespace {

constexpr int NMS_NT = 512;    // threads per CTA
constexpr int NMS_EPT = 4;     // scores per thread per streaming tile
constexpr int NMS_TILE = NMS_NT * NMS_EPT;
constexpr int NMS_SEL = 2048;  // candidates sorted per round
constexpr int NMS_CAP = 4096;  // shared key buffer (>= SEL + TILE)
constexpr int NMS_CH = 128;    // chunk width of the suppression matrix
constexpr int NMS_CW = NMS_CH / 32;
constexpr int NMS_PRESEL_MIN = 8192;   // rows longer than this get their first SEL candidates from topk_rows

struct Cand {
  float x1, y1, x2, y2;
};

__device__ __forceinline__ bool suppresses(const Cand& p, float pcls, float parea, const Cand& c,
                                           float ccls, float carea, float thr, bool diou) {
  if (ccls != pcls) return false;
  float xx1 = fmaxf(c.x1, p.x1), yy1 = fmaxf(c.y1, p.y1);
  float xx2 = fminf(c.x2, p.x2), yy2 = fminf(c.y2, p.y2);
  float w = fmaxf(xx2 - xx1 + 1.0f, 0.0f);
  float h = fmaxf(yy2 - yy1 + 1.0f, 0.0f);
  float inter = w * h;
  // exact shortcut: disjoint boxes give iou = +0 (denominator > 0), the DIoU term only lowers it, so
  // "iou <= thr" holds for every thr >= 0 — skip the two IEEE divisions
  if (inter == 0.0f && thr >= 0.0f && (carea + parea) > 0.0f) return false;
  float iou = inter / (carea + parea - inter + 1e-7f);
  if (diou) {
    float olx = fminf(c.x1, p.x1), oly = fminf(c.y1, p.y1);
    float orx = fmaxf(c.x2, p.x2), ory = fmaxf(c.y2, p.y2);
    float dx = c.x1 - p.x1, dy = c.y1 - p.y1;
    float inter_diag = dx * dx + dy * dy;
    float ox = orx - olx, oy = ory - oly;
    float outer_diag = (ox * ox + oy * oy) + 1e-7f;
    float v = iou - inter_diag / outer_diag;
    iou = (v != v) ? v : fminf(fmaxf(v, -1.0f), 1.0f);
  }
  return !(iou <= thr);
}

__global__ void __launch_bounds__(NMS_NT, 1)
nms_kernel(const float* __restrict__ scores, const float* __restrict__ boxes,
           const float* __restrict__ classes, int N, float thr, int D, int diou,
           float* __restrict__ out_scores, float* __restrict__ out_boxes,
           float* __restrict__ out_classes, int32_t* __restrict__ out_index, float* __restrict__ out_packed,
           const unsigned long long* __restrict__ presel_keys, const int* __restrict__ presel_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);  // [NMS_CAP]
  unsigned long long* keys2 = keys + NMS_CAP;                                   // [NMS_SEL] prune scratch
  Cand* kbox = reinterpret_cast<Cand*>(keys2 + NMS_SEL);                        // [D]
  float* kcls = reinterpret_cast<float*>(kbox + D);                             // [D]
  float* karea = kcls + D;                                                      // [D]

  __shared__ Cand cbox[NMS_CH];
  __shared__ float ccls[NMS_CH], carea[NMS_CH], cscore[NMS_CH];
  __shared__ uint32_t cidx[NMS_CH];
  __shared__ __align__(16) uint32_t mat[NMS_CH][NMS_CW];
  __shared__ uint32_t alive_w[NMS_CW];
  __shared__ int keptpos[NMS_CH];
  __shared__ int s_cnt, s_nk;
  __shared__ unsigned long long s_thr, s_kth;
  __shared__ int s_scratch[260];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* sc = scores + (size_t)b * N;
  const Cand* bx = reinterpret_cast<const Cand*>(boxes) + (size_t)b * N;
  const float* cl = classes + (size_t)b * N;

  int nkept = 0;
  unsigned long long upper = ~0ull;  // process keys strictly below this
  bool exhausted = false;

  bool first = true;
  while (nkept < D && !exhausted) {
    int count;
    if (first && presel_keys) {
      // ---------------- round 0 of a long row: the sorted top-SEL keys were selected by topk_rows ----------------
      count = presel_count[b];
      for (int i = tid; i < count; i += NMS_NT) keys[i] = presel_keys[(size_t)b * NMS_SEL + i];
      if (count < NMS_SEL) exhausted = true;          // every candidate with score > 0 is in the list
      __syncthreads();
    } else {
      // ---------------- selection round: top-SEL keys below `upper` ----------------
      if (tid == 0) {
        s_cnt = 0;
        s_thr = 0ull;
      }
      __syncthreads();
      bool pruned = false;
      for (int base = 0; base < N; base += NMS_TILE) {
        unsigned long long k[NMS_EPT];
        bool take[NMS_EPT];
        const unsigned long long cur = s_thr;
#pragma unroll
        for (int e = 0; e < NMS_EPT; ++e) {
          int i = base + e * NMS_NT + tid;
          float s = (i < N) ? __ldg(sc + i) : 0.0f;
          k[e] = make_key(s, (uint32_t)i);
          take[e] = (s > 0.0f) && (k[e] < upper) && (k[e] > cur);
        }
        const int fill = topk_append<NMS_EPT>(keys, &s_cnt, k, take);
        if (__syncthreads_or(fill > NMS_CAP - NMS_TILE)) {  // block-uniform, race-free
          topk_prune_select<NMS_NT>(keys, keys2, &s_cnt, &s_thr, NMS_SEL, s_scratch, &s_kth);
          pruned = true;
        }
      }
      topk_prune<NMS_NT>(keys, &s_cnt, &s_thr, NMS_SEL);
      count = s_cnt;  // sorted, descending, in keys[0..count)
      if (!pruned && count <= NMS_SEL) exhausted = true;  // nothing was ever discarded
    }
    first = false;
    if (count == 0) break;

    // ---------------- consume in chunks of CH ----------------
    int pos = 0;
    while (pos < count && nkept < D) {
      const int m = min(NMS_CH, count - pos);
      if (tid < m) {
        unsigned long long key = keys[pos + tid];
        uint32_t idx = key_index(key);
        Cand c = bx[idx];
        cbox[tid] = c;
        ccls[tid] = __ldg(cl + idx);
        cscore[tid] = key_score(key);
        cidx[tid] = idx;
        carea[tid] = (c.x2 - c.x1 + 1.0f) * (c.y2 - c.y1 + 1.0f);
      }
      if (tid < NMS_CW) {
        int lo = tid * 32;
        int nb = m - lo;
        alive_w[tid] = nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
      }
      __syncthreads();

      // (1) against pivots kept in earlier chunks: 4 threads per candidate stride the pivot list
      {
        const int j = tid & (NMS_CH - 1);
        const int q = tid / NMS_CH;  // 0..3
        if (j < m && nkept > 0) {
          Cand c = cbox[j];
          float cc = ccls[j], ca = carea[j];
          bool dead = false;
          for (int k = q; k < nkept && !dead; k += NMS_NT / NMS_CH)
            dead = suppresses(kbox[k], kcls[k], karea[k], c, cc, ca, thr, diou != 0);
          if (dead) atomicAnd(&alive_w[j >> 5], ~(1u << (j & 31)));
        }
      }
      // (2) suppression matrix inside the chunk: thread -> (row i, word w)
      {
        const int i = tid / NMS_CW;
        const int w = tid % NMS_CW;
        uint32_t bits = 0;
        if (i < m && w * 32 + 31 > i) {          // words entirely at or before i hold no j > i
          Cand p = cbox[i];
          float pc = ccls[i], pa = carea[i];
          const int j0 = w * 32;
          for (int jj = 0; jj < 32; ++jj) {
            int j = j0 + jj;
            if (j > i && j < m) {
              if (suppresses(p, pc, pa, cbox[j], ccls[j], carea[j], thr, diou != 0))
                bits |= (1u << jj);
            }
          }
        }
        mat[i][w] = bits;
      }
      __syncthreads();

      // (3) sequential scan by ONE thread over the alive bit-mask (4 x 32 bits in registers): only
      //     alive candidates are visited (find-first-set), each costs one 128-bit row load + 4 ANDs
      if (tid == 0) {
        uint32_t a0 = alive_w[0], a1 = alive_w[1], a2 = alive_w[2], a3 = alive_w[3];
        int nk = 0;
        const int room = D - nkept;
#define NMS_SCAN_WORD(W, AW)                                                        \
        while (AW != 0u && nk < room) {                                             \
          const int bit = __ffs(AW) - 1;                                            \
          const int i = (W) * 32 + bit;                                             \
          keptpos[nk++] = i;                                                        \
          const uint4 row = *reinterpret_cast<const uint4*>(&mat[i][0]);            \
          a0 &= ~row.x; a1 &= ~row.y; a2 &= ~row.z; a3 &= ~row.w;                   \
          AW &= ~(1u << bit);                                                       \
        }
        NMS_SCAN_WORD(0, a0)
        NMS_SCAN_WORD(1, a1)
        NMS_SCAN_WORD(2, a2)
        NMS_SCAN_WORD(3, a3)
#undef NMS_SCAN_WORD
        s_nk = nk;
      }
      __syncthreads();

      // (4) append the new pivots to the output and to the pivot list
      const int nk = s_nk;
      if (tid < nk) {
        const int j = keptpos[tid];
        const int o = nkept + tid;
        kbox[o] = cbox[j];
        kcls[o] = ccls[j];
        karea[o] = carea[j];
        const size_t ob = (size_t)b * D + o;
        if (out_scores) {
          out_scores[ob] = cscore[j];
          reinterpret_cast<Cand*>(out_boxes)[ob] = cbox[j];
          out_classes[ob] = ccls[j];
        }
        if (out_index) out_index[ob] = (int32_t)cidx[j];
        if (out_packed) {      // [B,D,6] = (score, x1, y1, x2, y2, class): the block SSDDetector ships / all-gathers
          float* pk = out_packed + ob * 6;
          pk[0] = cscore[j];
          pk[1] = cbox[j].x1; pk[2] = cbox[j].y1; pk[3] = cbox[j].x2; pk[4] = cbox[j].y2;
          pk[5] = ccls[j];
        }
      }
      nkept += nk;
      pos += m;
      __syncthreads();
    }
    upper = keys[count - 1];
    __syncthreads();
  }
  (void)s_kth;

  // zero padding of the tail (reference rows start as torch.zeros)
  for (int o = nkept + tid; o < D; o += NMS_NT) {
    const size_t ob = (size_t)b * D + o;
    if (out_scores) {
      out_scores[ob] = 0.0f;
      reinterpret_cast<Cand*>(out_boxes)[ob] = Cand{0.f, 0.f, 0.f, 0.f};
      out_classes[ob] = 0.0f;
    }
    if (out_index) out_index[ob] = -1;
    if (out_packed) {
      float* pk = out_packed + ob * 6;
#pragma unroll
      for (int e = 0; e < 6; ++e) pk[e] = 0.0f;
    }
  }
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

static size_t nms_presel_bytes(int B) {
  return align_up((size_t)B * NMS_SEL * 8, 256) + align_up((size_t)B * 4, 256);
}

extern "C" size_t ssdsb_nms_workspace_bytes(int B, int N, int ndetections) {
  (void)ndetections;
  if (B < 1 || N <= NMS_PRESEL_MIN) return 0;  // short rows: everything lives in shared memory
  // long rows: the first NMS_SEL candidates of every row come from a multi-CTA exact selection (decode_large.cu)
  return nms_presel_bytes(B) + topk_rows_workspace_bytes(B, N, NMS_SEL) + 256;
}

extern "C" int ssdsb_nms(const float* d_scores, const float* d_boxes, const float* d_classes,
                         int B, int N, float nms_threshold, int ndetections, int using_diou,
                         float* d_out_scores, float* d_out_boxes, float* d_out_classes,
                         int32_t* d_out_index, float* d_out_packed, void* d_workspace, size_t workspace_bytes,
                         void* stream) {
  SSDSB_REQUIRE(B >= 0 && N >= 0, "nms: negative size (B=%d, N=%d)", B, N);
  SSDSB_REQUIRE(ndetections >= 1 && ndetections <= 4096, "nms: ndetections=%d outside [1,4096]",
                ndetections);
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE((d_out_scores && d_out_boxes && d_out_classes) || (d_out_packed && !d_out_scores && !d_out_boxes &&
                                                                     !d_out_classes),
                "nms: give the three separate outputs, the packed output, or both");
  SSDSB_REQUIRE(N == 0 || (d_scores && d_boxes && d_classes), "nms: NULL input");
  SSDSB_REQUIRE(((uintptr_t)d_boxes & 15) == 0 && (!d_out_boxes || ((uintptr_t)d_out_boxes & 15) == 0),
                "nms: boxes must be 16-byte aligned");
  const size_t smem = sizeof(unsigned long long) * (NMS_CAP + NMS_SEL) + (size_t)ndetections * (16 + 4 + 4);
  static_assert(NMS_CAP >= NMS_SEL + NMS_TILE, "buffer too small");
  SSDSB_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long* presel_keys = nullptr;
  const int* presel_count = nullptr;
  const size_t need = ssdsb_nms_workspace_bytes(B, N, ndetections);
  if (need > 0 && d_workspace && workspace_bytes >= need && !getenv("SSDSB_NMS_NO_PRESEL")) {
    // (callers that pass no workspace keep the single-CTA streaming selection: same results, slower on long rows)
    unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)d_workspace + 255) & ~(uintptr_t)255);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
    int* count = reinterpret_cast<int*>(ws + align_up((size_t)B * NMS_SEL * 8, 256));
    unsigned char* rest = ws + nms_presel_bytes(B);
    int rc = topk_rows(d_scores, B, N, FLT_TRUE_MIN, NMS_SEL, keys, count, rest,
                       workspace_bytes - (size_t)(rest - reinterpret_cast<unsigned char*>(d_workspace)), st);
    if (rc != SSDSB_OK) return rc;
    presel_keys = keys;
    presel_count = count;
  }
  nms_kernel<<<B, NMS_NT, smem, st>>>(
      d_scores, d_boxes, d_classes, N, nms_threshold, ndetections, using_diou, d_out_scores,
      d_out_boxes, d_out_classes, d_out_index, d_out_packed, presel_keys, presel_count);
  SSDSB_LAUNCH_CHECK("nms_kernel");
  return SSDSB_OK;
}
