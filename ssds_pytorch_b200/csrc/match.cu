// match: IoU arg-max matcher ("jaccard" + "match") + encode + depth + one-hot class target.
//
// reference: extract_targets box.py:362-405 -> snap_to_anchors_by_iou box.py:116-226
//   * padded rows (label <= -1) are dropped (:375); an image without targets gives all zeros
//     (:132-145);
//   * targets (x,y,w,h) -> (x1,y1,x1+w-1,y1+h-1) (:162); IoU with +1 pixel areas, no eps (:163-168);
//   * best target per anchor = first maximum (:171); box_target = box2delta for EVERY anchor (:172);
//   * depth = -1, 0 if iou < unmatch, class+1 if iou >= match (:177-181); optional ATSS centre
//     sampling: depth = min(depth, inside) (:184-191, get_sample_region :90-113);
//   * one-hot class target, background (iou < unmatch) has no class (:195-207).
// Output layouts are the reference's: cls [B,A,C,H,W], box [B,A,4,H,W], depth [B,A,1,H,W].
//
// The kernel is write-bound (4*(C+5) bytes per anchor): one thread owns one (a,y,x) anchor with x
// fastest, so each of the C+5 plane stores of a warp is one coalesced 128-byte line; targets are
// staged once per CTA in shared memory.
#include "common.cuh"

namespace ssdsb {
namespace {

constexpr int MATCH_NT = 256;
constexpr int MATCH_TCHUNK = 128;

struct Tgt {
  float x1, y1, x2, y2, area, cls;
  float sx1, sy1, sx2, sy2;  // centre-sampling region (only when radius > 0)
};

__global__ void __launch_bounds__(MATCH_NT)
match_kernel(const float* __restrict__ targets, int T, const float4* __restrict__ base, int A, int C,
             int stride, int H, int W, float match_thr, float unmatch_thr, float radius,
             float* __restrict__ cls_t, float* __restrict__ box_t, float* __restrict__ depth_t) {
  __shared__ Tgt s_t[MATCH_TCHUNK];
  __shared__ int s_n;
  const int b = blockIdx.y;
  const int HW = H * W;
  const int N = A * HW;
  const int i = blockIdx.x * MATCH_NT + threadIdx.x;  // (a, y, x), x fastest
  const bool active = i < N;
  const int a = active ? i / HW : 0;
  const int yx = active ? i % HW : 0;
  const int y = yx / W, x = yx % W;

  const float4 ba = __ldg(base + a);
  const float fx = (float)(x * stride), fy = (float)(y * stride);
  const float ax1 = fx + ba.x, ay1 = fy + ba.y, ax2 = fx + ba.z, ay2 = fy + ba.w;
  const float aarea = (ax2 - ax1 + 1.0f) * (ay2 - ay1 + 1.0f);
  const float px = fx + (float)(stride / 2), py = fy + (float)(stride / 2);  // box.py:185

  float best = -INFINITY;
  float bx1 = 0.f, by1 = 0.f, bx2 = 0.f, by2 = 0.f, bcls = 0.f;
  bool any = false, inside = false;
  const float* tg = targets + (size_t)b * T * 5;
  const float r = (float)((double)stride * (double)radius);  // python: stride * radius

  for (int t0 = 0; t0 < T; t0 += MATCH_TCHUNK) {
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // ordered compaction of the valid rows of this chunk (order matters for first-max ties)
    if (threadIdx.x < 32) {
      int n = 0;
      for (int tt = threadIdx.x; tt < MATCH_TCHUNK; tt += 32) {
        const int t = t0 + tt;
        const bool valid = (t < T) && (tg[t * 5 + 4] > -1.0f);
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          const int slot = n + __popc(m & ((1u << threadIdx.x) - 1u));
          const float tx = tg[t * 5 + 0], ty = tg[t * 5 + 1], tw = tg[t * 5 + 2], th = tg[t * 5 + 3];
          Tgt g;
          g.x1 = tx; g.y1 = ty;
          g.x2 = tx + tw - 1.0f; g.y2 = ty + th - 1.0f;                       // box.py:162
          g.area = (g.x2 - g.x1 + 1.0f) * (g.y2 - g.y1 + 1.0f);              // box.py:166
          g.cls = tg[t * 5 + 4];
          const float cx = (g.x1 + g.x2) / 2.0f, cy = (g.y1 + g.y2) / 2.0f;  // box.py:98
          g.sx1 = fmaxf(cx - r, g.x1); g.sy1 = fmaxf(cy - r, g.y1);          // box.py:105
          g.sx2 = fminf(cx + r, g.x2); g.sy2 = fminf(cy + r, g.y2);          // box.py:108
          s_t[slot] = g;
        }
        n += __popc(m);
      }
      if (threadIdx.x == 0) s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    for (int k = 0; k < n; ++k) {
      const Tgt g = s_t[k];
      any = true;
      const float xx1 = fmaxf(ax1, g.x1), yy1 = fmaxf(ay1, g.y1);
      const float xx2 = fminf(ax2, g.x2), yy2 = fminf(ay2, g.y2);
      const float w = fmaxf(xx2 - xx1 + 1.0f, 0.0f), h = fmaxf(yy2 - yy1 + 1.0f, 0.0f);
      const float inter = w * h;
      const float ov = inter / (aarea + g.area - inter);                      // box.py:168
      if (ov > best) {                                                        // first maximum
        best = ov;
        bx1 = g.x1; by1 = g.y1; bx2 = g.x2; by2 = g.y2; bcls = g.cls;
      }
      if (radius > 0.0f) {
        const float m4 = fminf(fminf(px - g.sx1, py - g.sy1), fminf(g.sx2 - px, g.sy2 - py));
        inside = inside || (m4 > 0.0f);
      }
    }
  }
  if (!active) return;

  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, depth = 0.f;
  int ci = -1;  // class plane that gets the 1 (none by default)
  if (any) {
    // box2delta(best box, anchor)  box.py:61-71
    const float aw = ax2 - ax1 + 1.0f, ah = ay2 - ay1 + 1.0f;
    const float acx = ax1 + 0.5f * aw, acy = ay1 + 0.5f * ah;
    const float bw = bx2 - bx1 + 1.0f, bh = by2 - by1 + 1.0f;
    const float bcx = bx1 + 0.5f * bw, bcy = by1 + 0.5f * bh;
    d0 = (bcx - acx) / aw;
    d1 = (bcy - acy) / ah;
    d2 = (float)log((double)(bw / aw));
    d3 = (float)log((double)(bh / ah));
    depth = -1.0f;
    if (best < unmatch_thr) depth = 0.0f;
    if (best >= match_thr) depth = bcls + 1.0f;
    if (radius > 0.0f) depth = fminf(depth, inside ? 1.0f : 0.0f);
    ci = (best < unmatch_thr) ? -1 : (int)bcls;   // .long() truncation, box.py:201-203
  }
  const size_t ba_off = (size_t)b * A + a;
  depth_t[ba_off * HW + yx] = depth;
  float* bt = box_t + ba_off * 4 * HW + yx;
  bt[0] = d0; bt[(size_t)HW] = d1; bt[(size_t)2 * HW] = d2; bt[(size_t)3 * HW] = d3;
  if (cls_t) {
    float* ct = cls_t + ba_off * C * HW + yx;
    for (int c = 0; c < C; ++c) __stcs(ct + (size_t)c * HW, (c == ci) ? 1.0f : 0.0f);
  }
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_match_iou(const float* d_targets, int B, int T, const float* d_base_anchors,
                               int A, int C, int stride, int H, int W, float match_threshold,
                               float unmatch_threshold, float center_sampling_radius,
                               float* d_cls_target, float* d_box_target, float* d_depth,
                               void* stream) {
  SSDSB_REQUIRE(B >= 0 && T >= 0, "match: negative size");
  SSDSB_REQUIRE(A >= 1 && C >= 1 && stride >= 1 && H >= 1 && W >= 1,
                "match: non-positive dimension (A=%d C=%d stride=%d H=%d W=%d)", A, C, stride, H, W);
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_base_anchors && d_box_target && d_depth, "match: NULL argument");
  SSDSB_REQUIRE(T == 0 || d_targets, "match: NULL targets");
  SSDSB_REQUIRE(((uintptr_t)d_base_anchors & 15) == 0, "match: anchors must be 16-byte aligned");
  SSDSB_REQUIRE((long long)A * H * W < (1ll << 31), "match: too many anchors");
  SSDSB_REQUIRE(B <= 65535, "match: B=%d > 65535", B);
  const int N = A * H * W;
  dim3 grid((N + MATCH_NT - 1) / MATCH_NT, B);
  match_kernel<<<grid, MATCH_NT, 0, (cudaStream_t)stream>>>(
      d_targets, T, reinterpret_cast<const float4*>(d_base_anchors), A, C, stride, H, W,
      match_threshold, unmatch_threshold, center_sampling_radius, d_cls_target, d_box_target,
      d_depth);
  SSDSB_LAUNCH_CHECK("match_kernel");
  return SSDSB_OK;
}
