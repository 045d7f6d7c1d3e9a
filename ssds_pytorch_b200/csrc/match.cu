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
#include "match_core.cuh"

namespace ssdsb {
namespace {

constexpr int MATCH_NT = 256;

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

  MatchState m = match_init();
  const float* tg = targets + (size_t)b * T * 5;
  const float r = (float)((double)stride * (double)radius);  // python: stride * radius

  for (int t0 = 0; t0 < T; t0 += MATCH_TCHUNK) {
    __syncthreads();
    if (threadIdx.x < 32) stage_targets_warp(tg, T, t0, r, s_t, &s_n, threadIdx.x);
    __syncthreads();
    match_fold(m, s_t, s_n, ax1, ay1, ax2, ay2, aarea, px, py, radius);
  }
  if (!active) return;

  float d[4] = {0.f, 0.f, 0.f, 0.f};
  int ci = -1;  // class plane that gets the 1 (none by default)
  const float depth = match_depth(m, match_thr, unmatch_thr, radius, &ci);
  if (m.any) match_delta(m, ax1, ay1, ax2, ay2, d);
  const size_t ba_off = (size_t)b * A + a;
  depth_t[ba_off * HW + yx] = depth;
  float* bt = box_t + ba_off * 4 * HW + yx;
  bt[0] = d[0]; bt[(size_t)HW] = d[1]; bt[(size_t)2 * HW] = d[2]; bt[(size_t)3 * HW] = d[3];
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
