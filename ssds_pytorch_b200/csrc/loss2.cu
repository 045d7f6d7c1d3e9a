// Focal / SmoothL1 / IoU-family losses (forward) — the "next" row of SURVEY 8f, fused the same way as
// MultiBoxLoss: one pass over the logits / deltas, optional per-image reduction without materialising
// one-hot targets or masks.
//
// reference: ssds/core/criterion.py
//   FocalLoss.forward    :95-108   alpha_t * (1 - p_t)^gamma * BCEWithLogits
//   SmoothL1Loss.forward :138-151  where(|d| >= beta, |d| - beta/2, d^2 / (2 beta))
//   IOULoss.forward      :175-239  iou / giou / diou / ciou on (x, y, log w, log h) deltas
// and the caller-side masking / summation of pipeline_anchor_basic.py:76-97:
//   cls: sum(loss * (depth >= 0)), loc: sum(loss * (depth > 0)), both per level, / fg count.
// Layouts: logits [B,A,C,H,W], deltas [B,A,4,H,W], depth [B,A,1,H,W] (one thread per anchor, x fastest,
// so every channel-plane access of a warp is one coalesced line).
#include "common.cuh"

namespace ssdsb {
namespace {

constexpr int L2_NT = 256;
constexpr int RED_NT = 1024;

__device__ __forceinline__ float bce_logits2(float x, float t) {
  const float ls = fminf(x, 0.0f) - log1pf(expf(-fabsf(x)));
  return (1.0f - t) * x - ls;
}
__device__ __forceinline__ float focal_term(float x, float t, float alpha, float gamma) {
  const float p = 1.0f / (1.0f + expf(-x));                 // pred_logits.sigmoid()
  const float ce = bce_logits2(x, t);
  const float a = t * alpha + (1.0f - t) * (1.0f - alpha);
  const float pt = (t == 1.0f) ? p : 1.0f - p;
  const float q = 1.0f - pt;
  const float w = (gamma == 2.0f) ? q * q : powf(q, gamma);
  return a * w * ce;
}

// MODE 0: unreduced [B,A,C,H,W] from logits + target; MODE 1: per-anchor sum with the class from depth
template <int MODE>
__global__ void __launch_bounds__(L2_NT)
focal_kernel(const float* __restrict__ logits, const float* __restrict__ target,
             const float* __restrict__ depth, int A, int C, int HW, float alpha, float gamma,
             float* __restrict__ out, float* __restrict__ per_anchor) {
  const int b = blockIdx.y;
  const int N = A * HW;
  const int i = blockIdx.x * L2_NT + threadIdx.x;
  if (i >= N) return;
  const int a = i / HW, yx = i % HW;
  const size_t off = ((size_t)b * A + a) * C * HW + yx;
  int cpos = -1;
  if (MODE == 1) {
    const float d = __ldg(depth + (size_t)b * N + i);
    cpos = (d > 0.0f) ? (int)d - 1 : -1;
  }
  float sum = 0.0f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float x = __ldcs(logits + off + (size_t)c * HW);
    const float t = (MODE == 0) ? __ldcs(target + off + (size_t)c * HW) : ((c == cpos) ? 1.0f : 0.0f);
    const float fl = focal_term(x, t, alpha, gamma);
    if (MODE == 0) out[off + (size_t)c * HW] = fl;
    else sum += fl;
  }
  if (MODE == 1) per_anchor[(size_t)b * N + i] = sum;
}

enum { LOC_SMOOTHL1 = 0, LOC_IOU = 1, LOC_GIOU = 2, LOC_DIOU = 3, LOC_CIOU = 4 };

__device__ __forceinline__ float smooth_l1(float p, float t, float beta) {
  const float x = fabsf(p - t);
  return (x >= beta) ? x - 0.5f * beta : 0.5f * x * x / beta;
}

// torch.clamp semantics: NaN passes through (fminf/fmaxf would drop it).  The reference's ciou is NaN for
// identical boxes (v = 0, 1 - iou = 0 -> alpha = 0/0); that is reproduced, not repaired.
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

__device__ __forceinline__ float iou_family(const float (&p)[4], const float (&t)[4], int type) {
  // delta2ltrb (criterion.py:233-239): ctr = d[:2], wh = exp(d[2:])
  const float pw = expf(p[2]), ph = expf(p[3]), tw = expf(t[2]), th = expf(t[3]);
  const float plx = p[0] - 0.5f * pw, ply = p[1] - 0.5f * ph, prx = p[0] + 0.5f * pw, pry = p[1] + 0.5f * ph;
  const float tlx = t[0] - 0.5f * tw, tly = t[1] - 0.5f * th, trx = t[0] + 0.5f * tw, try_ = t[1] + 0.5f * th;
  const float lx = fmaxf(plx, tlx), ly = fmaxf(ply, tly), rx = fminf(prx, trx), ry = fminf(pry, try_);
  const float area_i = ((rx - lx) * (ry - ly)) * ((lx < rx && ly < ry) ? 1.0f : 0.0f);
  const float area_a = pw * ph, area_b = tw * th;
  const float area_u = area_a + area_b - area_i;
  const float iou = (area_i + 1e-7f) / (area_u + 1e-7f);
  if (type == LOC_IOU) return 1.0f - clampf(iou, 0.0f, 1.0f);
  const float olx = fminf(plx, tlx), oly = fminf(ply, tly), orx = fmaxf(prx, trx), ory = fmaxf(pry, try_);
  if (type == LOC_GIOU) {
    const float area_o = ((orx - olx) * (ory - oly)) * ((olx < orx && oly < ory) ? 1.0f : 0.0f) + 1e-7f;
    const float g = iou - (area_o - area_u) / area_o;
    return 1.0f - clampf(g, -1.0f, 1.0f);
  }
  const float dx = p[0] - t[0], dy = p[1] - t[1];
  const float inter_diag = dx * dx + dy * dy;
  const float ox = orx - olx, oy = ory - oly;
  const float outer_diag = (ox * ox + oy * oy) + 1e-7f;
  if (type == LOC_DIOU) {
    const float d = iou - inter_diag / outer_diag;
    return 1.0f - clampf(d, -1.0f, 1.0f);
  }
  const float da = atanf(tw / th) - atanf(pw / ph);
  const float v = (float)(4.0 / (3.14159265358979323846 * 3.14159265358979323846)) * (da * da);
  const float S = 1.0f - iou;
  const float al = v / (S + v);
  const float c = iou - (inter_diag / outer_diag + al * v);
  return 1.0f - clampf(c, -1.0f, 1.0f);
}

// REDUCED 0: unreduced output ([B,A,4,H,W] for SmoothL1, [B,A,1,H,W] otherwise);
// REDUCED 1: per-anchor value (sum of the 4 SmoothL1 terms), 0 where depth <= 0
template <int REDUCED>
__global__ void __launch_bounds__(L2_NT)
loc_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                const float* __restrict__ depth, int A, int HW, int type, float beta,
                float* __restrict__ out, float* __restrict__ per_anchor) {
  const int b = blockIdx.y;
  const int N = A * HW;
  const int i = blockIdx.x * L2_NT + threadIdx.x;
  if (i >= N) return;
  const int a = i / HW, yx = i % HW;
  const size_t off = ((size_t)b * A + a) * 4 * HW + yx;
  if (REDUCED == 1) {
    if (!(__ldg(depth + (size_t)b * N + i) > 0.0f)) {
      per_anchor[(size_t)b * N + i] = 0.0f;
      return;
    }
  }
  float p[4], t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    p[k] = __ldcs(pred + off + (size_t)k * HW);
    t[k] = __ldcs(target + off + (size_t)k * HW);
  }
  if (type == LOC_SMOOTHL1) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float l = smooth_l1(p[k], t[k], beta);
      if (REDUCED == 0) out[off + (size_t)k * HW] = l;
      s += l;
    }
    if (REDUCED == 1) per_anchor[(size_t)b * N + i] = s;
  } else {
    const float l = iou_family(p, t, type);
    if (REDUCED == 0) out[((size_t)b * A + a) * HW + yx] = l;
    else per_anchor[(size_t)b * N + i] = l;
  }
}

// per image: sum of per_anchor[i] over anchors with depth >= min_depth (0 for cls, already masked for
// loc -> pass -inf), deterministic (fixed order, double accumulation); also counts positives
__global__ void __launch_bounds__(RED_NT)
masked_sum_kernel(const float* __restrict__ per_anchor, const float* __restrict__ depth, int N,
                  int need_nonneg, float* __restrict__ loss_sum, float* __restrict__ num_pos) {
  __shared__ double s_part[RED_NT / 32];
  __shared__ int s_cnt[RED_NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  int cnt = 0;
  for (int i = tid; i < N; i += RED_NT) {
    const float d = __ldg(depth + (size_t)b * N + i);
    if (!need_nonneg || d >= 0.0f) acc += (double)__ldg(per_anchor + (size_t)b * N + i);
    cnt += (d > 0.0f) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_down_sync(0xffffffffu, acc, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if ((tid & 31) == 0) {
    s_part[tid >> 5] = acc;
    s_cnt[tid >> 5] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    int c = 0;
    for (int w = 0; w < RED_NT / 32; ++w) {
      t += s_part[w];
      c += s_cnt[w];
    }
    loss_sum[b] = (float)t;
    if (num_pos) num_pos[b] = (float)c;
  }
}

int check_shape(const char* what, int B, int A, int C, int H, int W) {
  SSDSB_REQUIRE(B >= 0 && A >= 1 && C >= 1 && H >= 1 && W >= 1, "%s: bad shape B=%d A=%d C=%d H=%d W=%d", what,
                B, A, C, H, W);
  SSDSB_REQUIRE((long long)A * H * W < (1ll << 31) && B <= 65535, "%s: too many anchors / images", what);
  return SSDSB_OK;
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" size_t ssdsb_loss_sum_workspace_bytes(int B, int A, int H, int W) {
  if (B < 0 || A < 1 || H < 1 || W < 1) return 0;
  return align_up((size_t)B * A * H * W * 4, 16);
}

extern "C" int ssdsb_focal_loss(const float* d_logits, const float* d_target, int B, int A, int C, int H,
                                int W, float alpha, float gamma, float* d_out, void* stream) {
  int rc = check_shape("focal_loss", B, A, C, H, W);
  if (rc != SSDSB_OK) return rc;
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_logits && d_target && d_out, "focal_loss: NULL argument");
  const int HW = H * W, N = A * HW;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  focal_kernel<0><<<grid, L2_NT, 0, (cudaStream_t)stream>>>(d_logits, d_target, nullptr, A, C, HW, alpha,
                                                             gamma, d_out, nullptr);
  SSDSB_LAUNCH_CHECK("focal_kernel<0>");
  return SSDSB_OK;
}

extern "C" int ssdsb_focal_loss_sum(const float* d_logits, const float* d_depth, int B, int A, int C, int H,
                                    int W, float alpha, float gamma, float* d_loss_sum, float* d_num_pos,
                                    void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape("focal_loss_sum", B, A, C, H, W);
  if (rc != SSDSB_OK) return rc;
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_logits && d_depth && d_loss_sum && d_num_pos, "focal_loss_sum: NULL argument");
  const size_t need = ssdsb_loss_sum_workspace_bytes(B, A, H, W);
  if (!d_workspace || workspace_bytes < need)
    return fail(SSDSB_ERR_WORKSPACE, "focal_loss_sum: workspace %zu B given, %zu B needed", workspace_bytes, need);
  const int HW = H * W, N = A * HW;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  float* pa = reinterpret_cast<float*>(d_workspace);
  focal_kernel<1><<<grid, L2_NT, 0, st>>>(d_logits, nullptr, d_depth, A, C, HW, alpha, gamma, nullptr, pa);
  SSDSB_LAUNCH_CHECK("focal_kernel<1>");
  masked_sum_kernel<<<B, RED_NT, 0, st>>>(pa, d_depth, N, 1, d_loss_sum, d_num_pos);
  SSDSB_LAUNCH_CHECK("masked_sum_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_loc_loss(const float* d_pred, const float* d_target, int B, int A, int H, int W,
                              int type, float beta, float* d_out, void* stream) {
  int rc = check_shape("loc_loss", B, A, 4, H, W);
  if (rc != SSDSB_OK) return rc;
  SSDSB_REQUIRE(type >= LOC_SMOOTHL1 && type <= LOC_CIOU, "loc_loss: type=%d outside [0,4]", type);
  SSDSB_REQUIRE(type != LOC_SMOOTHL1 || beta > 0.0f, "loc_loss: beta must be > 0");
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_pred && d_target && d_out, "loc_loss: NULL argument");
  const int HW = H * W, N = A * HW;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  loc_loss_kernel<0><<<grid, L2_NT, 0, (cudaStream_t)stream>>>(d_pred, d_target, nullptr, A, HW, type, beta,
                                                                d_out, nullptr);
  SSDSB_LAUNCH_CHECK("loc_loss_kernel<0>");
  return SSDSB_OK;
}

extern "C" int ssdsb_loc_loss_sum(const float* d_pred, const float* d_target, const float* d_depth, int B,
                                  int A, int H, int W, int type, float beta, float* d_loss_sum,
                                  void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape("loc_loss_sum", B, A, 4, H, W);
  if (rc != SSDSB_OK) return rc;
  SSDSB_REQUIRE(type >= LOC_SMOOTHL1 && type <= LOC_CIOU, "loc_loss_sum: type=%d outside [0,4]", type);
  SSDSB_REQUIRE(type != LOC_SMOOTHL1 || beta > 0.0f, "loc_loss_sum: beta must be > 0");
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_pred && d_target && d_depth && d_loss_sum, "loc_loss_sum: NULL argument");
  const size_t need = ssdsb_loss_sum_workspace_bytes(B, A, H, W);
  if (!d_workspace || workspace_bytes < need)
    return fail(SSDSB_ERR_WORKSPACE, "loc_loss_sum: workspace %zu B given, %zu B needed", workspace_bytes, need);
  const int HW = H * W, N = A * HW;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  float* pa = reinterpret_cast<float*>(d_workspace);
  loc_loss_kernel<1><<<grid, L2_NT, 0, st>>>(d_pred, d_target, d_depth, A, HW, type, beta, nullptr, pa);
  SSDSB_LAUNCH_CHECK("loc_loss_kernel<1>");
  masked_sum_kernel<<<B, RED_NT, 0, st>>>(pa, d_depth, N, 0, d_loss_sum, nullptr);
  SSDSB_LAUNCH_CHECK("masked_sum_kernel");
  return SSDSB_OK;
}
