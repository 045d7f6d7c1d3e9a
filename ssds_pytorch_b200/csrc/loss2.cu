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
#include "loss_math.cuh"

namespace ssdsb {
namespace {

constexpr int L2_NT = 256;
constexpr int RED_NT = 1024;

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

// ------------------------------------------------------------------------------------------------
// backward: d/d(logits) and d/d(pred deltas) of   sum_b scale[b] * (masked per-image loss sum)
// i.e. of the quantity the fused *_sum entry points return, with the caller's 1/fg_targets (and the
// upstream gradient) folded into scale[b].  Same chain as autograd on the reference modules.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(L2_NT)
focal_backward_kernel(const float* __restrict__ logits, const float* __restrict__ depth,
                      const float* __restrict__ scale, int A, int C, int HW, float alpha, float gamma,
                      float* __restrict__ grad) {
  const int b = blockIdx.y;
  const int N = A * HW;
  const int i = blockIdx.x * L2_NT + threadIdx.x;
  if (i >= N) return;
  const int a = i / HW, yx = i % HW;
  const size_t off = ((size_t)b * A + a) * C * HW + yx;
  const float d = __ldg(depth + (size_t)b * N + i);
  const bool live = d >= 0.0f;                       // (depth >= 0) mask of the caller
  const int cpos = (d > 0.0f) ? (int)d - 1 : -1;
  const float sc = __ldg(scale + b);
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    float g = 0.0f;
    if (live) {
      const float x = __ldcs(logits + off + (size_t)c * HW);
      const float t = (c == cpos) ? 1.0f : 0.0f;
      const float p = 1.0f / (1.0f + expf(-x));
      const float ce = bce_logits2(x, t);
      const float a_t = t * alpha + (1.0f - t) * (1.0f - alpha);
      const float q = (t == 1.0f) ? 1.0f - p : p;            // 1 - p_t
      const float dq = (t == 1.0f) ? -p * (1.0f - p) : p * (1.0f - p);
      float w, dw;                                            // q^gamma and d/dq
      if (gamma == 2.0f) {
        w = q * q;
        dw = 2.0f * q;
      } else {
        w = powf(q, gamma);
        dw = gamma * powf(q, gamma - 1.0f);
      }
      g = a_t * (w * (p - t) + ce * dw * dq) * sc;
    }
    __stcs(grad + off + (size_t)c * HW, g);
  }
}

// forward-mode dual number carrying the 4 partials w.r.t. the predicted deltas: the IoU-family losses
// are differentiated by running the forward formula on duals (exact, no hand-derived expressions).
// Tie / boundary conventions are torch's: maximum/minimum split the gradient evenly on ties, clamp
// passes the gradient on [lo, hi] inclusive, comparisons that gate a product are constants.
struct Dual {
  float v, d[4];
};
__device__ __forceinline__ Dual dconst(float v) { return Dual{v, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ Dual dvar(float v, int k) {
  Dual r = dconst(v);
  r.d[k] = 1.0f;
  return r;
}
__device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b) {
  return Dual{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2], a.d[3] + b.d[3]}};
}
__device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b) {
  return Dual{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2], a.d[3] - b.d[3]}};
}
__device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b) {
  Dual r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
__device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b) {
  Dual r;
  r.v = a.v / b.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v;
  return r;
}
__device__ __forceinline__ Dual dscale(const Dual& a, float s) {
  return Dual{a.v * s, {a.d[0] * s, a.d[1] * s, a.d[2] * s, a.d[3] * s}};
}
__device__ __forceinline__ Dual dexp(const Dual& a) {
  const float e = expf(a.v);
  return Dual{e, {a.d[0] * e, a.d[1] * e, a.d[2] * e, a.d[3] * e}};
}
__device__ __forceinline__ Dual datan(const Dual& a) {
  const float s = 1.0f / (1.0f + a.v * a.v);
  return Dual{atanf(a.v), {a.d[0] * s, a.d[1] * s, a.d[2] * s, a.d[3] * s}};
}
__device__ __forceinline__ Dual dmax(const Dual& a, float c) {     // maximum(a, constant)
  if (a.v > c) return a;
  if (a.v < c) return dconst(c);
  return Dual{c, {0.5f * a.d[0], 0.5f * a.d[1], 0.5f * a.d[2], 0.5f * a.d[3]}};
}
__device__ __forceinline__ Dual dmin(const Dual& a, float c) {
  if (a.v < c) return a;
  if (a.v > c) return dconst(c);
  return Dual{c, {0.5f * a.d[0], 0.5f * a.d[1], 0.5f * a.d[2], 0.5f * a.d[3]}};
}
__device__ __forceinline__ Dual dclamp(const Dual& a, float lo, float hi) {
  Dual r = a;
  r.v = clampf(a.v, lo, hi);
  if (!(a.v >= lo && a.v <= hi)) r.d[0] = r.d[1] = r.d[2] = r.d[3] = 0.0f;
  return r;
}

__device__ __forceinline__ Dual iou_family_dual(const float (&pv)[4], const float (&t)[4], int type) {
  const Dual px = dvar(pv[0], 0), py = dvar(pv[1], 1);
  const Dual pw = dexp(dvar(pv[2], 2)), ph = dexp(dvar(pv[3], 3));
  const float tw = expf(t[2]), th = expf(t[3]);
  const Dual plx = px - dscale(pw, 0.5f), ply = py - dscale(ph, 0.5f);
  const Dual prx = px + dscale(pw, 0.5f), pry = py + dscale(ph, 0.5f);
  const float tlx = t[0] - 0.5f * tw, tly = t[1] - 0.5f * th, trx = t[0] + 0.5f * tw, try_ = t[1] + 0.5f * th;
  const Dual lx = dmax(plx, tlx), ly = dmax(ply, tly), rx = dmin(prx, trx), ry = dmin(pry, try_);
  const float en = (lx.v < rx.v && ly.v < ry.v) ? 1.0f : 0.0f;
  const Dual area_i = dscale((rx - lx) * (ry - ly), en);
  const Dual area_a = pw * ph;
  const float area_b = tw * th;
  const Dual area_u = area_a + dconst(area_b) - area_i;
  const Dual iou = (area_i + dconst(1e-7f)) / (area_u + dconst(1e-7f));
  if (type == LOC_IOU) return dconst(1.0f) - dclamp(iou, 0.0f, 1.0f);
  const Dual olx = dmin(plx, tlx), oly = dmin(ply, tly), orx = dmax(prx, trx), ory = dmax(pry, try_);
  if (type == LOC_GIOU) {
    const float eo = (olx.v < orx.v && oly.v < ory.v) ? 1.0f : 0.0f;
    const Dual area_o = dscale((orx - olx) * (ory - oly), eo) + dconst(1e-7f);
    const Dual g = iou - (area_o - area_u) / area_o;
    return dconst(1.0f) - dclamp(g, -1.0f, 1.0f);
  }
  const Dual dx = px - dconst(t[0]), dy = py - dconst(t[1]);
  const Dual inter_diag = dx * dx + dy * dy;
  const Dual ox = orx - olx, oy = ory - oly;
  const Dual outer_diag = ox * ox + oy * oy + dconst(1e-7f);
  if (type == LOC_DIOU) {
    const Dual dd = iou - inter_diag / outer_diag;
    return dconst(1.0f) - dclamp(dd, -1.0f, 1.0f);
  }
  const Dual da = dconst(atanf(tw / th)) - datan(pw / ph);
  const Dual v = dscale(da * da, (float)(4.0 / (3.14159265358979323846 * 3.14159265358979323846)));
  const float al = v.v / ((1.0f - iou.v) + v.v);          // alpha is computed under no_grad (criterion.py:221-223)
  const Dual c = iou - (inter_diag / outer_diag + dscale(v, al));
  return dconst(1.0f) - dclamp(c, -1.0f, 1.0f);
}

__global__ void __launch_bounds__(L2_NT)
loc_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                    const float* __restrict__ depth, const float* __restrict__ scale, int A, int HW, int type,
                    float beta, float* __restrict__ grad) {
  const int b = blockIdx.y;
  const int N = A * HW;
  const int i = blockIdx.x * L2_NT + threadIdx.x;
  if (i >= N) return;
  const int a = i / HW, yx = i % HW;
  const size_t off = ((size_t)b * A + a) * 4 * HW + yx;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (__ldg(depth + (size_t)b * N + i) > 0.0f) {
    const float sc = __ldg(scale + b);
    float p[4], t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      p[k] = __ldcs(pred + off + (size_t)k * HW);
      t[k] = __ldcs(target + off + (size_t)k * HW);
    }
    if (type == LOC_SMOOTHL1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float dlt = p[k] - t[k];
        const float x = fabsf(dlt);
        // where(x >= beta, x - beta/2, x^2 / (2 beta)); d|d|/dd = sign(d) with sign(0) = 0
        g[k] = ((x >= beta) ? ((dlt > 0.0f) - (dlt < 0.0f)) : dlt / beta) * sc;
      }
    } else {
      const Dual l = iou_family_dual(p, t, type);
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = l.d[k] * sc;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) __stcs(grad + off + (size_t)k * HW, g[k]);
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

extern "C" int ssdsb_focal_loss_sum_backward(const float* d_logits, const float* d_depth, int B, int A, int C,
                                             int H, int W, float alpha, float gamma, const float* d_scale,
                                             float* d_grad_logits, void* stream) {
  int rc = check_shape("focal_loss_sum_backward", B, A, C, H, W);
  if (rc != SSDSB_OK) return rc;
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_logits && d_depth && d_scale && d_grad_logits, "focal_loss_sum_backward: NULL argument");
  const int HW = H * W, N = A * HW;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  focal_backward_kernel<<<grid, L2_NT, 0, (cudaStream_t)stream>>>(d_logits, d_depth, d_scale, A, C, HW, alpha, gamma,
                                                                   d_grad_logits);
  SSDSB_LAUNCH_CHECK("focal_backward_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_loc_loss_sum_backward(const float* d_pred, const float* d_target, const float* d_depth,
                                           int B, int A, int H, int W, int type, float beta,
                                           const float* d_scale, float* d_grad_pred, void* stream) {
  int rc = check_shape("loc_loss_sum_backward", B, A, 4, H, W);
  if (rc != SSDSB_OK) return rc;
  SSDSB_REQUIRE(type >= LOC_SMOOTHL1 && type <= LOC_CIOU, "loc_loss_sum_backward: type=%d outside [0,4]", type);
  SSDSB_REQUIRE(type != LOC_SMOOTHL1 || beta > 0.0f, "loc_loss_sum_backward: beta must be > 0");
  if (B == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_pred && d_target && d_depth && d_scale && d_grad_pred, "loc_loss_sum_backward: NULL argument");
  const int HW = H * W, N = A * HW;
  dim3 grid((N + L2_NT - 1) / L2_NT, B);
  loc_backward_kernel<<<grid, L2_NT, 0, (cudaStream_t)stream>>>(d_pred, d_target, d_depth, d_scale, A, HW, type, beta,
                                                                 d_grad_pred);
  SSDSB_LAUNCH_CHECK("loc_backward_kernel");
  return SSDSB_OK;
}
