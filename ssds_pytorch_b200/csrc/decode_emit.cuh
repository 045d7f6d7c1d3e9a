// Turning a selected (score, flat index) key of one level into the reference's decode outputs — shared by
// decode.cu (top_n <= 1024) and decode_large.cu (single-pass large top_n).  reference: box.py:443-471.
// Translation units including this header are built with -fmad=false.
#pragma once
#include "common.cuh"

namespace ssdsb {

__device__ __forceinline__ float clampf_nanprop(float t, float lo, float hi) {
  // torch.max(m, torch.min(t, M)) — NaN propagates
  return (t != t) ? t : fmaxf(lo, fminf(t, hi));
}

// writes slot `o` of the concatenated outputs from `key` (valid = false: the zero padding row)
__device__ __forceinline__ void emit_detection(const ssdsb_level& lv, int b, bool valid, unsigned long long key,
                                               int rescore, size_t o, float* __restrict__ out_scores,
                                               float* __restrict__ out_boxes, float* __restrict__ out_classes,
                                               int32_t* __restrict__ out_index) {
  float score = 0.f, cls = 0.f, x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  int32_t flat = -1;
  if (valid) {
    const int W = lv.W, H = lv.H, C = lv.C;
    const int HW = H * W;
    const float stride_f = (float)lv.stride;
    const float Mx = (float)W * stride_f - 1.0f;  // box.py:83  size=[W,H] * stride - 1
    const float My = (float)H * stride_f - 1.0f;
    const float* loc = lv.loc + (size_t)b * lv.A * 4 * HW;
    score = key_score(key);
    const uint32_t idx = key_index(key);
    flat = (int32_t)idx;
    const int x = idx % W;                       // box.py:452-454
    const int y = (idx / W) % H;
    const int c = (idx / W / H) % C;             // box.py:448
    const int a = idx / C / H / W;
    cls = (float)c;
    const float d0 = __ldg(loc + (size_t)(a * 4 + 0) * HW + y * W + x);
    const float d1 = __ldg(loc + (size_t)(a * 4 + 1) * HW + y * W + x);
    const float d2 = __ldg(loc + (size_t)(a * 4 + 2) * HW + y * W + x);
    const float d3 = __ldg(loc + (size_t)(a * 4 + 3) * HW + y * W + x);
    const float4 an = __ldg(reinterpret_cast<const float4*>(lv.anchors) + a);
    // grid anchor: (x,y,x,y)*stride + anchors[a]   box.py:459-462
    const float gx1 = (float)x * stride_f + an.x, gy1 = (float)y * stride_f + an.y;
    const float gx2 = (float)x * stride_f + an.z, gy2 = (float)y * stride_f + an.w;
    // delta2box  box.py:74-87
    const float aw = gx2 - gx1 + 1.0f, ah = gy2 - gy1 + 1.0f;
    const float cx = gx1 + 0.5f * aw, cy = gy1 + 0.5f * ah;
    const float pcx = d0 * aw + cx, pcy = d1 * ah + cy;
    const float pw = (float)exp((double)d2) * aw, ph = (float)exp((double)d3) * ah;
    x1 = clampf_nanprop(pcx - 0.5f * pw, 0.0f, Mx);
    y1 = clampf_nanprop(pcy - 0.5f * ph, 0.0f, My);
    x2 = clampf_nanprop(pcx + 0.5f * pw - 1.0f, 0.0f, Mx);
    y2 = clampf_nanprop(pcy + 0.5f * ph - 1.0f, 0.0f, My);
    if (rescore) {  // box.py:464-471
      const float gcx = (gx1 + gx2) / 2.0f, gcy = (gy1 + gy2) / 2.0f;
      const float ltx = fabsf(gcx - x1), lty = fabsf(gcy - y1);
      const float rbx = fabsf(x2 - gcx), rby = fabsf(y2 - gcy);
      const float qx = fminf(ltx, rbx) / fmaxf(ltx, rbx);
      const float qy = fminf(lty, rby) / fmaxf(lty, rby);
      score = score * sqrtf(qx * qy);
    }
  }
  out_scores[o] = score;
  out_classes[o] = cls;
  reinterpret_cast<float4*>(out_boxes)[o] = make_float4(x1, y1, x2, y2);
  if (out_index) out_index[o] = flat;
}

}  // namespace ssdsb
