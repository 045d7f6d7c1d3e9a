// Elementwise loss arithmetic shared by loss2.cu (per-level criteria) and loss_step.cu (the fused
// training-step loss), so both evaluate the reference's formulas with the same operation order.
// reference: ssds/core/criterion.py FocalLoss :95-108, SmoothL1Loss :138-151, IOULoss :175-239.
// Translation units including this header are built with -fmad=false.
#pragma once
#include "common.cuh"

namespace ssdsb {

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

}  // namespace ssdsb
