// Per-anchor IoU arg-max matching shared by match.cu (ssdsb_match_iou) and loss_step.cu (the fused
// training-step loss): target staging + the reference's exact arithmetic, so both produce bit-identical
// depth / box targets.
//
// reference: snap_to_anchors_by_iou box.py:116-226 (see match.cu for the line-by-line notes).
// Translation units including this header are built with -fmad=false.
#pragma once
#include "common.cuh"

namespace ssdsb {

constexpr int MATCH_TCHUNK = 128;

struct Tgt {
  float x1, y1, x2, y2, area, cls;
  float sx1, sy1, sx2, sy2;  // centre-sampling region (only when radius > 0)
};

// Ordered compaction of the valid rows [t0, t0 + MATCH_TCHUNK) of one image's targets into s_t (order
// matters for first-max ties).  Called by one full warp (lanes 0..31); returns the count to lane 0's caller
// through *s_n.  r = stride * radius.
__device__ __forceinline__ void stage_targets_warp(const float* __restrict__ tg, int T, int t0, float r,
                                                   Tgt* s_t, int* s_n, int lane) {
  int n = 0;
  const int t_end = min(T, t0 + MATCH_TCHUNK);
  for (int tb = t0; tb < t_end; tb += 32) {            // warp-uniform trip count; usually ONE trip (T <= 32)
    const int t = tb + lane;
    // all five fields are fetched before the ballot: one memory latency per trip instead of two dependent ones
    float tx = 0.f, ty = 0.f, tw = 0.f, th = 0.f, tc = -1.0f;
    if (t < t_end) {
      tx = __ldg(tg + t * 5 + 0); ty = __ldg(tg + t * 5 + 1); tw = __ldg(tg + t * 5 + 2);
      th = __ldg(tg + t * 5 + 3); tc = __ldg(tg + t * 5 + 4);
    }
    const bool valid = (t < t_end) && (tc > -1.0f);
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const int slot = n + __popc(m & ((1u << lane) - 1u));
      Tgt g;
      g.x1 = tx; g.y1 = ty;
      g.x2 = tx + tw - 1.0f; g.y2 = ty + th - 1.0f;                       // box.py:162
      g.area = (g.x2 - g.x1 + 1.0f) * (g.y2 - g.y1 + 1.0f);              // box.py:166
      g.cls = tc;
      const float cx = (g.x1 + g.x2) / 2.0f, cy = (g.y1 + g.y2) / 2.0f;  // box.py:98
      g.sx1 = fmaxf(cx - r, g.x1); g.sy1 = fmaxf(cy - r, g.y1);          // box.py:105
      g.sx2 = fminf(cx + r, g.x2); g.sy2 = fminf(cy + r, g.y2);          // box.py:108
      s_t[slot] = g;
    }
    n += __popc(m);
  }
  if (lane == 0) *s_n = n;
}

struct MatchState {
  float best, bx1, by1, bx2, by2, bcls;
  bool any, inside;
};

__device__ __forceinline__ MatchState match_init() {
  MatchState m;
  m.best = -INFINITY;
  m.bx1 = m.by1 = m.bx2 = m.by2 = m.bcls = 0.f;
  m.any = false;
  m.inside = false;
  return m;
}

// fold the n staged targets into the running best match of the anchor (ax1..ay2, area aarea, centre px,py)
__device__ __forceinline__ void match_fold(MatchState& m, const Tgt* s_t, int n, float ax1, float ay1, float ax2,
                                           float ay2, float aarea, float px, float py, float radius) {
  for (int k = 0; k < n; ++k) {
    m.any = true;
    if (radius <= 0.0f && m.best >= 0.0f) {
      // Exact shortcut for disjoint pairs (the vast majority): inter == 0 gives ov = +0, -0 or NaN, none of which
      // can beat a best >= 0, so the pair cannot change the state.  x-disjoint pairs cost 2 loads + 5 ops.
      const float gx1 = s_t[k].x1, gx2 = s_t[k].x2;
      if (!(fminf(ax2, gx2) - fmaxf(ax1, gx1) + 1.0f > 0.0f)) continue;
      const float gy1 = s_t[k].y1, gy2 = s_t[k].y2;
      if (!(fminf(ay2, gy2) - fmaxf(ay1, gy1) + 1.0f > 0.0f)) continue;
    }
    const Tgt g = s_t[k];
    const float xx1 = fmaxf(ax1, g.x1), yy1 = fmaxf(ay1, g.y1);
    const float xx2 = fminf(ax2, g.x2), yy2 = fminf(ay2, g.y2);
    const float w = fmaxf(xx2 - xx1 + 1.0f, 0.0f), h = fmaxf(yy2 - yy1 + 1.0f, 0.0f);
    const float inter = w * h;
    const float uni = aarea + g.area - inter;
    // 0 / positive == +0 exactly: disjoint pairs skip the IEEE division
    const float ov = (inter == 0.0f && uni > 0.0f) ? 0.0f : inter / uni;      // box.py:168
    if (ov > m.best) {                                                        // first maximum
      m.best = ov;
      m.bx1 = g.x1; m.by1 = g.y1; m.bx2 = g.x2; m.by2 = g.y2; m.bcls = g.cls;
    }
    if (radius > 0.0f) {
      const float m4 = fminf(fminf(px - g.sx1, py - g.sy1), fminf(g.sx2 - px, g.sy2 - py));
      m.inside = m.inside || (m4 > 0.0f);
    }
  }
}

// depth (box.py:177-191) and the class plane that gets the 1 (-1: none; box.py:195-207)
__device__ __forceinline__ float match_depth(const MatchState& m, float match_thr, float unmatch_thr, float radius,
                                             int* ci) {
  *ci = -1;
  if (!m.any) return 0.0f;
  float depth = -1.0f;
  if (m.best < unmatch_thr) depth = 0.0f;
  if (m.best >= match_thr) depth = m.bcls + 1.0f;
  if (radius > 0.0f) depth = fminf(depth, m.inside ? 1.0f : 0.0f);
  *ci = (m.best < unmatch_thr) ? -1 : (int)m.bcls;   // .long() truncation, box.py:201-203
  return depth;
}

// box2delta(best box, anchor)  box.py:61-71
__device__ __forceinline__ void match_delta(const MatchState& m, float ax1, float ay1, float ax2, float ay2,
                                            float (&d)[4]) {
  const float aw = ax2 - ax1 + 1.0f, ah = ay2 - ay1 + 1.0f;
  const float acx = ax1 + 0.5f * aw, acy = ay1 + 0.5f * ah;
  const float bw = m.bx2 - m.bx1 + 1.0f, bh = m.by2 - m.by1 + 1.0f;
  const float bcx = m.bx1 + 0.5f * bw, bcy = m.by1 + 0.5f * bh;
  d[0] = (bcx - acx) / aw;
  d[1] = (bcy - acy) / ah;
  d[2] = (float)log((double)(bw / aw));
  d[3] = (float)log((double)(bh / ah));
}

}  // namespace ssdsb
