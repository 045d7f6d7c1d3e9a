// Anchors ("PriorBox") and the box codec — write-/stream-bound elementwise kernels, 128-bit I/O.
//
// reference: ssds/modeling/layers/box.py
//   generate_anchors :46-58, grid materialisation :151-159, box2delta :61-71, delta2box :74-87.
// fp32, same operation order as the torch expressions (TU built with -fmad=false); rintf is
// round-half-even like torch.round (stride 15 / ratio 0.5 -> hs = rint(10.5) = 10).
#include "common.cuh"

namespace ssdsb {
namespace {

constexpr int MAX_RS = 32;
struct AnchorSpec {
  float ratios[MAX_RS];
  float scales[MAX_RS];
  int n_ratios, n_scales, stride;
};

__global__ void base_anchor_kernel(const __grid_constant__ AnchorSpec s, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int A = s.n_ratios * s.n_scales;
  if (i >= A) return;
  const float scale = s.scales[i / s.n_ratios];  // scale-major (box.py:49-50)
  const float ratio = s.ratios[i % s.n_ratios];  // ratio-minor (box.py:51)
  const float st = (float)s.stride;
  const float ws = rintf(sqrtf(st * st / ratio));  // box.py:54
  const float hs = rintf(ws * ratio);              // box.py:55
  float4 a;
  a.x = 0.5f * (st - ws * scale);                  // box.py:56
  a.y = 0.5f * (st - hs * scale);
  a.z = 0.5f * (st + ws * scale) - 1.0f;           // box.py:57
  a.w = 0.5f * (st + hs * scale) - 1.0f;
  out[i] = a;
}

// out[a][x][y] = base[a] + (x,y,x,y)*stride   — x-major like torch.meshgrid(ij) in box.py:151-159
__global__ void anchor_grid_kernel(const float4* __restrict__ base, int A, int stride, int W, int H,
                                   float4* __restrict__ out) {
  const size_t total = (size_t)A * W * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i % H);
    const int x = (int)((i / H) % W);
    const int a = (int)(i / ((size_t)H * W));
    const float4 b = __ldg(base + a);
    const float fx = (float)(x * stride), fy = (float)(y * stride);
    __stcs(out + i, make_float4(fx + b.x, fy + b.y, fx + b.z, fy + b.w));
  }
}

__global__ void box2delta_kernel(const float4* __restrict__ boxes, const float4* __restrict__ anchors,
                                 int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 b = boxes[i], a = anchors[i];
  const float aw = a.z - a.x + 1.0f, ah = a.w - a.y + 1.0f;
  const float acx = a.x + 0.5f * aw, acy = a.y + 0.5f * ah;
  const float bw = b.z - b.x + 1.0f, bh = b.w - b.y + 1.0f;
  const float bcx = b.x + 0.5f * bw, bcy = b.y + 0.5f * bh;
  out[i] = make_float4((bcx - acx) / aw, (bcy - acy) / ah, (float)log((double)(bw / aw)),
                       (float)log((double)(bh / ah)));
}

__device__ __forceinline__ float clampf_nanprop(float t, float lo, float hi) {
  return (t != t) ? t : fmaxf(lo, fminf(t, hi));
}

__global__ void delta2box_kernel(const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                                 int n, float Mx, float My, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d = deltas[i], a = anchors[i];
  const float aw = a.z - a.x + 1.0f, ah = a.w - a.y + 1.0f;
  const float cx = a.x + 0.5f * aw, cy = a.y + 0.5f * ah;
  const float pcx = d.x * aw + cx, pcy = d.y * ah + cy;
  const float pw = (float)exp((double)d.z) * aw, ph = (float)exp((double)d.w) * ah;
  out[i] = make_float4(clampf_nanprop(pcx - 0.5f * pw, 0.f, Mx), clampf_nanprop(pcy - 0.5f * ph, 0.f, My),
                       clampf_nanprop(pcx + 0.5f * pw - 1.0f, 0.f, Mx),
                       clampf_nanprop(pcy + 0.5f * ph - 1.0f, 0.f, My));
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_generate_anchors(int stride, const float* h_ratios, int n_ratios,
                                      const float* h_scales, int n_scales, float* d_out,
                                      void* stream) {
  SSDSB_REQUIRE(stride >= 1, "generate_anchors: stride=%d", stride);
  SSDSB_REQUIRE(h_ratios && h_scales && d_out, "generate_anchors: NULL argument");
  SSDSB_REQUIRE(n_ratios >= 1 && n_ratios <= MAX_RS && n_scales >= 1 && n_scales <= MAX_RS,
                "generate_anchors: n_ratios=%d / n_scales=%d outside [1,%d]", n_ratios, n_scales,
                MAX_RS);
  SSDSB_REQUIRE(((uintptr_t)d_out & 15) == 0, "generate_anchors: output must be 16-byte aligned");
  AnchorSpec s;
  for (int i = 0; i < n_ratios; ++i) s.ratios[i] = h_ratios[i];
  for (int i = 0; i < n_scales; ++i) s.scales[i] = h_scales[i];
  s.n_ratios = n_ratios;
  s.n_scales = n_scales;
  s.stride = stride;
  const int A = n_ratios * n_scales;
  base_anchor_kernel<<<(A + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      s, reinterpret_cast<float4*>(d_out));
  SSDSB_LAUNCH_CHECK("base_anchor_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_anchor_grid(const float* d_base, int A, int stride, int W, int H, float* d_out,
                                 void* stream) {
  SSDSB_REQUIRE(d_base && d_out, "anchor_grid: NULL argument");
  SSDSB_REQUIRE(A >= 1 && stride >= 1 && W >= 1 && H >= 1, "anchor_grid: non-positive dimension");
  SSDSB_REQUIRE(((uintptr_t)d_base & 15) == 0 && ((uintptr_t)d_out & 15) == 0,
                "anchor_grid: pointers must be 16-byte aligned");
  const size_t total = (size_t)A * W * H;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  anchor_grid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(d_base), A, stride, W, H, reinterpret_cast<float4*>(d_out));
  SSDSB_LAUNCH_CHECK("anchor_grid_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_box2delta(const float* d_boxes, const float* d_anchors, int n, float* d_out,
                               void* stream) {
  SSDSB_REQUIRE(n >= 0, "box2delta: n=%d", n);
  if (n == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_boxes && d_anchors && d_out, "box2delta: NULL argument");
  SSDSB_REQUIRE((((uintptr_t)d_boxes | (uintptr_t)d_anchors | (uintptr_t)d_out) & 15) == 0,
                "box2delta: pointers must be 16-byte aligned");
  box2delta_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(d_boxes), reinterpret_cast<const float4*>(d_anchors), n,
      reinterpret_cast<float4*>(d_out));
  SSDSB_LAUNCH_CHECK("box2delta_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_delta2box(const float* d_deltas, const float* d_anchors, int n, int size_w,
                               int size_h, int stride, float* d_out, void* stream) {
  SSDSB_REQUIRE(n >= 0, "delta2box: n=%d", n);
  if (n == 0) return SSDSB_OK;
  SSDSB_REQUIRE(d_deltas && d_anchors && d_out, "delta2box: NULL argument");
  SSDSB_REQUIRE((((uintptr_t)d_deltas | (uintptr_t)d_anchors | (uintptr_t)d_out) & 15) == 0,
                "delta2box: pointers must be 16-byte aligned");
  const float Mx = (float)size_w * (float)stride - 1.0f;
  const float My = (float)size_h * (float)stride - 1.0f;
  delta2box_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(d_deltas), reinterpret_cast<const float4*>(d_anchors), n, Mx,
      My, reinterpret_cast<float4*>(d_out));
  SSDSB_LAUNCH_CHECK("delta2box_kernel");
  return SSDSB_OK;
}
