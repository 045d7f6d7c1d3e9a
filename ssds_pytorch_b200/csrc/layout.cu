// HBM-bound layout/pooling kernels around the conv stack (128-bit vectorised, streaming hints).
//
//   pack_image_s2d   reference ssds/ssds.py:48-57 (HWC->CHW, (x-mean)/std) fused with the 2x2
//                    space-to-depth packing the tcgen05 stem consumes: out[n,i,j,(a*2+b)*3+c] =
//                    (x[n,c,2i+a,2j+b] - mean)/std as bf16, channels 12..15 = 0.
//   maxpool3x3s2     resnet.py:45 nn.MaxPool2d(3, 2, 1) on NHWC bf16, 8 channels per thread.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace ssdsb {
namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int FMT>  // 0: fp32 NCHW, 1: uint8 NHWC
__global__ void __launch_bounds__(256)
pack_image_s2d_kernel(const void* __restrict__ src, int N, int H, int W, float mean, float stdv,
                      int row_px, int left_pad, uint4* __restrict__ out) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)N * Ho * Wo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Wo);
    const int ii = (int)((i / Wo) % Ho);
    const int n = (int)(i / ((size_t)Wo * Ho));
    float v[12];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int y = 2 * ii + a, x = 2 * j + b;
          float f;
          if (FMT == 0)
            f = __ldg(reinterpret_cast<const float*>(src) + (((size_t)n * 3 + c) * H + y) * W + x);
          else
            f = (float)__ldg(reinterpret_cast<const unsigned char*>(src) +
                             (((size_t)n * H + y) * W + x) * 3 + c);
          v[(a * 2 + b) * 3 + c] = (f - mean) / stdv;   // ssds.py:57
        }
    uint4 lo, hi;
    lo.x = pack2(v[0], v[1]); lo.y = pack2(v[2], v[3]); lo.z = pack2(v[4], v[5]); lo.w = pack2(v[6], v[7]);
    hi.x = pack2(v[8], v[9]); hi.y = pack2(v[10], v[11]); hi.z = 0u; hi.w = 0u;
    const size_t o = ((size_t)n * Ho + ii) * row_px + left_pad + j;   // pixel index in the padded rows
    out[o * 2 + 0] = lo;
    out[o * 2 + 1] = hi;
  }
}

// uint8 NHWC fast path (W % 4 == 0): a thread packs TWO horizontally adjacent s2d pixels = 2 rows x 4 source pixels
// x 3 bytes = 2 x 12 contiguous, 4-byte aligned bytes -> 6 32-bit loads instead of 24 byte loads, 4 16-byte stores.
__global__ void __launch_bounds__(256)
pack_image_s2d_u8x2_kernel(const uint32_t* __restrict__ src, int N, int H, int W, float mean, float stdv,
                           int row_px, int left_pad, uint4* __restrict__ out) {
  const int Ho = H >> 1, Wo2 = W >> 2;
  const size_t total = (size_t)N * Ho * Wo2;
  const int row_words = (W * 3) >> 2;                  // 32-bit words per source row (W % 4 == 0)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int j2 = (int)(i % Wo2);
    const int ii = (int)((i / Wo2) % Ho);
    const int n = (int)(i / ((size_t)Wo2 * Ho));
    float v[2][12];                                     // [s2d pixel][(a*2+b)*3+c]
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const uint32_t* row = src + ((size_t)n * H + 2 * ii + a) * row_words + (size_t)j2 * 3;
      const uint32_t w0 = __ldg(row), w1 = __ldg(row + 1), w2 = __ldg(row + 2);
      const uint32_t words[3] = {w0, w1, w2};
#pragma unroll
      for (int k = 0; k < 12; ++k) {                    // byte k of the 12: source pixel k/3, channel k%3
        const float f = (float)((words[k >> 2] >> ((k & 3) * 8)) & 0xffu);
        const int px = k / 3, c = k % 3;                // px 0,1 -> first s2d pixel (b = px), px 2,3 -> second
        v[px >> 1][(a * 2 + (px & 1)) * 3 + c] = (f - mean) / stdv;     // ssds.py:57
      }
    }
    const size_t o = ((size_t)n * Ho + ii) * row_px + left_pad + 2 * j2;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 lo, hi;
      lo.x = pack2(v[q][0], v[q][1]); lo.y = pack2(v[q][2], v[q][3]); lo.z = pack2(v[q][4], v[q][5]);
      lo.w = pack2(v[q][6], v[q][7]);
      hi.x = pack2(v[q][8], v[q][9]); hi.y = pack2(v[q][10], v[q][11]); hi.z = 0u; hi.w = 0u;
      out[(o + q) * 2 + 0] = lo;
      out[(o + q) * 2 + 1] = hi;
    }
  }
}

__device__ __forceinline__ uint32_t max_bf162(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}

// 3x3 / stride 2 / pad 1 max pooling, NHWC bf16, 8 channels per thread.  Separable and run-based: a
// thread owns POOL_RUN vertically consecutive outputs of one (n, wo, 8-channel) column; every input row
// is reduced horizontally once (3 loads) and shared by the two outputs it belongs to, i.e.
// (2*RUN+1)*3/RUN = 6.75 loads per output instead of 9 (the kernel is L1/L2-request bound, not DRAM bound:
// r1l launch list 172 us for 671 MB).  -inf is the identity; every window contains its centre pixel.
constexpr int POOL_RUN = 4;

__device__ __forceinline__ uint4 max_u4(const uint4& a, const uint4& b) {
  return make_uint4(max_bf162(a.x, b.x), max_bf162(a.y, b.y), max_bf162(a.z, b.z), max_bf162(a.w, b.w));
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8, int Ho, int Wo,
                    uint4* __restrict__ y) {
  const int hblocks = (Ho + POOL_RUN - 1) / POOL_RUN;
  const size_t total = (size_t)N * hblocks * Wo * C8;
  const uint4 ninf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int wo = (int)((i / C8) % Wo);
    const int hb = (int)((i / ((size_t)C8 * Wo)) % hblocks);
    const int n = (int)(i / ((size_t)C8 * Wo * hblocks));
    const int ho0 = hb * POOL_RUN;
    const int w0 = 2 * wo - 1;
    auto hmax = [&](int h) -> uint4 {          // max over the 3 horizontal taps of input row h
      if (h < 0 || h >= H) return ninf;
      const uint4* row = x + (((size_t)n * H + h) * W) * C8 + c;
      uint4 m = __ldg(row + (size_t)(w0 + 1) * C8);               // centre column always exists
      if (w0 >= 0) m = max_u4(m, __ldg(row + (size_t)w0 * C8));
      if (w0 + 2 < W) m = max_u4(m, __ldg(row + (size_t)(w0 + 2) * C8));
      return m;
    };
    uint4 carry = hmax(2 * ho0 - 1);
#pragma unroll
    for (int o = 0; o < POOL_RUN; ++o) {
      const int ho = ho0 + o;
      if (ho >= Ho) break;
      const uint4 r1 = hmax(2 * ho);
      const uint4 r2 = hmax(2 * ho + 1);
      y[(((size_t)n * Ho + ho) * Wo + wo) * C8 + c] = max_u4(max_u4(carry, r1), r2);
      carry = r2;
    }
  }
}

// fine[n,h,w,:] += coarse[n,h/2,w/2,:]   (F.interpolate(scale_factor=2, mode="nearest") + lateral,
// fpn.py:80-87), fp32 add of bf16 operands, 8 channels per thread
__global__ void __launch_bounds__(256)
upsample2x_add_kernel(const uint4* __restrict__ coarse, uint4* __restrict__ fine, int N, int H, int W,
                      int C8) {
  const size_t total = (size_t)N * H * W * C8;
  const int Hc = H >> 1, Wc = W >> 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int w = (int)((i / C8) % W);
    const int h = (int)((i / ((size_t)C8 * W)) % H);
    const int n = (int)(i / ((size_t)C8 * W * H));
    const uint4 a = fine[i];
    const uint4 b = __ldg(coarse + (((size_t)n * Hc + (h >> 1)) * Wc + (w >> 1)) * C8 + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = __uint_as_float(aw[e] << 16) + __uint_as_float(bw[e] << 16);
      const float hi = __uint_as_float(aw[e] & 0xffff0000u) + __uint_as_float(bw[e] & 0xffff0000u);
      o[e] = pack2(lo, hi);
    }
    fine[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[n,h,w,:] = cat(fine[n,h,w,:], coarse[n,h/2,w/2,:])  — YOLOv3's top-down merge
// torch.cat((features[i], F.interpolate(transforms[i](xx), scale_factor=2)), dim=1) (yolo.py:70-72) in NHWC:
// one pass, 8 channels (16 bytes) per thread, both sources read once.
__global__ void __launch_bounds__(256)
upsample2x_concat_kernel(const uint4* __restrict__ fine, const uint4* __restrict__ coarse, uint4* __restrict__ out,
                         int N, int H, int W, int Cf8, int Cc8) {
  const int Co8 = Cf8 + Cc8;
  const size_t total = (size_t)N * H * W * Co8;
  const int Hc = H >> 1, Wc = W >> 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Co8);
    const size_t pix = i / Co8;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    out[i] = (c < Cf8) ? __ldg(fine + pix * Cf8 + c)
                       : __ldg(coarse + (((size_t)n * Hc + (h >> 1)) * Wc + (w >> 1)) * Cc8 + (c - Cf8));
  }
}

// BiFPN weighted fusion (reference ssds/modeling/ssds/bifpn.py:41-62), NHWC bf16, 8 channels/thread:
//   mode 0 (top-down):  out = w0*a + w1*nearest_up2(b)            a [N,H,W,C], b [N,H/2,W/2,C]
//   mode 1 (bottom-up): out = w0*a + w1*maxpool2x2(b) [+ w2*c]    a,c [N,H,W,C], b [N,2H,2W,C]
// fp32 arithmetic in the reference's order, one rounding to bf16 at the store.
__global__ void __launch_bounds__(256)
bifpn_fuse_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                  int mode, float w0, float w1, float w2, int N, int H, int W, int C8,
                  uint4* __restrict__ out) {
  const size_t total = (size_t)N * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C8);
    const int w = (int)((i / C8) % W);
    const int h = (int)((i / ((size_t)C8 * W)) % H);
    const int n = (int)(i / ((size_t)C8 * W * H));
    const uint4 av = __ldg(a + i);
    const uint32_t as[4] = {av.x, av.y, av.z, av.w};
    float bv[8];
    if (mode == 0) {
      const uint4 q = __ldg(b + (((size_t)n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * C8 + ch);
      const uint32_t qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bv[e * 2] = __uint_as_float(qs[e] << 16);
        bv[e * 2 + 1] = __uint_as_float(qs[e] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) bv[e] = -INFINITY;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const uint4 q = __ldg(b + (((size_t)n * (2 * H) + 2 * h + dy) * (2 * W) + 2 * w + dx) * C8 + ch);
          const uint32_t qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            bv[e * 2] = fmaxf(bv[e * 2], __uint_as_float(qs[e] << 16));
            bv[e * 2 + 1] = fmaxf(bv[e * 2 + 1], __uint_as_float(qs[e] & 0xffff0000u));
          }
        }
    }
    float r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      r[e * 2] = w0 * __uint_as_float(as[e] << 16) + w1 * bv[e * 2];
      r[e * 2 + 1] = w0 * __uint_as_float(as[e] & 0xffff0000u) + w1 * bv[e * 2 + 1];
    }
    if (c) {
      const uint4 cv = __ldg(c + i);
      const uint32_t cs[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        r[e * 2] += w2 * __uint_as_float(cs[e] << 16);
        r[e * 2 + 1] += w2 * __uint_as_float(cs[e] & 0xffff0000u);
      }
    }
    out[i] = make_uint4(pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7]));
  }
}

// Depthwise 3x3 (pad 1, stride 1|2) + folded BN + ReLU/ReLU6 on NHWC bf16 (torchvision
// InvertedResidual / reference SepConvBNReLU).  One thread = one output pixel x 8 channels: nine
// 16-byte input loads (neighbouring threads hit the same lines in L1/L2), bf16 weights [9][C],
// fp32 bias [C], fp32 accumulation.  HBM-bound: in + out bytes.
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias,
                 int N, int H, int W, int C8, int stride, int Ho, int Wo, int relu, uint4* __restrict__ y) {
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int wo = (int)((i / C8) % Wo);
    const int ho = (int)((i / ((size_t)C8 * Wo)) % Ho);
    const int n = (int)(i / ((size_t)C8 * Wo * Ho));
    float acc[8];
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + c * 2);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + c * 2 + 1);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int h = ho * stride - 1 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ww = wo * stride - 1 + dx;
        if (ww < 0 || ww >= W) continue;
        const uint4 xv = __ldg(x + (((size_t)n * H + h) * W + ww) * C8 + c);
        const uint4 wv = __ldg(w + (size_t)(dy * 3 + dx) * C8 + c);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[e * 2 + 0] += __uint_as_float(xs[e] << 16) * __uint_as_float(ws[e] << 16);
          acc[e * 2 + 1] += __uint_as_float(xs[e] & 0xffff0000u) * __uint_as_float(ws[e] & 0xffff0000u);
        }
      }
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.0f);
      if (relu == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fminf(acc[e], 6.0f);
      }
    }
    y[i] = make_uint4(pack2(acc[0], acc[1]), pack2(acc[2], acc[3]), pack2(acc[4], acc[5]),
                      pack2(acc[6], acc[7]));
  }
}


// Depthwise 3x3, row-streaming variant (the one ssdsb_dwconv3x3_nhwc_bf16 launches): a thread owns 4 channels of
// one output column and walks DOWN a chunk of output rows.  Every input row it needs is loaded once (3 x 8-byte
// loads: left / centre / right column; neighbours' loads hit L1) and feeds the three output rows it touches from
// registers; the 9 x 4 folded weights live in registers for the whole chunk.
//
// The kernel is bound by load latency, not by DRAM or issue slots (r2 profiles: 1.6 TB/s at 25 % issue activity
// when each row's loads were issued right before their use), so the loop is an explicit software pipeline: a ring
// of PF raw input rows (3 x uint2 each) is always in flight ahead of the row being consumed, loads are predicated
// instructions (no branches, no out-of-range addresses dereferenced) so that nothing stops the scheduler from
// keeping them ahead, and the launch cuts the rows into as few chunks as still fill the machine (long threads, few
// waves) instead of many short ones.
//   stride 1: input row i feeds outputs i (dy 0), i-1 (dy 1), i-2 (dy 2): three accumulators whose roles rotate
//             with period 3 — PF is a multiple of 3 so the rotation is static under the unroll;
//   stride 2: output o = rows 2o-1, 2o, 2o+1; row 2o+1 is also the top row of output o+1.
// fp32 accumulation in the same order as the per-output kernel above (bias, then dy-major / dx-minor taps).
__device__ __forceinline__ void bf4_to_f(const uint2 v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}

// predicated 8-byte read-only load: zero when `pred` is false, and the address is then never dereferenced
__device__ __forceinline__ uint2 ldg_pred_u2(const uint2* p, bool pred) {
  uint2 v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
      "@q ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
      : "=r"(v.x), "=r"(v.y)
      : "l"(p), "r"((int)pred));
  return v;
}

// two fp32 lanes per 64-bit register pair: fma.rn.f32x2 = two IEEE FMAs in one issue slot (bit-identical to two
// fmaf; the kernel is issue / latency bound); ReLU / ReLU6 as max + min on the packed bf16 pair AFTER rounding
// (rounding is monotone, 0 and 6 are bf16 numbers: clamp(round(x)) == round(clamp(x)))
typedef unsigned long long dwf2_t;
__device__ __forceinline__ dwf2_t dw_pk2(float a, float b) {
  dwf2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void dw_upk2(dwf2_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ dwf2_t dw_ffma2(dwf2_t a, dwf2_t b, dwf2_t c) {
  dwf2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void bf4_to_f2(const uint2 v, dwf2_t (&f)[2]) {
  f[0] = dw_pk2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
  f[1] = dw_pk2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}
__device__ __forceinline__ uint32_t dw_clamp_bf16x2(uint32_t v, uint32_t lo2, uint32_t hi2) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\tmax.bf16x2 t, %1, %2;\n\tmin.bf16x2 %0, t, %3;\n\t}" : "=r"(r) : "r"(v), "r"(lo2), "r"(hi2));
  return r;
}

struct DwRow {
  uint2 l, m, r;
};

template <int S, int PF>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_stream_kernel(const uint2* __restrict__ x, const uint2* __restrict__ w, const float* __restrict__ bias,
                        int N, int H, int W, int C4, int Ho, int Wo, int relu, int rows_per, int chunks,
                        uint2* __restrict__ y) {
  static_assert(S == 2 || PF % 3 == 0, "stride 1: the accumulator ring has period 3");
  const unsigned total = (unsigned)N * chunks * Wo * C4;
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % (unsigned)C4);
  unsigned t = idx / (unsigned)C4;
  const int wo = (int)(t % (unsigned)Wo);
  t /= (unsigned)Wo;
  const int ch = (int)(t % (unsigned)chunks);
  const int n = (int)(t / (unsigned)chunks);
  const int ho0 = ch * rows_per;
  const int rows = min(Ho, ho0 + rows_per) - ho0;
  dwf2_t wf[9][2], b[2];
#pragma unroll
  for (int k = 0; k < 9; ++k) bf4_to_f2(__ldg(w + (size_t)k * C4 + c), wf[k]);
  {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c);
    b[0] = dw_pk2(bv.x, bv.y);
    b[1] = dw_pk2(bv.z, bv.w);
  }
  const uint32_t lo2 = relu ? 0x00000000u : 0xff80ff80u;          // bf16x2 (0 | -inf), (6 | +inf)
  const uint32_t hi2 = (relu == 2) ? 0x40c040c0u : 0x7f807f80u;
  const int w0 = wo * S - 1;                       // leftmost input column of this output column
  const bool has_l = w0 >= 0, has_r = w0 + 2 < W;  // the centre column wo*S is always inside
  const int h_first = ho0 * S - 1;                 // first input row this thread streams (may be -1)
  const int n_in = (S == 1) ? rows + 2 : 2 * rows + 1;
  const long long row_stride = (long long)W * C4;
  // `lp` walks down the input rows at column w0 (it may point outside the tensor: then the load is predicated off)
  const uint2* lp = x + ((long long)n * H + h_first) * row_stride + (long long)w0 * C4 + c;
  int li = 0;                                      // index (0 .. n_in) of the next row to load
  auto load_next = [&]() -> DwRow {
    const bool v = (li < n_in) && ((unsigned)(h_first + li) < (unsigned)H);
    DwRow r;
    r.l = ldg_pred_u2(lp, v && has_l);
    r.m = ldg_pred_u2(lp + C4, v);
    r.r = ldg_pred_u2(lp + 2 * C4, v && has_r);
    lp += row_stride;
    ++li;
    return r;
  };
  auto fma_row = [&](dwf2_t (&a)[2], const dwf2_t (&f)[3][2], int dy) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      a[0] = dw_ffma2(f[dx][0], wf[dy * 3 + dx][0], a[0]);
      a[1] = dw_ffma2(f[dx][1], wf[dy * 3 + dx][1], a[1]);
    }
  };
  auto unpack = [&](const DwRow& r, dwf2_t (&f)[3][2]) {
    bf4_to_f2(r.l, f[0]); bf4_to_f2(r.m, f[1]); bf4_to_f2(r.r, f[2]);
  };
  uint2* yp = y + ((long long)n * Ho + ho0) * ((long long)Wo * C4) + (long long)wo * C4 + c;
  const long long out_stride = (long long)Wo * C4;
  auto emit = [&](const dwf2_t (&a)[2]) {
    float o[4];
    dw_upk2(a[0], o[0], o[1]);
    dw_upk2(a[1], o[2], o[3]);
    *yp = make_uint2(dw_clamp_bf16x2(pack2(o[0], o[1]), lo2, hi2), dw_clamp_bf16x2(pack2(o[2], o[3]), lo2, hi2));
    yp += out_stride;
  };

  if (S == 1) {
    DwRow ring[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) ring[j] = load_next();
    dwf2_t acc[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      acc[a][0] = b[0];
      acc[a][1] = b[1];
    }
    for (int i0 = 0; i0 < n_in; i0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int i = i0 + j;                       // input row h_first + i
        dwf2_t f[3][2];
        unpack(ring[j], f);
        ring[j] = load_next();                      // row i + PF
        // rows outside [0, n_in) were loaded as zeros; outputs outside [0, rows) are accumulated but never emitted
        fma_row(acc[j % 3], f, 0);                  // starts output i
        fma_row(acc[(j + 2) % 3], f, 1);            // continues output i-1
        fma_row(acc[(j + 1) % 3], f, 2);            // finishes output i-2
        if (i >= 2 && i < n_in) emit(acc[(j + 1) % 3]);
        acc[(j + 1) % 3][0] = b[0];
        acc[(j + 1) % 3][1] = b[1];
      }
    }
  } else {
    dwf2_t a[2] = {b[0], b[1]};
    {
      dwf2_t f[3][2];
      const DwRow r0 = load_next();                 // row 2*ho0 - 1: the top row of the first output
      unpack(r0, f);
      fma_row(a, f, 0);
    }
    DwRow ring[2 * PF];
#pragma unroll
    for (int j = 0; j < 2 * PF; ++j) ring[j] = load_next();
    for (int o0 = 0; o0 < rows; o0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        dwf2_t f1[3][2], f2[3][2];
        unpack(ring[2 * j], f1);
        unpack(ring[2 * j + 1], f2);
        ring[2 * j] = load_next();
        ring[2 * j + 1] = load_next();
        fma_row(a, f1, 1);
        fma_row(a, f2, 2);
        if (o0 + j < rows) emit(a);
        a[0] = b[0];
        a[1] = b[1];
        fma_row(a, f2, 0);
      }
    }
  }
}

// how the launch cuts Ho output rows into chunks: as few as fill the machine evenly (a thread's 2 halo rows and its
// 36 weight loads are per chunk; waves of 148 x 512 threads should be close to whole)
inline void dw_pick_chunks(long long base_threads, int Ho, int sms, int* rows_per, int* chunks) {
  const double wave = (double)sms * 512.0;
  double best = -1.0;
  int best_rp = Ho;
  const int min_rows = Ho < 4 ? Ho : 4;
  for (int c = 1; c <= 32; ++c) {
    const int rp = (Ho + c - 1) / c;
    if (rp < min_rows) break;
    if (rp > 48 && c < 32) continue;                        // bound the length of one thread
    const int cc = (Ho + rp - 1) / rp;
    const double waves = (double)base_threads * cc / wave;
    const double full = waves <= 1.0 ? waves : waves / (double)(long long)(waves + 0.999999);
    // few waves cannot hide their ramp-up / tail: prefer >= 4
    const double depth = waves >= 4.0 ? 1.0 : 0.85 + 0.15 * waves / 4.0;
    const double score = full * depth * ((double)rp / (rp + 1.0));
    if (score > best + 1e-9) {
      best = score;
      best_rp = rp;
    }
  }
  *rows_per = best_rp;
  *chunks = (Ho + best_rp - 1) / best_rp;
}

// 5x5 / stride 1 / pad 2 max-pool on NHWC bf16 with independent channel strides for input and output: the SPP
// block of YOLOv4 (reference ssds/modeling/ssds/yolo.py:161-184: max-pools of 5, 9, 13 concatenated after x) is
// three cascaded calls (5 o 5 = 9, 5 o 9 = 13 for stride-1 max-pools with -inf padding), each reading one channel
// slice of the concatenated buffer and writing the next.  One thread = one output pixel x 8 channels; the maps are
// the smallest backbone level (<= 40 x 40), so 25 16-byte loads per thread from L1/L2 are not worth tiling.
__global__ void __launch_bounds__(256)
maxpool5x5s1_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C8, int x_cs,
                    __nv_bfloat16* __restrict__ y, int y_cs) {
  const size_t total = (size_t)N * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int w = (int)((i / C8) % W);
    const int h = (int)((i / ((size_t)C8 * W)) % H);
    const int n = (int)(i / ((size_t)C8 * W * H));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int dy = -2; dy <= 2; ++dy) {
      const int hh = h + dy;
      if (hh < 0 || hh >= H) continue;
      for (int dx = -2; dx <= 2; ++dx) {
        const int ww = w + dx;
        if (ww < 0 || ww >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + hh) * W + ww) * x_cs) + c);
        const uint32_t vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          m[e * 2] = fmaxf(m[e * 2], __uint_as_float(vs[e] << 16));
          m[e * 2 + 1] = fmaxf(m[e * 2 + 1], __uint_as_float(vs[e] & 0xffff0000u));
        }
      }
    }
    *(reinterpret_cast<uint4*>(y + (((size_t)n * H + h) * W + w) * y_cs) + c) =
        make_uint4(pack2(m[0], m[1]), pack2(m[2], m[3]), pack2(m[4], m[5]), pack2(m[6], m[7]));
  }
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_maxpool5x5s1_nhwc_bf16(const void* d_x, int N, int H, int W, int C, int x_cstride, void* d_y,
                                            int y_cstride, void* stream) {
  SSDSB_REQUIRE(d_x && d_y, "maxpool5x5s1: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool5x5s1: bad shape");
  SSDSB_REQUIRE(x_cstride >= C && y_cstride >= C && x_cstride % 8 == 0 && y_cstride % 8 == 0,
                "maxpool5x5s1: channel strides must be >= C and multiples of 8");
  SSDSB_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_y) & 15) == 0, "maxpool5x5s1: pointers must be 16-byte aligned");
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxpool5x5s1_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_x), N, H, W,
                                                                C / 8, x_cstride,
                                                                reinterpret_cast<__nv_bfloat16*>(d_y), y_cstride);
  SSDSB_LAUNCH_CHECK("maxpool5x5s1_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_bifpn_fuse_nhwc_bf16(const void* d_a, const void* d_b, const void* d_c, int mode,
                                          float w0, float w1, float w2, int N, int H, int W, int C,
                                          void* d_out, void* stream) {
  SSDSB_REQUIRE(d_a && d_b && d_out, "bifpn_fuse: NULL argument");
  SSDSB_REQUIRE(mode == 0 || mode == 1, "bifpn_fuse: mode=%d (0 top-down, 1 bottom-up)", mode);
  SSDSB_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "bifpn_fuse: bad shape");
  SSDSB_REQUIRE(mode == 1 || ((H % 2) == 0 && (W % 2) == 0), "bifpn_fuse: top-down needs an even-sized map");
  SSDSB_REQUIRE(mode == 1 || d_c == nullptr, "bifpn_fuse: the third input only exists bottom-up (mode 1)");
  SSDSB_REQUIRE((((uintptr_t)d_a | (uintptr_t)d_b | (uintptr_t)d_c | (uintptr_t)d_out) & 15) == 0,
                "bifpn_fuse: pointers must be 16-byte aligned");
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bifpn_fuse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(d_a), reinterpret_cast<const uint4*>(d_b),
      reinterpret_cast<const uint4*>(d_c), mode, w0, w1, w2, N, H, W, C / 8,
      reinterpret_cast<uint4*>(d_out));
  SSDSB_LAUNCH_CHECK("bifpn_fuse_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_dwconv3x3_nhwc_bf16(const void* d_x, const void* d_w, const float* d_bias, int N,
                                         int H, int W, int C, int stride, int relu, void* d_y,
                                         void* stream) {
  SSDSB_REQUIRE(d_x && d_w && d_bias && d_y, "dwconv3x3: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "dwconv3x3: bad shape");
  SSDSB_REQUIRE(stride == 1 || stride == 2, "dwconv3x3: stride=%d", stride);
  SSDSB_REQUIRE(relu >= 0 && relu <= 2, "dwconv3x3: relu=%d (0 none, 1 ReLU, 2 ReLU6)", relu);
  SSDSB_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_w | (uintptr_t)d_y | (uintptr_t)d_bias) & 15) == 0,
                "dwconv3x3: pointers must be 16-byte aligned");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (getenv("SSDSB_DW_SIMPLE")) {                 // the per-output variant (A/B runs, parity reference)
    const size_t total = (size_t)N * Ho * Wo * (C / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    dwconv3x3_kernel<<<blocks, 256, 0, st>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<const uint4*>(d_w), d_bias, N, H, W, C / 8,
        stride, Ho, Wo, relu, reinterpret_cast<uint4*>(d_y));
  } else {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      SSDSB_CUDA(cudaGetDevice(&dev));
      SSDSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int rows_per = 0, chunks = 0;
    dw_pick_chunks((long long)N * Wo * (C / 4), Ho, sms, &rows_per, &chunks);
    if (const char* e = getenv("SSDSB_DW_ROWS")) {          // experiment knob (profiling only)
      const int v = atoi(e);
      if (v >= 1) {
        rows_per = v < Ho ? v : Ho;
        chunks = (Ho + rows_per - 1) / rows_per;
      }
    }
    const long long total = (long long)N * chunks * Wo * (C / 4);
    SSDSB_REQUIRE(total < (1ll << 31), "dwconv3x3: tensor too large for 32-bit thread indexing");
    const unsigned blocks = (unsigned)((total + 255) / 256);
    static const int deep = getenv("SSDSB_DW_SHALLOW") ? 0 : 1;   // A/B: prefetch ring 6 vs 3 rows (s1), 3 vs 2 (s2)
    const uint2* xx = reinterpret_cast<const uint2*>(d_x);
    const uint2* ww = reinterpret_cast<const uint2*>(d_w);
    uint2* yy = reinterpret_cast<uint2*>(d_y);
    if (stride == 1) {
      if (deep)
        dwconv3x3_stream_kernel<1, 6><<<blocks, 256, 0, st>>>(xx, ww, d_bias, N, H, W, C / 4, Ho, Wo, relu, rows_per,
                                                              chunks, yy);
      else
        dwconv3x3_stream_kernel<1, 3><<<blocks, 256, 0, st>>>(xx, ww, d_bias, N, H, W, C / 4, Ho, Wo, relu, rows_per,
                                                              chunks, yy);
    } else {
      if (deep)
        dwconv3x3_stream_kernel<2, 3><<<blocks, 256, 0, st>>>(xx, ww, d_bias, N, H, W, C / 4, Ho, Wo, relu, rows_per,
                                                              chunks, yy);
      else
        dwconv3x3_stream_kernel<2, 2><<<blocks, 256, 0, st>>>(xx, ww, d_bias, N, H, W, C / 4, Ho, Wo, relu, rows_per,
                                                              chunks, yy);
    }
  }
  SSDSB_LAUNCH_CHECK("dwconv3x3_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_upsample2x_add_nhwc_bf16(const void* d_coarse, void* d_fine, int N, int H, int W,
                                              int C, void* stream) {
  SSDSB_REQUIRE(d_coarse && d_fine, "upsample2x_add: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0 && C >= 8 && C % 8 == 0,
                "upsample2x_add: fine map must be even-sized with C %% 8 == 0 (N=%d H=%d W=%d C=%d)", N, H, W, C);
  SSDSB_REQUIRE((((uintptr_t)d_coarse | (uintptr_t)d_fine) & 15) == 0,
                "upsample2x_add: pointers must be 16-byte aligned");
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(d_coarse), reinterpret_cast<uint4*>(d_fine), N, H, W, C / 8);
  SSDSB_LAUNCH_CHECK("upsample2x_add_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_upsample2x_concat_nhwc_bf16(const void* d_fine, const void* d_coarse, int N, int H, int W,
                                                 int Cf, int Cc, void* d_out, void* stream) {
  SSDSB_REQUIRE(d_fine && d_coarse && d_out, "upsample2x_concat: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0 && Cf >= 8 && Cf % 8 == 0 && Cc >= 8 &&
                    Cc % 8 == 0,
                "upsample2x_concat: the fine map must be even-sized, channel counts multiples of 8 "
                "(N=%d H=%d W=%d Cf=%d Cc=%d)", N, H, W, Cf, Cc);
  SSDSB_REQUIRE((((uintptr_t)d_fine | (uintptr_t)d_coarse | (uintptr_t)d_out) & 15) == 0,
                "upsample2x_concat: pointers must be 16-byte aligned");
  const size_t total = (size_t)N * H * W * ((Cf + Cc) / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_concat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(d_fine), reinterpret_cast<const uint4*>(d_coarse),
      reinterpret_cast<uint4*>(d_out), N, H, W, Cf / 8, Cc / 8);
  SSDSB_LAUNCH_CHECK("upsample2x_concat_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_pack_image_s2d(const void* d_src, int src_format, int N, int H, int W,
                                    float mean, float stdv, int out_row_pixels, int left_pad,
                                    void* d_out, void* stream) {
  if (out_row_pixels == 0) out_row_pixels = W / 2;
  SSDSB_REQUIRE(left_pad >= 0 && out_row_pixels >= W / 2 + left_pad,
                "pack_image_s2d: out_row_pixels=%d too small for W/2=%d + left_pad=%d", out_row_pixels,
                W / 2, left_pad);
  SSDSB_REQUIRE(d_src && d_out, "pack_image_s2d: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0,
                "pack_image_s2d: H and W must be even (N=%d H=%d W=%d)", N, H, W);
  SSDSB_REQUIRE(src_format == 0 || src_format == 1, "pack_image_s2d: src_format=%d", src_format);
  SSDSB_REQUIRE(stdv != 0.0f, "pack_image_s2d: std must be non-zero");
  SSDSB_REQUIRE(((uintptr_t)d_out & 15) == 0, "pack_image_s2d: output must be 16-byte aligned");
  const size_t total = (size_t)N * (H / 2) * (W / 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (src_format == 0)
    pack_image_s2d_kernel<0><<<blocks, 256, 0, st>>>(d_src, N, H, W, mean, stdv, out_row_pixels,
                                                     left_pad, reinterpret_cast<uint4*>(d_out));
  else if ((W % 4) == 0 && ((uintptr_t)d_src & 3) == 0) {
    const size_t total2 = (size_t)N * (H / 2) * (W / 4);
    int blocks2 = (int)((total2 + 255) / 256);
    if (blocks2 > 148 * 16) blocks2 = 148 * 16;
    pack_image_s2d_u8x2_kernel<<<blocks2, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d_src), N, H, W, mean, stdv,
                                                        out_row_pixels, left_pad, reinterpret_cast<uint4*>(d_out));
  } else
    pack_image_s2d_kernel<1><<<blocks, 256, 0, st>>>(d_src, N, H, W, mean, stdv, out_row_pixels,
                                                     left_pad, reinterpret_cast<uint4*>(d_out));
  SSDSB_LAUNCH_CHECK("pack_image_s2d_kernel");
  return SSDSB_OK;
}

extern "C" int ssdsb_maxpool3x3s2_nhwc_bf16(const void* d_x, int N, int H, int W, int C, void* d_y,
                                            void* stream) {
  SSDSB_REQUIRE(d_x && d_y, "maxpool: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool: bad shape");
  SSDSB_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_y) & 15) == 0, "maxpool: pointers must be 16-byte aligned");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)N * ((Ho + POOL_RUN - 1) / POOL_RUN) * Wo * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxpool3x3s2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(d_x), N, H, W, C / 8, Ho, Wo, reinterpret_cast<uint4*>(d_y));
  SSDSB_LAUNCH_CHECK("maxpool3x3s2_kernel");
  return SSDSB_OK;
}
