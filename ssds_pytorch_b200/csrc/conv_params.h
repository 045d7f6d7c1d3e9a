// Kernel-side parameter block of the implicit-GEMM convolution (internal).
#pragma once
#include <stdint.h>

namespace ssdsb {

enum { CONV_OUT_NHWC_BF16 = 0, CONV_OUT_HEAD_NCHW_F32 = 1 };

struct ConvKernelParams {
  int num_k_blocks;  // taps * kc_per_tap
  int kc_per_tap;    // Cin / BLOCK_K
  int KW, taps;
  int stride, pad_w, pad_h;
  int stages, n_staging, tma_store;
  int store_lag;     // TMA stores the store engine keeps in flight before re-arming a staging slot
  int chunk_cin;     // block-diagonal (grouped) conv: input-channel offset per n-tile (0 = dense)
  int ntile_cout;    // output channels covered by one n-tile (BLOCK_N, or the chunk width)
  int ways;          // M tiles processed together with interleaved MMAs (1, 2 or 4 accumulators)
  int b_resident;    // weights of this CTA's n-tile stay in shared memory for the whole kernel
  int mma_converged; // MMA warp issues from converged code through elect.sync (default) or from lane 0 only (A/B)
  int BW, BH, BN;    // output-pixel patch of one M tile (product <= 128)
  int tiles_w, tiles_h, tiles_n, n_tiles;
  int Ho, Wo, N;
  int Cout;
  int out_cstride, res_cstride;
  int relu, mode, n_loc, sigmoid;
  const float* bias;
  const void* residual;
  void* y;
  void* y2;
};

}  // namespace ssdsb
