// Fused MobileNetV2 inverted residual (reference ssds/modeling/nets/mobilenet.py:40-76 — torchvision's
// InvertedResidual:  [1x1 expand + BN + ReLU6] -> 3x3 depthwise (stride 1|2) + BN + ReLU6 -> 1x1 project + BN
// [+ x]) in ONE launch, sm_100a only.  The 6x-expanded tensor never leaves the SM:
//
//   per output tile (BW x BH pixels of one image) and per chunk of HC hidden channels
//     TMA      x patch ((BW-1)S+3) x ((BH-1)S+3) pixels x Cin, zero-filled outside the image      -> smem
//     tcgen05  expand GEMM  [patch pixels x Cin] x [Cin x HC]  (PM = ceil(patch/128) M tiles)      -> TMEM
//     convert  4 warps: TMEM -> +bias -> ReLU6 -> bf16, ZERO for pixels outside the image
//              (the depthwise conv pads the EXPANDED tensor with zeros)                            -> smem "hpatch"
//     dw       8 warps: 3x3 depthwise from hpatch on the CUDA cores (row streaming, fp32, same tap order as
//              layout.cu's depthwise kernels) -> +bias -> ReLU6 -> bf16, written as the K-major swizzled
//              A operand of the project GEMM                                                       -> smem
//     tcgen05  project GEMM  [128 x HC] x [HC x Cout], accumulated over the chunks                 -> TMEM
//   per tile  epilogue (the convert warps): TMEM -> +bias [+ residual tile, TMA-loaded] -> bf16 -> staged TMA store.
//
// Rounding points are exactly those of the three separate launches (bf16 after expand, after depthwise, after
// project; fp32 accumulation in the same K order), so the result is bit-identical to
// ssdsb_conv2d_bf16 -> ssdsb_dwconv3x3_nhwc_bf16 -> ssdsb_conv2d_bf16 (tests/test_gpu_conv.py).
// Without an expand layer (t = 1, the first block) the x patch IS the depthwise input: TMA writes it straight
// into hpatch.
//
// Warp roles (512 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w3 store / residual engine,
// w4-7 convert + epilogue (TMEM lane quarter = warp % 4), w8-15 depthwise.
// The MMA thread runs expand(g) one chunk ahead of project(g-1), so convert(g+1) overlaps dw(g).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace ssdsb {
namespace {

constexpr int MB_NT = 512;
constexpr int MB_DW_THREADS = 256;
constexpr int MB_STAGING_BYTES = BLOCK_M * 128;
constexpr int MB_MAX_BYTES = 232448;

struct MbParams {
  int H, W, Ho, Wo;
  int Cin, hid, Cout;
  int S, has_expand, has_res;
  int BW, BH, PW, PH, PP, PM;
  int tiles_w, tiles_h, num_tiles;
  int nch;                 // hidden chunks per tile (hid / HC)
  int xkb, blk_x;          // k-blocks of the expand GEMM: Cin / blk_x channels each (32 | 64)
  int xkb_bytes, x_bytes;  // one k-block of the x patch / the whole patch
  int n_xbuf;
  int block_n;             // project accumulator columns (64 | 128 | 256)
  int R, lag;              // staging slots, TMA stores kept in flight
  int nseg, rs;            // depthwise: row segments per tile, output rows per segment
  int relu_e, relu_d, relu_p;
  int we_bytes, wp_bytes, hp_bytes, a_bytes;
  int off_x, off_we, off_wp, off_hp, off_a, off_stage, off_bar;
  const float* b_exp;
  const float* b_dw;
  const float* b_proj;
  const uint2* w_dw;       // [9][hid / 4] bf16 quads
};

__device__ __forceinline__ uint64_t make_smem_desc_rt(uint32_t smem_addr, int row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((8u * (uint32_t)row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 o;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o.x), "=r"(o.y) : "r"(addr) : "memory");
  return o;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mb_bf4_to_f(const uint2 v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}

// byte offset of channel quad `qd` (4 bf16 = 8 bytes) of row `p` in a K-major tile with HC*2-byte rows,
// 128B- (HC = 64) or 64B- (HC = 32) swizzled: the layout TMA writes and UMMA reads
template <int HC>
__device__ __forceinline__ uint32_t swz_off(int p, int qd) {
  const int sw = (HC == 64) ? (p & 7) : ((p >> 1) & 3);
  return (uint32_t)(p * (HC * 2) + ((((qd >> 1) ^ sw)) << 4) + ((qd & 1) << 3));
}

// mbarrier wait with a watchdog: a protocol bug must end in a trap with the barrier's id, not in a hung GPU
// (ids: 1 we_empty 2 hp_empty 3 wp_empty 4 a_full 5 wp_full 6 we_full 7 eacc_empty 8 slot_full 9 pacc_full
//  10 eacc_full 11 hp_full 12 a_empty 13 x_empty 14 pacc_empty 15 x_full)
__device__ __noinline__ void mb_wait_fail(int id, uint32_t parity) {
  printf("mbconv_kernel: block %d thread %d timed out on barrier %d (parity %u)\n", (int)blockIdx.x,
         (int)threadIdx.x, id, parity);
  __trap();
}
__device__ __forceinline__ void mbw(uint64_t* bar, uint32_t parity, int id) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) mb_wait_fail(id, parity);
}

struct MbTile {
  int w0, h0, n;
};
__device__ __forceinline__ MbTile mb_tile(const MbParams& p, int tile) {
  MbTile t;
  t.w0 = (tile % p.tiles_w) * p.BW;
  t.h0 = ((tile / p.tiles_w) % p.tiles_h) * p.BH;
  t.n = tile / (p.tiles_w * p.tiles_h);
  return t;
}

template <int HC>
__global__ void __launch_bounds__(MB_NT, 1)
mbconv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
              const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmY,
              const __grid_constant__ CUtensorMap tmR, const __grid_constant__ MbParams p) {
  constexpr int QD = HC / 4;            // channel quads per chunk
  constexpr int RB = HC * 2;            // bytes per row of hpatch / of the project A operand
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* xbuf = smem + p.off_x;
  unsigned char* webuf = smem + p.off_we;
  unsigned char* wpbuf = smem + p.off_wp;
  unsigned char* hpbuf = smem + p.off_hp;
  unsigned char* abuf = smem + p.off_a;
  unsigned char* staging = smem + p.off_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* x_full = bars;            // [2]
  uint64_t* x_empty = bars + 2;       // [2]
  uint64_t* we_full = bars + 4;       // [2]
  uint64_t* we_empty = bars + 6;      // [2]
  uint64_t* wp_full = bars + 8;       // [2]
  uint64_t* wp_empty = bars + 10;     // [2]
  uint64_t* eacc_full = bars + 12;    // [2]
  uint64_t* eacc_empty = bars + 14;   // [2]
  uint64_t* hp_full = bars + 16;      // [2]
  uint64_t* hp_empty = bars + 18;     // [2]
  uint64_t* a_full = bars + 20;       // [2]
  uint64_t* a_empty = bars + 22;      // [2]
  uint64_t* pacc_full = bars + 24;    // [1]
  uint64_t* pacc_empty = bars + 25;   // [1]
  uint64_t* slot_ready = bars + 26;   // [4]
  uint64_t* slot_full = bars + 30;    // [4]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 34);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWe);
    tma_prefetch_desc(&tmWp);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&we_full[i], 1);
      mbar_init(&we_empty[i], 1);
      mbar_init(&wp_full[i], 1);
      mbar_init(&wp_empty[i], 1);
      mbar_init(&eacc_full[i], 1);
      mbar_init(&eacc_empty[i], 4);                     // the 4 convert warps
      mbar_init(&hp_full[i], p.has_expand ? 4u : 1u);   // convert warps, or the TMA transaction
      mbar_init(&hp_empty[i], 8);                       // the 8 depthwise warps
      mbar_init(&a_full[i], 8);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(pacc_full, 1);
    mbar_init(pacc_empty, 8);                           // 4 epilogue warps x 2 column halves
    for (int i = 0; i < 4; ++i) {
      mbar_init(&slot_ready[i], 1);
      mbar_init(&slot_full[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // TMEM columns: [0, block_n) project accumulator; then 2 x PM expand accumulators of HC columns
  const uint32_t eacc_col0 = (uint32_t)p.block_n;
  const int nch = p.nch;
  const int nchunks_out = (p.Cout + 63) >> 6;
  const int rowb_x = p.blk_x * 2;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int g = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const MbTile t = mb_tile(p, tile);
        const int iw0 = t.w0 * p.S - 1, ih0 = t.h0 * p.S - 1;
        if (p.has_expand) {
          const int xb = tcount % p.n_xbuf;
          mbw(&x_empty[xb], (uint32_t)(((tcount / p.n_xbuf) & 1) ^ 1), 13);
          mbar_expect_tx(&x_full[xb], (uint32_t)(p.xkb * p.PP * rowb_x));
          for (int kb = 0; kb < p.xkb; ++kb)
            tma_load_4d(xbuf + xb * p.x_bytes + kb * p.xkb_bytes, &tmX, &x_full[xb], kb * p.blk_x, iw0, ih0, t.n);
        }
        for (int h = 0; h < nch; ++h, ++g) {
          const int s = g & 1;
          const uint32_t ph = (uint32_t)((g >> 1) & 1);
          if (p.has_expand) {
            mbw(&we_empty[s], ph ^ 1, 1);
            mbar_expect_tx(&we_full[s], (uint32_t)(p.xkb * HC * rowb_x));
            for (int kb = 0; kb < p.xkb; ++kb)
              tma_load_2d(webuf + s * p.we_bytes + kb * HC * rowb_x, &tmWe, &we_full[s], kb * p.blk_x, h * HC);
          } else {
            mbw(&hp_empty[s], ph ^ 1, 2);
            mbar_expect_tx(&hp_full[s], (uint32_t)(p.PP * RB));
            tma_load_4d(hpbuf + s * p.hp_bytes, &tmX, &hp_full[s], h * HC, iw0, ih0, t.n);
          }
          mbw(&wp_empty[s], ph ^ 1, 3);
          mbar_expect_tx(&wp_full[s], (uint32_t)(p.block_n * RB));
          tma_load_2d(wpbuf + s * p.wp_bytes, &tmWp, &wp_full[s], h * HC, 0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_e = make_idesc(BLOCK_M, HC);
      const uint32_t idesc_p = make_idesc(BLOCK_M, p.block_n);
      auto project = [&](int gp) {
        const int s = gp & 1;
        const uint32_t ph = (uint32_t)((gp >> 1) & 1);
        const int hh = gp % nch, tc = gp / nch;
        mbw(&a_full[s], ph, 4);
        mbw(&wp_full[s], ph, 5);
        if (hh == 0) mbw(pacc_empty, (uint32_t)((tc & 1) ^ 1), 14);
        tcgen05_fence_after();
        const uint64_t a_desc = make_smem_desc_rt(smem_u32(abuf + s * p.a_bytes), RB);
        const uint64_t b_desc = make_smem_desc_rt(smem_u32(wpbuf + s * p.wp_bytes), RB);
#pragma unroll
        for (int k = 0; k < HC / UMMA_K; ++k)
          umma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc_p,
                    (hh > 0 || k > 0) ? 1u : 0u);
        tcgen05_commit(&a_empty[s]);
        tcgen05_commit(&wp_empty[s]);
        if (hh == nch - 1) tcgen05_commit(pacc_full);
      };
      int g = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const int xb = tcount % p.n_xbuf;
        for (int h = 0; h < nch; ++h, ++g) {
          if (p.has_expand) {
            const int s = g & 1;
            const uint32_t ph = (uint32_t)((g >> 1) & 1);
            if (h == 0) mbw(&x_full[xb], (uint32_t)((tcount / p.n_xbuf) & 1), 15);
            mbw(&we_full[s], ph, 6);
            mbw(&eacc_empty[s], ph ^ 1, 7);
            tcgen05_fence_after();
            const uint32_t xa = smem_u32(xbuf + xb * p.x_bytes);
            const uint32_t wa = smem_u32(webuf + s * p.we_bytes);
            for (int m = 0; m < p.PM; ++m) {
              const uint32_t d_tmem = tmem_base + eacc_col0 + (uint32_t)((s * p.PM + m) * HC);
              for (int kb = 0; kb < p.xkb; ++kb) {
                const uint64_t a_desc =
                    make_smem_desc_rt(xa + (uint32_t)(kb * p.xkb_bytes + m * BLOCK_M * rowb_x), rowb_x);
                const uint64_t b_desc = make_smem_desc_rt(wa + (uint32_t)(kb * HC * rowb_x), rowb_x);
                const int ksteps = p.blk_x / UMMA_K;
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc_e,
                            (kb > 0 || k > 0) ? 1u : 0u);
              }
            }
            tcgen05_commit(&we_empty[s]);
            tcgen05_commit(&eacc_full[s]);
            if (h == nch - 1) tcgen05_commit(&x_empty[xb]);
          }
          if (g > 0) project(g - 1);
        }
      }
      if (g > 0) project(g - 1);
    }
  } else if (warp == 3) {
    // ================= store / residual engine =================
    if (lane == 0) {
      const int R = p.R, lag = p.lag;
      const uint32_t box_bytes = (uint32_t)(p.BW * p.BH) * 128u;
      int a_tile = blockIdx.x, a_chunk = 0, a_slot = 0;
      auto arm_next = [&]() {
        if (a_tile >= p.num_tiles) return;
        const MbTile ta = mb_tile(p, a_tile);
        const int slot = a_slot;
        if (++a_slot == R) a_slot = 0;
        if (p.has_res) {
          mbar_expect_tx(&slot_ready[slot], box_bytes);
          tma_load_4d(staging + slot * MB_STAGING_BYTES, &tmR, &slot_ready[slot], a_chunk * 64, ta.w0, ta.h0, ta.n);
        } else {
          mbar_arrive(&slot_ready[slot]);
        }
        if (++a_chunk == nchunks_out) {
          a_chunk = 0;
          a_tile += gridDim.x;
        }
      };
      for (int i = 0; i < R; ++i) arm_next();
      int gs = 0;
      SlotRing sring = {0, 0u};
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const MbTile t = mb_tile(p, tile);
        for (int c = 0; c < nchunks_out; ++c, ++gs) {
          mbw(&slot_full[sring.slot], sring.phase, 8);
          tma_store_4d(&tmY, staging + sring.slot * MB_STAGING_BYTES, c * 64, t.w0, t.h0, t.n);
          tma_store_commit();
          sring.advance(R);
          if (gs >= lag) {
            tma_store_wait_read_n(lag);
            arm_next();
          }
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== convert (per chunk) + epilogue (per tile) ===============================
    const int q = warp & 3;
    const int r_tile = q * 32 + lane;              // row of an M tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t staging_addr = smem_u32(staging);
    const float hi_e = (p.relu_e == 2) ? 6.0f : __int_as_float(0x7f800000);
    const float lo_e = p.relu_e ? 0.0f : -__int_as_float(0x7f800000);
    SlotRing ring = {0, 0u};
    auto epilogue = [&](int tc) {
      mbw(pacc_full, (uint32_t)(tc & 1), 9);
      tcgen05_fence_after();
      for (int c = 0; c < nchunks_out; ++c) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          SlotRing rr = ring;
          staged_epilogue_item(t_lane + (uint32_t)(c * 64), 1, p.b_proj + c * 64 + half * 32, staging_addr,
                               MB_STAGING_BYTES, p.R, rr, slot_ready, slot_full,
                               c == nchunks_out - 1 ? pacc_empty : nullptr, p.has_res != 0, p.relu_p, r_tile, half,
                               lane, p.Cout - c * 64 - half * 32);
          if (half == 1) ring = rr;
        }
      }
    };
    // patch pixel of each of this thread's rows (fixed for the whole kernel)
    int pr_h[3], pr_w[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int r = m * BLOCK_M + r_tile;
      pr_h[m] = r / p.PW;
      pr_w[m] = r - pr_h[m] * p.PW;
    }
    int g = 0, tcount = 0;
    bool pending = false;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      if (!p.has_expand) {
        epilogue(tcount);
        continue;
      }
      const MbTile t = mb_tile(p, tile);
      const int iw0 = t.w0 * p.S - 1, ih0 = t.h0 * p.S - 1;
      bool inimg[3];
#pragma unroll
      for (int m = 0; m < 3; ++m)
        inimg[m] = (unsigned)(ih0 + pr_h[m]) < (unsigned)p.H && (unsigned)(iw0 + pr_w[m]) < (unsigned)p.W;
      for (int h = 0; h < nch; ++h, ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        mbw(&eacc_full[s], ph, 10);
        mbw(&hp_empty[s], ph ^ 1, 2);
        tcgen05_fence_after();
        const uint32_t hp_addr = smem_u32(hpbuf + s * p.hp_bytes);
        const float* bias = p.b_exp + h * HC;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          if (m >= p.PM || m * BLOCK_M + q * 32 >= p.PP) break;        // warp-uniform
          const int r = m * BLOCK_M + r_tile;
          const bool keep = inimg[m];
#pragma unroll
          for (int half = 0; half < HC / 32; ++half) {
            uint32_t v[32];
            tmem_ld32(t_lane + eacc_col0 + (uint32_t)((s * p.PM + m) * HC + half * 32), v);
            float4 bv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) bv[e] = __ldg(reinterpret_cast<const float4*>(bias + half * 32) + e);
            tmem_ld_wait_dep(v);
            if (r < p.PP) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float4 b4 = bv[gq * 2 + e];
                  f[e * 4 + 0] = __uint_as_float(v[gq * 8 + e * 4 + 0]) + b4.x;
                  f[e * 4 + 1] = __uint_as_float(v[gq * 8 + e * 4 + 1]) + b4.y;
                  f[e * 4 + 2] = __uint_as_float(v[gq * 8 + e * 4 + 2]) + b4.z;
                  f[e * 4 + 3] = __uint_as_float(v[gq * 8 + e * 4 + 3]) + b4.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = keep ? fminf(fmaxf(f[e], lo_e), hi_e) : 0.0f;
                uint4 o;
                o.x = pack_bf16(f[0], f[1]);
                o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]);
                o.w = pack_bf16(f[6], f[7]);
                // 16-byte chunk (half * 4 + gq) of row r = channel quads 2*(half*4+gq), +1
                sts_u4(hp_addr + swz_off<HC>(r, 2 * (half * 4 + gq)), o);
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&eacc_empty[s]);
          mbar_arrive(&hp_full[s]);
        }
        // the previous tile's epilogue runs after this tile's first chunk has been handed to the depthwise
        // warps: its project MMAs only complete after dw of the previous tile's last chunk
        if (pending && h == 0) {
          epilogue(tcount - 1);
          pending = false;
        }
      }
      pending = true;
    }
    if (pending) epilogue(tcount - 1);
  } else if (warp >= 8) {
    // =============================== depthwise ===============================
    const int tid = threadIdx.x - 256;
    const float hi_d = (p.relu_d == 2) ? 6.0f : __int_as_float(0x7f800000);
    const float lo_d = p.relu_d ? 0.0f : -__int_as_float(0x7f800000);
    const int n_items = p.nseg * p.BW * QD;
    const int hq = p.hid >> 2;
    int g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int h = 0; h < nch; ++h, ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        mbw(&hp_full[s], ph, 11);
        mbw(&a_empty[s], ph ^ 1, 12);
        const uint32_t hp_addr = smem_u32(hpbuf + s * p.hp_bytes);
        const uint32_t a_addr = smem_u32(abuf + s * p.a_bytes);
        for (int item = tid; item < n_items; item += MB_DW_THREADS) {
          const int qd = item & (QD - 1);
          const int tt = item / QD;
          const int seg = tt / p.BW;
          const int bw = tt - seg * p.BW;
          const int r0 = seg * p.rs;
          const int rows = min(p.rs, p.BH - r0);
          if (rows <= 0) continue;
          float wf[9][4], b[4];
          {
            const uint2* wp = p.w_dw + h * QD + qd;
#pragma unroll
            for (int k = 0; k < 9; ++k) mb_bf4_to_f(__ldg(wp + (size_t)k * hq), wf[k]);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b_dw + h * HC) + qd);
            b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
          }
          auto load_row = [&](int prow, int pcol, float (&f)[3][4]) {
            const int pix = prow * p.PW + pcol;
            mb_bf4_to_f(lds_u2(hp_addr + swz_off<HC>(pix, qd)), f[0]);
            mb_bf4_to_f(lds_u2(hp_addr + swz_off<HC>(pix + 1, qd)), f[1]);
            mb_bf4_to_f(lds_u2(hp_addr + swz_off<HC>(pix + 2, qd)), f[2]);
          };
          auto fma_row = [&](float (&a)[4], const float (&f)[3][4], int dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] = fmaf(f[dx][e], wf[dy * 3 + dx][e], a[e]);
          };
          auto emit = [&](int oh, const float (&a)[4]) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fminf(fmaxf(a[e], lo_d), hi_d);
            sts_u2(a_addr + swz_off<HC>(oh * p.BW + bw, qd), pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
          };
          if (p.S == 1) {
            // patch row r0 + i feeds outputs r0 + i (dy 0), r0 + i - 1 (dy 1), r0 + i - 2 (dy 2)
            float acc[3][4];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[a][e] = b[e];
            const int n_in = rows + 2;
            for (int i0 = 0; i0 < n_in; i0 += 3) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const int i = i0 + j;
                if (i < n_in) {
                  float f[3][4];
                  load_row(r0 + i, bw, f);
                  fma_row(acc[j], f, 0);
                  fma_row(acc[(j + 2) % 3], f, 1);
                  fma_row(acc[(j + 1) % 3], f, 2);
                  if (i >= 2) emit(r0 + i - 2, acc[(j + 1) % 3]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[(j + 1) % 3][e] = b[e];
              }
            }
          } else {
            float a[4] = {b[0], b[1], b[2], b[3]};
            {
              float f[3][4];
              load_row(2 * r0, 2 * bw, f);
              fma_row(a, f, 0);
            }
            for (int o = 0; o < rows; ++o) {
              float f1[3][4], f2[3][4];
              load_row(2 * (r0 + o) + 1, 2 * bw, f1);
              load_row(2 * (r0 + o) + 2, 2 * bw, f2);
              fma_row(a, f1, 1);
              fma_row(a, f2, 2);
              emit(r0 + o, a);
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] = b[e];
              fma_row(a, f2, 0);
            }
          }
        }
        fence_proxy_async();              // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[s]);
          mbar_arrive(&hp_empty[s]);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct MbGeom {
  int BW, BH, PW, PH, PP, PM, nseg, rs;
  double cost;
};

// output tile BW x BH (<= 128 pixels) whose input patch fits PM <= 3 M tiles; the cost model only has to rank
// candidates: depthwise rounds of 256 threads x rows streamed, convert work per M tile, a fixed per-tile part
MbGeom pick_geometry(int Ho, int Wo, int S, int HC, int nch, bool has_expand, int tmem_cols_left, int sms, int N) {
  MbGeom best = {};
  best.cost = -1.0;
  const int qd = HC / 4;
  for (int BW = 1; BW <= Wo && BW <= 64; ++BW) {
    for (int BH = 1; BH <= Ho && BW * BH <= BLOCK_M; ++BH) {
      const int PW = (BW - 1) * S + 3, PH = (BH - 1) * S + 3;
      const int PP = PW * PH;
      if (PP > 3 * BLOCK_M || PW > 256 || PH > 256) continue;
      const int PM = (PP + BLOCK_M - 1) / BLOCK_M;
      if (has_expand && 2 * PM * HC > tmem_cols_left) continue;
      double dw_best = 1e30;
      int nseg_best = 1;
      for (int nseg = 1; nseg <= 4 && nseg <= BH; ++nseg) {
        const int rs = (BH + nseg - 1) / nseg;
        const int rounds = (nseg * BW * qd + MB_DW_THREADS - 1) / MB_DW_THREADS;
        const double c = rounds * (rs * (S == 1 ? 44.0 : 56.0) + 70.0);
        if (c < dw_best) {
          dw_best = c;
          nseg_best = nseg;
        }
      }
      const double conv = has_expand ? PM * (HC / 32) * 110.0 : 0.0;
      const double tile_cost = nch * (dw_best + 0.5 * conv + 40.0) + 500.0;
      const long long tiles = (long long)N * ((Wo + BW - 1) / BW) * ((Ho + BH - 1) / BH);
      const double waves = (double)((tiles + sms - 1) / sms);
      const double cost = waves * tile_cost;
      if (best.cost < 0 || cost < best.cost - 1e-9) {
        best.BW = BW; best.BH = BH; best.PW = PW; best.PH = PH; best.PP = PP; best.PM = PM;
        best.nseg = nseg_best;
        best.rs = (BH + nseg_best - 1) / nseg_best;
        best.cost = cost;
      }
    }
  }
  return best;
}

CUresult encode_tm(EncodeTiledFn encode, CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims,
                   const cuuint64_t* strides, const cuuint32_t* box, int inner_bytes) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle swz = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

thread_local int g_mb_last[12] = {0};

int pick_block_n_mb(int cout) { return cout <= 64 ? 64 : (cout <= 128 ? 128 : 256); }

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_mbconv_last_launch(int* out12) {
  SSDSB_REQUIRE(out12, "mbconv_last_launch: NULL argument");
  for (int i = 0; i < 12; ++i) out12[i] = g_mb_last[i];
  return SSDSB_OK;
}

extern "C" int ssdsb_mbconv_bf16(const ssdsb_mbconv_desc* d, const void* x, const void* w_exp, const float* b_exp,
                                 const void* w_dw, const float* b_dw, const void* w_proj, const float* b_proj,
                                 void* y, void* stream) {
  SSDSB_REQUIRE(d && x && w_dw && b_dw && w_proj && b_proj && y, "mbconv: NULL argument");
  const bool has_expand = w_exp != nullptr;
  SSDSB_REQUIRE(!has_expand || b_exp, "mbconv: expand weights without bias");
  SSDSB_REQUIRE(d->N >= 1 && d->H >= 1 && d->W >= 1, "mbconv: bad shape");
  SSDSB_REQUIRE(d->stride == 1 || d->stride == 2, "mbconv: stride=%d", d->stride);
  SSDSB_REQUIRE(d->Cin % 32 == 0 && d->hid % 32 == 0 && d->Cout % 32 == 0 && d->Cin >= 32 && d->hid >= 32 &&
                    d->Cout >= 32,
                "mbconv: channel counts must be multiples of 32 (Cin=%d hid=%d Cout=%d)", d->Cin, d->hid, d->Cout);
  SSDSB_REQUIRE(has_expand || d->hid == d->Cin, "mbconv: without an expand layer hid must equal Cin");
  SSDSB_REQUIRE(!d->residual || (d->stride == 1 && d->Cin == d->Cout), "mbconv: residual needs stride 1, Cin == Cout");
  SSDSB_REQUIRE(d->relu_expand >= 0 && d->relu_expand <= 2 && d->relu_dw >= 0 && d->relu_dw <= 2 &&
                    d->relu_project >= 0 && d->relu_project <= 2,
                "mbconv: activation codes are 0 none, 1 ReLU, 2 ReLU6");
  SSDSB_REQUIRE(d->w_exp_rows >= (has_expand ? d->hid : 0) && d->w_proj_rows >= d->Cout, "mbconv: weight rows");
  SSDSB_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w_dw | (uintptr_t)w_proj | (uintptr_t)w_exp |
                  (uintptr_t)b_dw | (uintptr_t)b_proj | (uintptr_t)b_exp) & 15) == 0,
                "mbconv: pointers must be 16-byte aligned");
  const int out_cs = d->out_cstride ? d->out_cstride : d->Cout;
  SSDSB_REQUIRE(out_cs >= d->Cout && out_cs % 8 == 0, "mbconv: bad output channel stride");
  if (d->Cout > 256)
    return fail(SSDSB_ERR_UNSUPPORTED, "mbconv: Cout=%d > 256 needs two project accumulators (use the separate launches)",
                d->Cout);
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(SSDSB_ERR_CUDA, "mbconv: cuTensorMapEncodeTiled entry point not found");
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SSDSB_CUDA(cudaGetDevice(&dev));
    SSDSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int S = d->stride;
  const int Ho = (d->H - 1) / S + 1, Wo = (d->W - 1) / S + 1;
  const int block_n = pick_block_n_mb(d->Cout);
  const int nchunks_out = (d->Cout + 63) / 64;

  MbParams p = {};
  p.H = d->H; p.W = d->W; p.Ho = Ho; p.Wo = Wo;
  p.Cin = d->Cin; p.hid = d->hid; p.Cout = d->Cout;
  p.S = S; p.has_expand = has_expand ? 1 : 0; p.has_res = d->residual ? 1 : 0;
  p.block_n = block_n;
  p.relu_e = d->relu_expand; p.relu_d = d->relu_dw; p.relu_p = d->relu_project;
  p.b_exp = b_exp; p.b_dw = b_dw; p.b_proj = b_proj;
  p.w_dw = reinterpret_cast<const uint2*>(w_dw);
  p.blk_x = has_expand ? (d->Cin % 64 == 0 ? 64 : 32) : 0;
  p.xkb = has_expand ? d->Cin / p.blk_x : 0;

  // configuration search: hidden chunk 64 (if it divides hid) before 32; two x buffers before one; 4 staging slots
  // before the minimum — first one that fits shared memory and TMEM
  int HC = 0, smem_total = 0;
  const char* force_hc = getenv("SSDSB_MB_HC");
  for (int hc_try : {64, 32}) {
    if (d->hid % hc_try) continue;
    if (force_hc && atoi(force_hc) != hc_try) continue;
    const int nch = d->hid / hc_try;
    MbGeom gm = pick_geometry(Ho, Wo, S, hc_try, nch, has_expand, 512 - block_n, sms, d->N);
    if (gm.cost < 0) continue;
    if (const char* e = getenv("SSDSB_MB_TILE")) {              // experiment knob: "BWxBH"
      int bw = 0, bh = 0;
      if (sscanf(e, "%dx%d", &bw, &bh) == 2 && bw >= 1 && bh >= 1 && bw * bh <= BLOCK_M && bw <= Wo && bh <= Ho) {
        const int PW = (bw - 1) * S + 3, PH = (bh - 1) * S + 3;
        const int PM = (PW * PH + BLOCK_M - 1) / BLOCK_M;
        if (PM <= 3 && (!has_expand || 2 * PM * hc_try <= 512 - block_n)) {
          gm.BW = bw; gm.BH = bh; gm.PW = PW; gm.PH = PH; gm.PP = PW * PH; gm.PM = PM;
          gm.nseg = 1; gm.rs = bh;
        }
      }
    }
    const int pp_pad = (gm.PP + 15) / 16 * 16;
    const int rowb_x = p.blk_x * 2;
    const int xkb_bytes = pp_pad * rowb_x;
    const int x_bytes = p.xkb * xkb_bytes;
    const int we_bytes = has_expand ? p.xkb * hc_try * rowb_x : 0;
    const int wp_bytes = block_n * hc_try * 2;
    const int hp_bytes = pp_pad * hc_try * 2;
    const int a_bytes = BLOCK_M * hc_try * 2;
    bool done = false;
    for (int n_xbuf = has_expand ? 2 : 0; n_xbuf >= (has_expand ? 1 : 0) && !done; --n_xbuf) {
      for (int R : {4, 2}) {
        const int lag = R >= 4 ? (nchunks_out <= 2 ? 2 : 1) : 0;
        // the x M tiles of the last k-block are read up to PM*128 rows: what follows must be inside the window
        const int total = n_xbuf * x_bytes + 2 * we_bytes + 2 * wp_bytes + 2 * hp_bytes + 2 * a_bytes +
                          R * MB_STAGING_BYTES + 512 + 1024;
        if (total > MB_MAX_BYTES) continue;
        HC = hc_try;
        p.nch = nch;
        p.BW = gm.BW; p.BH = gm.BH; p.PW = gm.PW; p.PH = gm.PH; p.PP = gm.PP; p.PM = gm.PM;
        p.nseg = gm.nseg; p.rs = gm.rs;
        p.xkb_bytes = xkb_bytes; p.x_bytes = x_bytes; p.n_xbuf = n_xbuf ? n_xbuf : 1;
        p.we_bytes = we_bytes; p.wp_bytes = wp_bytes; p.hp_bytes = hp_bytes; p.a_bytes = a_bytes;
        p.R = R; p.lag = lag;
        p.off_x = 0;
        p.off_we = p.off_x + n_xbuf * x_bytes;
        p.off_wp = p.off_we + 2 * we_bytes;
        p.off_hp = p.off_wp + 2 * wp_bytes;
        p.off_a = p.off_hp + 2 * hp_bytes;
        p.off_stage = p.off_a + 2 * a_bytes;
        p.off_bar = p.off_stage + R * MB_STAGING_BYTES;
        smem_total = total;
        done = true;
        break;
      }
    }
    if (done) break;
  }
  if (!HC)
    return fail(SSDSB_ERR_UNSUPPORTED, "mbconv: no configuration fits shared memory / TMEM (Cin=%d hid=%d Cout=%d)",
                d->Cin, d->hid, d->Cout);
  p.tiles_w = (Wo + p.BW - 1) / p.BW;
  p.tiles_h = (Ho + p.BH - 1) / p.BH;
  p.num_tiles = p.tiles_w * p.tiles_h * d->N;

  alignas(64) CUtensorMap tmX, tmWe, tmWp, tmY, tmR;
  {
    // x as [N][H][W][Cin]: the patch box; without an expand layer it carries HC channels per load into hpatch
    const int inner = has_expand ? p.blk_x : HC;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->W * d->Cin * 2, (cuuint64_t)d->H * d->W * d->Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1};
    CUresult r = encode_tm(encode, &tmX, 4, x, dims, strides, box, inner * 2);
    if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "mbconv: x tensor map failed (CUresult %d)", (int)r);
  }
  if (has_expand) {
    cuuint64_t dims[2] = {(cuuint64_t)d->Cin, (cuuint64_t)d->w_exp_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d->Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.blk_x, (cuuint32_t)HC};
    CUresult r = encode_tm(encode, &tmWe, 2, w_exp, dims, strides, box, p.blk_x * 2);
    if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "mbconv: expand weight tensor map failed (CUresult %d)", (int)r);
  } else {
    tmWe = tmX;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->hid, (cuuint64_t)d->w_proj_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d->hid * 2};
    cuuint32_t box[2] = {(cuuint32_t)HC, (cuuint32_t)block_n};
    CUresult r = encode_tm(encode, &tmWp, 2, w_proj, dims, strides, box, HC * 2);
    if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "mbconv: project weight tensor map failed (CUresult %d)", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)out_cs * 2, (cuuint64_t)Wo * out_cs * 2, (cuuint64_t)Ho * Wo * out_cs * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    CUresult r = encode_tm(encode, &tmY, 4, y, dims, strides, box, 128);
    if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "mbconv: output tensor map failed (CUresult %d)", (int)r);
  }
  if (p.has_res) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->W * d->Cin * 2, (cuuint64_t)d->H * d->W * d->Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    CUresult r = encode_tm(encode, &tmR, 4, x, dims, strides, box, 128);
    if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "mbconv: residual tensor map failed (CUresult %d)", (int)r);
  } else {
    tmR = tmY;
  }

  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const int info[12] = {HC, p.BW, p.BH, p.PM, p.nch, p.n_xbuf, p.R, p.nseg, p.rs, grid, smem_total, block_n};
  for (int i = 0; i < 12; ++i) g_mb_last[i] = info[i];
  cudaStream_t st = (cudaStream_t)stream;
  if (HC == 64) {
    static bool configured = false;
    if (!configured) {
      SSDSB_CUDA(cudaFuncSetAttribute(mbconv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_MAX_BYTES));
      configured = true;
    }
    mbconv_kernel<64><<<grid, MB_NT, smem_total, st>>>(tmX, tmWe, tmWp, tmY, tmR, p);
  } else {
    static bool configured = false;
    if (!configured) {
      SSDSB_CUDA(cudaFuncSetAttribute(mbconv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_MAX_BYTES));
      configured = true;
    }
    mbconv_kernel<32><<<grid, MB_NT, smem_total, st>>>(tmX, tmWe, tmWp, tmY, tmR, p);
  }
  SSDSB_LAUNCH_CHECK("mbconv_kernel");
  return SSDSB_OK;
}
