// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces, for the SSD conv stack (reference ssds/modeling/ssds/ssd.py:42-74 over
// ssds/modeling/nets/resnet.py:41-56 and layers/basic_layers.py:27-57), every
// nn.Conv2d -> BatchNorm2d(eval) -> [+ residual] -> ReLU chain and the 3x3 multibox heads
// (ssd.py:100-103, + sigmoid in eval, ssd.py:72-73) by ONE kernel launch per convolution.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * Wt[cout, tap, cin]
//
//   * activations: NHWC bf16.  An M-tile is a (BW x BH x BN) patch of 128 output pixels; the A
//     operand of one (tap, 64-channel) K-block is ONE 4-D TMA box load at the shifted coordinate
//     (c, w*s+kw-p, h*s+kh-p, n): TMA zero-fills the halo (conv padding) and the ragged edge, and
//     its traversal stride implements stride-2 convolutions.  No im2col buffer exists anywhere.
//   * weights: BN-folded bf16 [Cout_pad][tap][Cin] (K-major) -> 2-D TMA box (64, BLOCK_N).
//   * both land in 128B-swizzled shared memory, STAGES deep; one elected thread issues
//     tcgen05.mma (M=128, N=BLOCK_N, K=16) with fp32 accumulators in TMEM, double buffered so
//     the epilogue of tile i overlaps the MMAs of tile i+1 (persistent CTAs, 1 per SM).
//   * epilogue warps: tcgen05.ld -> +bias (folded BN / conv bias) -> +residual -> ReLU -> bf16
//     NHWC, or for the multibox head -> fp32 NCHW split into loc / sigmoid(conf), the exact
//     tensors `model(x)` returns in the reference.
//   * 7x7/s2 stem (Cin=3): the image is pre-packed 2x2 space-to-depth into NHWC16 (layout.cu), on
//     which the stem is a 4x4/s1 convolution; its K-block is one tap x 16 channels (32-byte rows,
//     SWIZZLE_32B, one K=16 MMA per block) — same kernel, BLOCK_K = 16.
//
//   * NHWC epilogue: the 128x64 bf16 sub-tile is staged in 128B-swizzled shared memory and written
//     with ONE 4-D TMA store (full 128-byte lines, ragged edges clipped by TMA); the residual
//     sub-tile is TMA-loaded into the same staging buffer two chunks ahead and added in place.
//
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-11 epilogue
// (two warps per TMEM lane quarter, each owning one 32-column half of every 64-column chunk: a lone
// warp per scheduler cannot hide its own instruction latency, which made the epilogue the bottleneck).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_params.h"
#include "umma.cuh"

namespace ssdsb {
namespace {


// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
constexpr int MAX_STAGES = 16;
constexpr int MAX_STAGING = 8;
constexpr int STAGING_BYTES = BLOCK_M * 128;   // 128 rows x 64 bf16

template <int BLOCK_N, int BLOCK_K>
struct ConvSmem {
  static constexpr int ROW_BYTES = BLOCK_K * 2;
  static constexpr int A_STAGE_BYTES = BLOCK_M * ROW_BYTES;
  static constexpr int B_STAGE_BYTES = BLOCK_N * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int MAX_BYTES = 232448;  // 227 KiB opt-in limit per CTA
  static_assert(STAGE_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned");
  // one pipeline stage = one k-block of a group: the A tiles of its `ways` M tiles + (unless the weights
  // are resident) ONE B tile, behind one full / one empty barrier
  static int stage_bytes(bool b_resident, int ways) {
    return ways * A_STAGE_BYTES + (b_resident ? 0 : B_STAGE_BYTES);
  }
  static int stages_for(int n_staging, bool b_resident, int num_k_blocks, int ways) {
    const int fixed = 1024 + BAR_BYTES + n_staging * STAGING_BYTES +
                      (b_resident ? num_k_blocks * B_STAGE_BYTES : 0);
    int s = (MAX_BYTES - fixed) / stage_bytes(b_resident, ways);
    return s > MAX_STAGES ? MAX_STAGES : s;
  }
  static int bytes(int stages, int n_staging, bool b_resident, int num_k_blocks, int ways) {
    return stages * stage_bytes(b_resident, ways) + n_staging * STAGING_BYTES +
           (b_resident ? num_k_blocks * B_STAGE_BYTES : 0) + 1024 + BAR_BYTES;
  }
};

struct TileCoord {
  int n_tile, w0, h0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvKernelParams& p, int m, int n_tile) {
  TileCoord t;
  t.n_tile = n_tile;
  const int tw = m % p.tiles_w;
  const int th = (m / p.tiles_w) % p.tiles_h;
  const int tn = m / (p.tiles_w * p.tiles_h);
  t.w0 = tw * p.BW;
  t.h0 = th * p.BH;
  t.n0 = tn * p.BN;
  return t;
}

// The MMA-issuing loop of conv_igemm_kernel for a converged warp (see the call site).
template <int BLOCK_N, int BLOCK_K, int WAYS>
__device__ __noinline__ void mma_issue_converged(const ConvKernelParams& p, uint32_t smem_addr, uint32_t bres_addr,
                                                 uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, uint64_t* b_full, uint32_t tmem_base,
                                                 int num_groups, int stage_bytes) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  constexpr int A_STAGE_BYTES = S::A_STAGE_BYTES;
  constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
  const int STAGES = p.stages;
  const uint32_t tmem_base_u = __reduce_max_sync(0xffffffffu, tmem_base);      // REDUX writes a uniform register
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  if (p.b_resident && (int)blockIdx.x < num_groups) mbar_wait(b_full, 0);
  for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
    tcgen05_fence_after();
    const uint32_t d_tmem = tmem_base_u + (uint32_t)(acc * WAYS * BLOCK_N);
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t sa = smem_addr + (uint32_t)(stage * stage_bytes);
      uint64_t a_desc[WAYS];
#pragma unroll
      for (int w = 0; w < WAYS; ++w) a_desc[w] = make_smem_desc<S::ROW_BYTES>(sa + (uint32_t)(w * A_STAGE_BYTES));
      const uint64_t b_desc = p.b_resident
          ? make_smem_desc<S::ROW_BYTES>(bres_addr + (uint32_t)(kb * S::B_STAGE_BYTES))
          : make_smem_desc<S::ROW_BYTES>(sa + (uint32_t)(WAYS * A_STAGE_BYTES));
#pragma unroll
      for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
        // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (>>4) address field;
        // consecutive MMAs go to different accumulators
#pragma unroll
        for (int w = 0; w < WAYS; ++w)
          umma_bf16_elect(d_tmem + (uint32_t)(w * BLOCK_N), a_desc[w] + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k),
                          idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      tcgen05_commit_elect(&empty_bar[stage]);    // frees the stage when these MMAs retire
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    tcgen05_commit_elect(&tmem_full[acc]);      // accumulators ready for the epilogue
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
}

template <int BLOCK_N, int BLOCK_K, int WAYS>
__global__ void __launch_bounds__(CONV_NT, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ ConvKernelParams p) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  constexpr int A_STAGE_BYTES = S::A_STAGE_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int STAGES = p.stages;
  const int stage_bytes = WAYS * A_STAGE_BYTES + (p.b_resident ? 0 : S::B_STAGE_BYTES);
  unsigned char* staging = smem + STAGES * stage_bytes;
  unsigned char* bres = staging + p.n_staging * STAGING_BYTES;     // resident weights (optional)
  uint64_t* bars = reinterpret_cast<uint64_t*>(bres + (p.b_resident ? p.num_k_blocks * S::B_STAGE_BYTES : 0));
  uint64_t* full_bar = bars;                           // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;             // [MAX_STAGES]
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;         // [2]
  uint64_t* tmem_empty = bars + 2 * MAX_STAGES + 2;    // [2]
  uint64_t* slot_ready = bars + 2 * MAX_STAGES + 4;                  // [MAX_STAGING] slot free (+ residual landed)
  uint64_t* slot_full = bars + 2 * MAX_STAGES + 4 + MAX_STAGING;     // [MAX_STAGING] 8 epilogue warps wrote it
  uint64_t* b_full = bars + 2 * MAX_STAGES + 4 + 2 * MAX_STAGING;     // [1] resident weights landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 5 + 2 * MAX_STAGING);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // WAYS M-tiles of one n-tile form a group: their MMAs are issued interleaved into WAYS separate
  // accumulators, because back-to-back MMAs on ONE accumulator retire only every ~146 cycles whatever
  // N is (tools/umma_bench.cu: N=64 146 -> 78 cycles/MMA with 4 accumulators, N=128 146 -> 116 with 2)
  // (WAYS is a template parameter: the single-thread producer / MMA loops must stay in registers.)
  // A trailing partial group simply runs "ghost" tiles whose coordinates are out of range: TMA
  // zero-fills their loads and clips their stores.
  constexpr uint32_t tmem_cols = (2 * WAYS * BLOCK_N < 32) ? 32u : (uint32_t)(2 * WAYS * BLOCK_N);
  static_assert(2 * WAYS * BLOCK_N <= 512, "TMEM has 512 columns");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmY);
    if (p.residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);  // one arrive per epilogue warp
    }
    for (int r = 0; r < MAX_STAGING; ++r) {
      mbar_init(&slot_ready[r], 1);
      mbar_init(&slot_full[r], 8);   // one arrive per epilogue warp
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_holder)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();       // everything above overlapped the previous kernel's tail; its outputs are needed from here
  pdl_trigger();

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_groups = (m_tiles + WAYS - 1) / WAYS;
  const int num_groups = m_groups * p.n_tiles;
  const int rows = p.BW * p.BH * p.BN;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.b_resident && (int)blockIdx.x < num_groups) {
        // weight-stationary: the grid is a multiple of n_tiles, so this CTA only ever sees one n_tile
        const int n_tile = (int)blockIdx.x % p.n_tiles;
        mbar_expect_tx(b_full, (uint32_t)(p.num_k_blocks * S::B_STAGE_BYTES));
        for (int kb = 0; kb < p.num_k_blocks; ++kb)
          tma_load_2d(bres + kb * S::B_STAGE_BYTES, &tmB, b_full, kb * BLOCK_K, n_tile * BLOCK_N);
      }
      const uint32_t a_bytes = (uint32_t)rows * (BLOCK_K * 2);
      for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        const int n_tile = group % p.n_tiles;
        const int m0 = (group / p.n_tiles) * WAYS;
        int cw[WAYS], ch[WAYS], cn[WAYS];          // input-space origin of each way's tile
#pragma unroll
        for (int w = 0; w < WAYS; ++w) {
          const TileCoord t = tile_coord(p, m0 + w, n_tile);
          cw[w] = t.w0 * p.stride - p.pad_w;
          ch[w] = t.h0 * p.stride - p.pad_h;
          cn[w] = t.n0;
        }
        int tap = 0, kc = 0, kh = 0, kw = 0;       // incremental (no divisions in the K loop)
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          // one stage per k-block: the A tiles of all ways + the shared B tile, one barrier pair
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage],
                         (uint32_t)WAYS * a_bytes + (p.b_resident ? 0u : (uint32_t)S::B_STAGE_BYTES));
#pragma unroll
          for (int w = 0; w < WAYS; ++w)
            tma_load_4d(sa + w * A_STAGE_BYTES, &tmA, &full_bar[stage], kc * BLOCK_K + n_tile * p.chunk_cin,
                        cw[w] + kw, ch[w] + kh, cn[w]);
          if (!p.b_resident)
            tma_load_2d(sa + WAYS * A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n_tile * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
            if (++kw == p.KW) {
              kw = 0;
              ++kh;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (p.mma_converged) {
      // the whole warp runs the loop (waits, descriptor arithmetic: all uniform); elect.sync inside the helpers picks
      // the issuing lane.  A separate NON-INLINED function: inside this 12-role kernel ptxas keeps the descriptors in
      // vector registers and moves them through R2UR for every MMA; compiled on its own the loop stays on the
      // uniform datapath (bare UTCHMMA / UTCBAR, ~3 instructions per MMA).
      mma_issue_converged<BLOCK_N, BLOCK_K, WAYS>(p, smem_u32(smem), smem_u32(bres), full_bar, empty_bar, tmem_full,
                                                  tmem_empty, b_full, tmem_base, num_groups, stage_bytes);
    } else if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.b_resident && (int)blockIdx.x < num_groups) mbar_wait(b_full, 0);
      const uint32_t bres_addr = smem_u32(bres);
      for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * WAYS * BLOCK_N);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          uint64_t a_desc[WAYS];
#pragma unroll
          for (int w = 0; w < WAYS; ++w) a_desc[w] = make_smem_desc<S::ROW_BYTES>(sa + (uint32_t)(w * A_STAGE_BYTES));
          const uint64_t b_desc = p.b_resident
              ? make_smem_desc<S::ROW_BYTES>(bres_addr + (uint32_t)(kb * S::B_STAGE_BYTES))
              : make_smem_desc<S::ROW_BYTES>(sa + (uint32_t)(WAYS * A_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (>>4) address field;
            // consecutive MMAs go to different accumulators
#pragma unroll
            for (int w = 0; w < WAYS; ++w)
              umma_bf16(d_tmem + (uint32_t)(w * BLOCK_N), a_desc[w] + (uint64_t)(2 * k),
                        b_desc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[stage]);    // frees the stage when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tcgen05_commit(&tmem_full[acc]);      // accumulators ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ================= store / residual engine (staged NHWC epilogue only) =================
    // One thread owns every TMA operation of the epilogue so that the 8 epilogue warps never block
    // on the copy engine and never barrier with each other: per 64-column chunk g (slot g % R)
    //   slot_ready[slot]  <- this thread: slot reusable (+ residual sub-tile landed via TMA)
    //   slot_full[slot]   <- the 8 epilogue warps: sub-tile written (after fence.proxy.async)
    // and this thread then issues the 4-D TMA store and re-arms the slot as soon as the store has
    // been read out of shared memory.  Residuals are thereby prefetched R-1 chunks ahead.
    if (lane == 0 && p.tma_store) {
      const int R = p.n_staging;
      const bool has_res = (p.residual != nullptr);
      int a_group = blockIdx.x, a_way = 0, a_chunk = 0, a_slot = 0;      // arming iterator
      auto arm_next = [&]() {
        if (a_group >= num_groups) return;
        const int a_m0 = (a_group / p.n_tiles) * WAYS;
        const TileCoord ta = tile_coord(p, a_m0 + a_way, a_group % p.n_tiles);
        const int slot = a_slot;
        if (++a_slot == R) a_slot = 0;
        if (has_res) {
          mbar_expect_tx(&slot_ready[slot], (uint32_t)rows * 128u);
          tma_load_4d(staging + slot * STAGING_BYTES, &tmR, &slot_ready[slot],
                      ta.n_tile * BLOCK_N + a_chunk * 64, ta.w0, ta.h0, ta.n0);
        } else {
          mbar_arrive(&slot_ready[slot]);
        }
        const int nch = (min(BLOCK_N, p.Cout - ta.n_tile * BLOCK_N) + 63) >> 6;   // last chunk may be half full
        if (++a_chunk == nch) {
          a_chunk = 0;
          if (++a_way == WAYS) {
            a_way = 0;
            a_group += gridDim.x;
          }
        }
      };
      for (int i = 0; i < R; ++i) arm_next();
      // `store_lag` stores stay in flight: waiting for the smem read of store g right after issuing it
      // caps the store rate at one 16 KiB sub-tile per read latency (~1 us under load = 2.4 TB/s chip-wide).
      // The re-arm of a slot IS the residual prefetch for the chunk R slots later, so residual sub-tiles
      // are in flight R - store_lag chunks ahead.
      const int lag = p.store_lag;
      int g = 0;
      SlotRing sring = {0, 0u};
      for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        const int m0 = (group / p.n_tiles) * WAYS;
        for (int w = 0; w < WAYS; ++w) {
          const TileCoord t = tile_coord(p, m0 + w, group % p.n_tiles);
          const int nch = (min(BLOCK_N, p.Cout - t.n_tile * BLOCK_N) + 63) >> 6;
          for (int c = 0; c < nch; ++c, ++g) {
            mbar_wait(&slot_full[sring.slot], sring.phase);
            tma_store_4d(&tmY, staging + sring.slot * STAGING_BYTES, t.n_tile * BLOCK_N + c * 64, t.w0, t.h0,
                         t.n0);
            tma_store_commit();
            sring.advance(R);
            if (g >= lag) {
              tma_store_wait_read_n(lag);   // store g-lag has been read out of smem -> its slot is free
              arm_next();                   // = chunk g-lag+R
            }
          }
        }
      }
      tma_store_wait_all();             // smem must outlive the last store
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int q = warp & 3;                   // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - 4) >> 2;         // which 32-column half of each 64-column chunk
    const int r = q * 32 + lane;              // row of the tile == TMEM lane
    int acc = 0;
    uint32_t acc_phase = 0;
    // staging ring state (TMA-store path)
    const int R = p.n_staging;
    SlotRing ring = {0, 0u};
    const bool has_res = (p.residual != nullptr);
    const bool staged = (p.mode == CONV_OUT_NHWC_BF16) && p.tma_store;
    const uint32_t staging_addr = smem_u32(staging);
    // row r -> (bn, bh, bw) in TMA box order (w fastest); constant across tiles
    const int bw = r % p.BW;
    const int bh = (r / p.BW) % p.BH;
    const int bn = r / (p.BW * p.BH);

    for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
     const int m0 = (group / p.n_tiles) * WAYS;
     mbar_wait(&tmem_full[acc], acc_phase);
     tcgen05_fence_after();
     const int n_tile = group % p.n_tiles;
     for (int way = 0; way < WAYS; ++way) {
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) +
                             (uint32_t)((acc * WAYS + way) * BLOCK_N);
      if (staged) {
        // ---------- staged path: 64-column chunks through swizzled smem + TMA store ----------
        const int nchunks = (min(BLOCK_N, p.Cout - n_tile * BLOCK_N) + 63) >> 6;
        staged_epilogue_item(t_row, nchunks, p.bias + n_tile * BLOCK_N + half * 32, staging_addr,
                             STAGING_BYTES, R, ring, slot_ready, slot_full,
                             way == WAYS - 1 ? &tmem_empty[acc] : nullptr, has_res, p.relu, r, half, lane,
                             p.Cout - n_tile * BLOCK_N - half * 32);
        continue;
      }
      const TileCoord t = tile_coord(p, m0 + way, n_tile);
      const int w = t.w0 + bw, h = t.h0 + bh, n = t.n0 + bn;
      const bool row_ok = (r < rows) && (w < p.Wo) && (h < p.Ho) && (n < p.N);
      {
#pragma unroll 1
        for (int c0 = half * 32; c0 < p.ntile_cout; c0 += 64) {
          const int col0 = t.n_tile * p.ntile_cout + c0;
          if (col0 >= p.Cout) break;            // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (p.mode == CONV_OUT_NHWC_BF16) {
            // direct path (Cout not a multiple of 64): 16-byte stores from each row owner
            const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __ldg(p.bias + col0 + j);
            if (row_ok) {
              if (has_res) {
                const uint4* rp = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.residual) + pix * p.res_cstride + col0);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  const uint4 rv = __ldg(rp + gq);
                  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    f[gq * 8 + e * 2 + 0] += __uint_as_float(rw[e] << 16);
                    f[gq * 8 + e * 2 + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                  }
                }
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                if (p.relu == 2) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = fminf(f[j], 6.0f);
                }
              }
              uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) +
                                                   pix * p.out_cstride + col0);
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                uint4 o;
                o.x = pack_bf16(f[gq * 8 + 0], f[gq * 8 + 1]);
                o.y = pack_bf16(f[gq * 8 + 2], f[gq * 8 + 3]);
                o.z = pack_bf16(f[gq * 8 + 4], f[gq * 8 + 5]);
                o.w = pack_bf16(f[gq * 8 + 6], f[gq * 8 + 7]);
                yp[gq] = o;
              }
            }
          } else if (row_ok) {
            // multibox head: channels [0, n_loc) -> loc fp32 NCHW; [n_loc, Cout) -> sigmoid -> conf.
            // Consecutive lanes are consecutive pixels, so every column store is a coalesced line.
            const size_t hw = (size_t)p.Ho * p.Wo;
            const size_t sp = (size_t)h * p.Wo + w;
            const int nvalid = min(32, p.Cout - col0);
            if (col0 >= p.n_loc && nvalid == 32) {          // whole chunk is conf (the common case)
              float* dst = reinterpret_cast<float*>(p.y2) + ((size_t)n * (p.Cout - p.n_loc) +
                                                              (size_t)(col0 - p.n_loc)) * hw + sp;
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 b4 = __ldg(bp + e);
                float x0 = __uint_as_float(v[e * 4 + 0]) + b4.x, x1 = __uint_as_float(v[e * 4 + 1]) + b4.y;
                float x2 = __uint_as_float(v[e * 4 + 2]) + b4.z, x3 = __uint_as_float(v[e * 4 + 3]) + b4.w;
                if (p.sigmoid) {
                  x0 = __fdividef(1.0f, 1.0f + __expf(-x0));
                  x1 = __fdividef(1.0f, 1.0f + __expf(-x1));
                  x2 = __fdividef(1.0f, 1.0f + __expf(-x2));
                  x3 = __fdividef(1.0f, 1.0f + __expf(-x3));
                }
                __stcs(dst, x0); dst += hw;
                __stcs(dst, x1); dst += hw;
                __stcs(dst, x2); dst += hw;
                __stcs(dst, x3); dst += hw;
              }
            } else {
              float* loc = reinterpret_cast<float*>(p.y) + (size_t)n * p.n_loc * hw + sp;
              float* conf = reinterpret_cast<float*>(p.y2) + (size_t)n * (p.Cout - p.n_loc) * hw + sp;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                if (j < nvalid) {
                  const float x = __uint_as_float(v[j]) + __ldg(p.bias + col);
                  if (col < p.n_loc) {
                    loc[(size_t)col * hw] = x;
                  } else {
                    const float sg = p.sigmoid ? __fdividef(1.0f, 1.0f + __expf(-x)) : x;
                    __stcs(conf + (size_t)(col - p.n_loc) * hw, sg);
                  }
                }
              }
            }
          }
        }
      }
     }  // ways
      if (!staged) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------

int pick_block_n(int cout) {
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  return 256;
}

// what the last ssdsb_conv2d_bf16 call of this thread launched (ssdsb_conv_last_launch; tests assert
// that the BASELINE shapes really run the multi-way / weight-resident instantiations)
thread_local int g_last_launch[8] = {0, 0, 0, 0, 0, 0, 0, 0};

template <int BLOCK_N, int BLOCK_K, int WAYS>
int launch_ways(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                const CUtensorMap& tmR, const ConvKernelParams& kp, int grid, int smem, cudaStream_t st) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  const int m_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  const int info[8] = {BLOCK_N, BLOCK_K, WAYS, kp.b_resident, kp.stages, grid,
                       ((m_tiles + WAYS - 1) / WAYS) * kp.n_tiles, m_tiles % WAYS};
  for (int i = 0; i < 8; ++i) g_last_launch[i] = info[i];
  static bool configured = false;
  if (!configured) {
    SSDSB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, BLOCK_K, WAYS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, S::MAX_BYTES));
    configured = true;
  }
  SSDSB_CUDA(launch_pdl(conv_igemm_kernel<BLOCK_N, BLOCK_K, WAYS>, grid, CONV_NT, smem, st, tmA, tmB, tmY, tmR, kp));
  SSDSB_LAUNCH_CHECK("conv_igemm_kernel");
  return SSDSB_OK;
}

template <int BLOCK_N, int BLOCK_K>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
           const CUtensorMap& tmR, ConvKernelParams& kp, bool want_staging, int sms,
           cudaStream_t st) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  // 4 staging slots for residual layers (prefetch) and for short-K, write-heavy layers; 2 when the
  // operand ring needs the shared memory more (long-K, MMA-bound layers)
  // residual layers: 6 slots — the residual prefetch distance (slots - store_lag) must be >= 4 sub-tiles to
  // cover the HBM latency (r1l sweep: l1_conv3 222 us at distance 3, 200 us at >= 4, whatever the lag)
  kp.n_staging = want_staging ? (kp.residual ? 6 : (kp.num_k_blocks <= 8 ? 4 : 2)) : 0;
  if (const char* e = getenv("SSDSB_STAGING")) {            // experiment knob (profiling only)
    const int v = atoi(e);
    if (want_staging && v >= 2 && v <= MAX_STAGING && (kp.residual || kp.num_k_blocks <= 8)) kp.n_staging = v;
  }
  // two stores in flight (waiting for the read of the store just issued caps the store rate); with 2
  // slots the store must be read before its slot is re-armed
  kp.store_lag = kp.n_staging >= 4 ? 2 : 0;
  if (const char* e = getenv("SSDSB_STORE_LAG")) {          // experiment knob (profiling only)
    const int v = atoi(e);
    if (v >= 0 && v <= 6) kp.store_lag = v < kp.n_staging - 1 ? v : kp.n_staging - 1;
  }
  if (kp.store_lag < 0) kp.store_lag = 0;
  kp.tma_store = want_staging ? 1 : 0;
  kp.mma_converged = getenv("SSDSB_MMA_LANE0") ? 0 : 1;      // (A/B knob: the old single-lane issue loop)
  const int m_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  // interleaved accumulators: as many M tiles per group as TMEM allows (2 buffers x ways x BLOCK_N <= 512
  // columns), as long as there are enough groups to keep every SM busy
  kp.ways = 512 / (2 * BLOCK_N);
  if (kp.ways > 4) kp.ways = 4;
  if (const char* e = getenv("SSDSB_WAYS")) {               // experiment knob (profiling only)
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) kp.ways = v < kp.ways ? v : kp.ways;
  }
  while (kp.ways > 1 && ((m_tiles + kp.ways - 1) / kp.ways) * kp.n_tiles < 2 * sms) kp.ways >>= 1;
  const int tiles = ((m_tiles + kp.ways - 1) / kp.ways) * kp.n_tiles;   // groups
  int grid = tiles < sms ? tiles : sms;
  // weight-stationary mode: small weight slabs (<= 80 KiB per n-tile) are loaded once per CTA instead of
  // once per tile; needs every CTA to stay on one n_tile => grid must be a multiple of n_tiles
  const int bres_bytes = kp.num_k_blocks * S::B_STAGE_BYTES;
  kp.b_resident = 0;
  if (bres_bytes <= 80 * 1024 && tiles >= 2 * sms && bres_bytes < (1 << 20)) {
    const int g2 = grid - grid % kp.n_tiles;
    if (g2 >= sms / 2) {
      grid = g2;
      kp.b_resident = 1;
    }
  }
  kp.stages = S::stages_for(kp.n_staging, kp.b_resident, kp.num_k_blocks, kp.ways);
  if (kp.b_resident && kp.stages < 2) {                     // resident weights must not starve the A ring
    kp.b_resident = 0;
    grid = tiles < sms ? tiles : sms;
    kp.stages = S::stages_for(kp.n_staging, false, kp.num_k_blocks, kp.ways);
  }
  while (kp.ways > 1 && kp.stages < 2) {                    // (cannot happen with <= 4 ways of 16 KiB)
    kp.ways >>= 1;
    kp.stages = S::stages_for(kp.n_staging, kp.b_resident, kp.num_k_blocks, kp.ways);
  }
  if (kp.stages > kp.num_k_blocks * 4) kp.stages = kp.num_k_blocks * 4;
  if (kp.stages < 2) kp.stages = 2;
  const int smem = S::bytes(kp.stages, kp.n_staging, kp.b_resident, kp.num_k_blocks, kp.ways);
  const int groups = ((m_tiles + kp.ways - 1) / kp.ways) * kp.n_tiles;
  if (grid > groups) grid = groups;
  if constexpr (BLOCK_N <= 64) {
    if (kp.ways == 4) return launch_ways<BLOCK_N, BLOCK_K, 4>(tmA, tmB, tmY, tmR, kp, grid, smem, st);
  }
  if constexpr (BLOCK_N <= 128) {
    if (kp.ways == 2) return launch_ways<BLOCK_N, BLOCK_K, 2>(tmA, tmB, tmY, tmR, kp, grid, smem, st);
  }
  kp.ways = 1;
  return launch_ways<BLOCK_N, BLOCK_K, 1>(tmA, tmB, tmY, tmR, kp, grid, smem, st);
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_conv_last_launch(int* out8) {
  SSDSB_REQUIRE(out8, "conv_last_launch: NULL argument");
  for (int i = 0; i < 8; ++i) out8[i] = g_last_launch[i];
  return SSDSB_OK;
}

extern "C" int ssdsb_conv2d_bf16(const ssdsb_conv_desc* d, const void* x, const void* w,
                                 const float* bias, const void* residual, void* y, void* y2,
                                 void* stream) {
  SSDSB_REQUIRE(d && x && w && bias && y, "conv2d: NULL argument");
  SSDSB_REQUIRE(d->N >= 1 && d->H >= 1 && d->W >= 1 && d->Cin >= 1 && d->Cout >= 1,
                "conv2d: non-positive dimension");
  SSDSB_REQUIRE(d->KH >= 1 && d->KW >= 1 && d->stride >= 1 && d->stride <= 2 && d->pad >= 0,
                "conv2d: unsupported kernel geometry (KH=%d KW=%d stride=%d pad=%d)", d->KH, d->KW,
                d->stride, d->pad);
  const int Ho = d->Ho > 0 ? d->Ho : (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  const int Wo = d->Wo > 0 ? d->Wo : (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  SSDSB_REQUIRE(Ho >= 1 && Wo >= 1, "conv2d: empty output");
  const bool windowed = d->x_kind == SSDSB_CONV_X_WINDOWED_STEM;
  const int chunk = d->chunk;
  if (chunk) {
    SSDSB_REQUIRE(chunk % 32 == 0 && chunk <= 128 && d->Cin == d->Cout && d->Cin % chunk == 0 && !windowed &&
                      d->out_mode == CONV_OUT_NHWC_BF16,
                  "conv2d: chunked conv needs Cin == Cout, chunk %% 32 == 0, chunk <= 128 (chunk=%d Cin=%d)",
                  chunk, d->Cin);
    SSDSB_REQUIRE(d->w_rows >= (d->Cout / chunk) * 128, "conv2d: chunked weights need 128 rows per chunk");
  }
  // K-block = one swizzle row of channels: 64 (128B) when Cin allows it, else 32 (64B), else 16 (32B)
  const int cin_k = chunk ? chunk : d->Cin;       // channels contracted per output tile
  const int block_k = windowed ? 64 : (cin_k % 64 == 0 ? 64 : (cin_k % 32 == 0 ? 32 : 16));
  const int cs = d->x_cstride ? d->x_cstride : d->Cin;
  const int row_px = d->x_row_pixels ? d->x_row_pixels : d->W;
  if (windowed) {
    SSDSB_REQUIRE(d->Cin == 16 && d->KH == 4 && d->KW == 4 && d->stride == 1 && d->pad == 2 && cs == 16,
                  "conv2d: the windowed stem needs Cin=16, 4x4/s1/p2 on a packed s2d image");
    SSDSB_REQUIRE(row_px >= d->W + 3, "conv2d: windowed stem rows need >= W+3 pixels (2 left + 1 right pad)");
  } else {
    SSDSB_REQUIRE(d->Cin % block_k == 0, "conv2d: Cin=%d must be a multiple of 16", d->Cin);
    SSDSB_REQUIRE(row_px >= d->W, "conv2d: x_row_pixels < W");
  }
  SSDSB_REQUIRE(cs >= d->Cin && cs % 8 == 0, "conv2d: bad input channel stride %d", cs);
  SSDSB_REQUIRE(d->out_mode == CONV_OUT_NHWC_BF16 || d->out_mode == CONV_OUT_HEAD_NCHW_F32,
                "conv2d: bad out_mode");
  if (d->out_mode == CONV_OUT_NHWC_BF16) {
    SSDSB_REQUIRE(d->out_cstride >= d->Cout && d->out_cstride % 8 == 0 && d->Cout % 32 == 0,
                  "conv2d: NHWC output needs Cout %% 32 == 0 and a channel stride %% 8 == 0");
    SSDSB_REQUIRE(!residual || (d->res_cstride >= d->Cout && d->res_cstride % 8 == 0 &&
                                ((uintptr_t)residual & 15) == 0),
                  "conv2d: bad residual stride/alignment");
  } else {
    SSDSB_REQUIRE(y2 && d->n_loc >= 0 && d->n_loc <= d->Cout, "conv2d: head needs y2 and n_loc");
    SSDSB_REQUIRE(!residual, "conv2d: the head epilogue takes no residual");
  }
  SSDSB_REQUIRE(d->w_rows >= d->Cout, "conv2d: w_rows=%d < Cout=%d", d->w_rows, d->Cout);
  SSDSB_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0,
                "conv2d: pointers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(SSDSB_ERR_CUDA, "conv2d: cuTensorMapEncodeTiled entry point not found");

  // ---- M tile: (BW, BH, BN) output pixels, product <= 128 --------------------------------------
  int BW = Wo < 16 ? Wo : 16;
  if (Wo > 16 && Wo < 32) BW = Wo;                              // e.g. 19 -> 19x6 = 114 rows
  if (const char* e = getenv("SSDSB_TILE_W")) {                 // experiment knob (profiling only)
    const int v = atoi(e);
    if (v >= 1 && v <= BLOCK_M && v <= Wo) BW = v;
  }
  int BH = BLOCK_M / BW;
  if (BH > Ho) BH = Ho;
  int BN = BLOCK_M / (BW * BH);
  if (BN > d->N) BN = d->N;
  if (BN < 1) BN = 1;
  // TMA box limit: 256 per dimension (box = outputs * stride)
  SSDSB_REQUIRE(BW * d->stride <= 256 && BH * d->stride <= 256 && BN <= 256,
                "conv2d: tile too large for a TMA box");

  ConvKernelParams kp;
  kp.BW = BW; kp.BH = BH; kp.BN = BN;
  kp.tiles_w = (Wo + BW - 1) / BW;
  kp.tiles_h = (Ho + BH - 1) / BH;
  kp.tiles_n = (d->N + BN - 1) / BN;
  const int block_n = chunk ? 128 : pick_block_n(d->Cout);
  kp.n_tiles = chunk ? d->Cout / chunk : (d->Cout + block_n - 1) / block_n;
  kp.chunk_cin = chunk;
  kp.ntile_cout = chunk ? chunk : block_n;
  if (windowed) {          // taps = the 4 kernel rows; the 4 horizontal taps live inside the K-block
    kp.taps = 4; kp.KW = 1; kp.kc_per_tap = 1;
    kp.pad_w = 0; kp.pad_h = 2;
  } else {
    kp.taps = d->KH * d->KW; kp.KW = d->KW; kp.kc_per_tap = cin_k / block_k;
    kp.pad_w = d->pad; kp.pad_h = d->pad;
  }
  kp.num_k_blocks = kp.taps * kp.kc_per_tap;
  kp.stride = d->stride;
  kp.Ho = Ho; kp.Wo = Wo; kp.N = d->N;
  kp.Cout = d->Cout;
  kp.out_cstride = d->out_cstride;
  kp.res_cstride = d->res_cstride;
  kp.relu = d->relu;
  kp.mode = d->out_mode;
  kp.n_loc = d->n_loc;
  kp.sigmoid = d->sigmoid;
  kp.bias = bias;
  kp.residual = residual;
  kp.y = y; kp.y2 = y2;

  // ---- tensor maps ------------------------------------------------------------------------------
  const CUtensorMapSwizzle swz = block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmY;
  alignas(64) CUtensorMap tmR;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)row_px * cs * 2,
                             (cuuint64_t)d->H * row_px * cs * 2};
    cuuint32_t box[4] = {(cuuint32_t)block_k, (cuuint32_t)(BW * d->stride),
                         (cuuint32_t)(BH * d->stride), (cuuint32_t)BN};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    if (windowed) {   // window w = padded pixels [w, w+4) x 16 ch = 64 contiguous bf16, next window +32 B
      dims[0] = 64; dims[1] = (cuuint64_t)Wo;
      strides[0] = 32;
    }
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(SSDSB_ERR_CUDA, "conv2d: activation tensor map failed (CUresult %d)", (int)r);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)kp.num_k_blocks * block_k;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)d->w_rows};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(SSDSB_ERR_CUDA, "conv2d: weight tensor map failed (CUresult %d)", (int)r);
  }

  // staged TMA-store epilogue for every un-chunked NHWC layer (Cout % 32 == 0: a half-full last 64-column chunk is
  // clipped by the store's tensor map); SSDSB_DIRECT_RAGGED=1 sends Cout % 64 != 0 back to per-lane 16-byte stores
  const bool direct_ragged = getenv("SSDSB_DIRECT_RAGGED") != nullptr;        // (read per call: tests A/B it)
  const bool want_staging = d->out_mode == CONV_OUT_NHWC_BF16 && !chunk &&
                            ((d->Cout % 64) == 0 || !direct_ragged);
  if (want_staging) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)d->N};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    {
      cuuint64_t strides[3] = {(cuuint64_t)d->out_cstride * 2, (cuuint64_t)Wo * d->out_cstride * 2,
                               (cuuint64_t)Ho * Wo * d->out_cstride * 2};
      CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS)
        return fail(SSDSB_ERR_CUDA, "conv2d: output tensor map failed (CUresult %d)", (int)r);
    }
    if (residual) {
      cuuint64_t strides[3] = {(cuuint64_t)d->res_cstride * 2, (cuuint64_t)Wo * d->res_cstride * 2,
                               (cuuint64_t)Ho * Wo * d->res_cstride * 2};
      CUresult r = encode(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(residual), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS)
        return fail(SSDSB_ERR_CUDA, "conv2d: residual tensor map failed (CUresult %d)", (int)r);
    } else {
      tmR = tmY;
    }
  } else {
    tmY = tmA;
    tmR = tmA;
  }

  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SSDSB_CUDA(cudaGetDevice(&dev));
    SSDSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (block_k == 16) {
    switch (block_n) {
      case 64: return launch<64, 16>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
      case 128: return launch<128, 16>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
      default: return launch<256, 16>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
    }
  }
  if (block_k == 32) {
    switch (block_n) {
      case 64: return launch<64, 32>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
      case 128: return launch<128, 32>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
      default: return launch<256, 32>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
    }
  }
  switch (block_n) {
    case 64: return launch<64, 64>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
    case 128: return launch<128, 64>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
    default: return launch<256, 64>(tmA, tmB, tmY, tmR, kp, want_staging, sms, st);
  }
}
