// Two chained 1x1 convolutions in ONE persistent kernel (tcgen05 + TMEM + TMA), sm_100a only:
//
//   y1 = act1(conv1x1(x,  W1) + b1 [+ residual])        e.g. a bottleneck's conv3 + bn3 + add + ReLU
//   y2 = act2(conv1x1(y1, W2) + b2)                      e.g. the NEXT bottleneck's conv1 + bn1 + ReLU
//
// (reference: torchvision Bottleneck.forward as driven by ssds/modeling/nets/resnet.py:41-56; y1 is the
// block output every later consumer still needs, y2 is what the next block would otherwise compute by
// re-reading all of y1 from HBM.)  Both layers are pointwise, so tile t of layer 2 depends only on tile
// t of layer 1: each CTA walks its own M-tiles through
//
//   L1(t0) | L1(t1) | L1(t2) L2(t0) | L1(t3) L2(t1) | ...          (layer 2 lags LAG tiles behind)
//
// and loads the layer-2 A operand by TMA from the y1 tile it stored itself LAG tiles earlier — an L2
// hit, not a DRAM read — once the store engine has seen those stores complete (cp.async.bulk.wait_group
// without .read) and published the count through shared memory.  No cross-CTA synchronisation exists.
// y1 and y2 are bit-identical to two separate conv_igemm launches (same K-block order, same bf16 y1).
//
// Roles and barriers are those of conv_igemm.cu (w0 TMA producer, w1 MMA issuer, w2 TMEM allocator,
// w3 TMA store/residual engine, w4-11 epilogue); BLOCK_N = 256, BLOCK_K = 64, two 256-column
// accumulators that layer-1 and layer-2 work items use alternately.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace ssdsb {
namespace {

constexpr int P_BLOCK_N = 256;
constexpr int P_BLOCK_K = 64;
constexpr int P_MAX_STAGES = 8;
constexpr int P_MAX_STAGING = 8;
constexpr int P_STAGING_BYTES = BLOCK_M * 128;
constexpr int P_A_BYTES = BLOCK_M * 128;
constexpr int P_B_BYTES = P_BLOCK_N * 128;
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_BAR_BYTES = 512;
constexpr int P_MAX_SMEM = 232448;

struct PairLayer {
  int num_k_blocks;   // Cin / 64
  int n_tiles;        // ceil(Cout / 256)
  int Cout;
  int relu;
  const float* bias;
};

struct PairParams {
  PairLayer L[2];
  int has_res;        // layer 1 adds a residual (TMA-prefetched into the staging slot)
  int stages;
  int n_staging;      // staging slots (16 KiB each)
  int store_lag;      // TMA stores kept in flight before a slot is re-armed
  int lag;            // layer 2 runs this many of the CTA's tiles behind layer 1
  int done_lag;       // y1 completion of a tile is checked this many stores after its last store (<= 8)
  int BW, BH, BN, tiles_w, tiles_h, tiles_n;
  int Ho, Wo, N;
};

// The CTA's work items in issue order; every role walks the same sequence.
struct ItemIter {
  int n_my, lag, n1, n2;
  int i, sub;
  __device__ ItemIter(int n_my_, int lag_, int n1_, int n2_)
      : n_my(n_my_), lag(lag_), n1(n1_), n2(n2_), i(0), sub(0) {}
  __device__ bool next(int& layer, int& seq, int& n_tile) {
    while (i < n_my + lag) {
      const int c1 = (i < n_my) ? n1 : 0;
      const int c2 = (i >= lag) ? n2 : 0;
      if (sub < c1) {
        layer = 0; seq = i; n_tile = sub; ++sub;
        return true;
      }
      if (sub < c1 + c2) {
        layer = 1; seq = i - lag; n_tile = sub - c1; ++sub;
        return true;
      }
      ++i;
      sub = 0;
    }
    return false;
  }
};

struct PTile {
  int w0, h0, n0;
};
__device__ __forceinline__ PTile ptile(const PairParams& p, int m) {
  PTile t;
  t.w0 = (m % p.tiles_w) * p.BW;
  t.h0 = ((m / p.tiles_w) % p.tiles_h) * p.BH;
  t.n0 = (m / (p.tiles_w * p.tiles_h)) * p.BN;
  return t;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
template <int N>
__device__ __forceinline__ void tma_store_wait_done() {     // completion, not just the smem read
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_done_n(int n) {   // all but the n newest groups complete
  switch (n) {
    case 0: tma_store_wait_done<0>(); break;
    case 1: tma_store_wait_done<1>(); break;
    case 2: tma_store_wait_done<2>(); break;
    case 3: tma_store_wait_done<3>(); break;
    case 4: tma_store_wait_done<4>(); break;
    case 5: tma_store_wait_done<5>(); break;
    case 6: tma_store_wait_done<6>(); break;
    case 7: tma_store_wait_done<7>(); break;
    default: tma_store_wait_done<8>(); break;
  }
}

__global__ void __launch_bounds__(CONV_NT, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmR1,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmY2,
                 const __grid_constant__ PairParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int STAGES = p.stages;
  unsigned char* staging = smem + STAGES * P_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + p.n_staging * P_STAGING_BYTES);
  uint64_t* full_bar = bars;                              // [P_MAX_STAGES]
  uint64_t* empty_bar = bars + P_MAX_STAGES;              // [P_MAX_STAGES]
  uint64_t* tmem_full = bars + 2 * P_MAX_STAGES;          // [2]
  uint64_t* tmem_empty = bars + 2 * P_MAX_STAGES + 2;     // [2]
  uint64_t* slot_ready = bars + 2 * P_MAX_STAGES + 4;                   // [P_MAX_STAGING]
  uint64_t* slot_full = bars + 2 * P_MAX_STAGES + 4 + P_MAX_STAGING;    // [P_MAX_STAGING]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * P_MAX_STAGES + 4 + 2 * P_MAX_STAGING);
  uint32_t* y1_done = tmem_holder + 1;      // number of this CTA's tiles whose y1 is complete in global

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t tmem_cols = 2 * P_BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmY1);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmY2);
    if (p.has_res) tma_prefetch_desc(&tmR1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    for (int r = 0; r < P_MAX_STAGING; ++r) {
      mbar_init(&slot_ready[r], 1);
      mbar_init(&slot_full[r], 8);
    }
    *y1_done = 0;
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
  pdl_wait();
  pdl_trigger();

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int n_my = ((int)blockIdx.x < m_tiles) ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int rows = p.BW * p.BH * p.BN;
  const int n1 = p.L[0].n_tiles, n2 = p.L[1].n_tiles;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = (uint32_t)rows * 128u;
      ItemIter it(n_my, p.lag, n1, n2);
      int layer, seq, nt;
      while (it.next(layer, seq, nt)) {
        const PTile t = ptile(p, (int)blockIdx.x + seq * (int)gridDim.x);
        const PairLayer& Lr = p.L[layer];
        const uint32_t b_bytes = (uint32_t)min(P_BLOCK_N, Lr.Cout - nt * P_BLOCK_N) * 128u;
        const CUtensorMap* ma = layer ? &tmY1 : &tmA1;
        const CUtensorMap* mb = layer ? &tmB2 : &tmB1;
        if (layer == 1 && nt == 0) {
          while (ld_acquire_u32(y1_done) <= (uint32_t)seq) {
          }
          fence_proxy_async_all();      // y1 was written through the async proxy and is read through it
        }
        for (int kb = 0; kb < Lr.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * P_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
          tma_load_4d(sa, ma, &full_bar[stage], kb * P_BLOCK_K, t.w0, t.h0, t.n0);
          tma_load_2d(sa + P_A_BYTES, mb, &full_bar[stage], kb * P_BLOCK_K, nt * P_BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      ItemIter it(n_my, p.lag, n1, n2);
      int layer, seq, nt;
      while (it.next(layer, seq, nt)) {
        const PairLayer& Lr = p.L[layer];
        const uint32_t idesc = make_idesc(BLOCK_M, min(P_BLOCK_N, Lr.Cout - nt * P_BLOCK_N));
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P_BLOCK_N);
        for (int kb = 0; kb < Lr.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t a_desc = make_smem_desc<128>(smem_u32(smem + stage * P_STAGE_BYTES));
          const uint64_t b_desc = a_desc + (uint64_t)(P_A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < P_BLOCK_K / UMMA_K; ++k)
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tcgen05_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ======================= store / residual engine =======================
    if (lane == 0) {
      const int R = p.n_staging;
      // arming iterator: runs R chunks ahead of the stores
      ItemIter ait(n_my, p.lag, n1, n2);
      int a_layer = 0, a_seq = 0, a_nt = 0, a_chunk = 0, a_nch = 0, a_slot = 0;
      bool a_valid = false;
      auto arm_next = [&]() {
        if (!a_valid || a_chunk == a_nch) {
          a_valid = ait.next(a_layer, a_seq, a_nt);
          if (!a_valid) return;
          a_chunk = 0;
          a_nch = min(P_BLOCK_N, p.L[a_layer].Cout - a_nt * P_BLOCK_N) >> 6;
        }
        const int slot = a_slot;
        if (++a_slot == R) a_slot = 0;
        if (a_layer == 0 && p.has_res) {
          const PTile ta = ptile(p, (int)blockIdx.x + a_seq * (int)gridDim.x);
          mbar_expect_tx(&slot_ready[slot], (uint32_t)rows * 128u);
          tma_load_4d(staging + slot * P_STAGING_BYTES, &tmR1, &slot_ready[slot],
                      a_nt * P_BLOCK_N + a_chunk * 64, ta.w0, ta.h0, ta.n0);
        } else {
          mbar_arrive(&slot_ready[slot]);
        }
        ++a_chunk;
      };
      for (int i = 0; i < R; ++i) arm_next();

      uint32_t published = 0;   // tiles whose y1 completion has been published
      uint32_t pending = 0;     // tiles whose last y1 store has been issued
      int since = 0;            // stores issued since `pending` last grew
      int g = 0;
      SlotRing sring = {0, 0u};
      ItemIter it(n_my, p.lag, n1, n2);
      int layer, seq, nt;
      while (it.next(layer, seq, nt)) {
        const PTile t = ptile(p, (int)blockIdx.x + seq * (int)gridDim.x);
        const int nch = min(P_BLOCK_N, p.L[layer].Cout - nt * P_BLOCK_N) >> 6;
        const CUtensorMap* my = layer ? &tmY2 : &tmY1;
        if (layer == 1 && nt == 0 && published <= (uint32_t)seq) {
          // the producer needs y1 of this tile before anything of this item can reach us (short CTAs /
          // one-chunk tiles): flush instead of waiting for two more stores that cannot come
          tma_store_wait_done<0>();
          fence_proxy_async_all();
          st_release_u32(y1_done, pending);
          published = pending;
        }
        for (int c = 0; c < nch; ++c, ++g) {
          mbar_wait(&slot_full[sring.slot], sring.phase);
          tma_store_4d(my, staging + sring.slot * P_STAGING_BYTES, nt * P_BLOCK_N + c * 64, t.w0, t.h0, t.n0);
          tma_store_commit();
          sring.advance(R);
          ++since;
          if (g >= p.store_lag) {
            tma_store_wait_read_n(p.store_lag);   // store g-lag has left smem -> re-arm its slot (residual prefetch)
            arm_next();
          }
          if (layer == 0 && nt == n1 - 1 && c == nch - 1) {
            pending = (uint32_t)seq + 1u;
            since = 0;
          } else if (pending > published && since >= p.done_lag) {
            // y1 of tile `pending-1`: every store but the `done_lag` newest has completed; checking
            // late keeps this thread (which also issues the stores and re-arms the slots) from waiting
            tma_store_wait_done_n(p.done_lag);
            fence_proxy_async_all();
            st_release_u32(y1_done, pending);
            published = pending;
          }
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int R = p.n_staging;
    SlotRing ring = {0, 0u};
    const uint32_t staging_addr = smem_u32(staging);
    ItemIter it(n_my, p.lag, n1, n2);
    int layer, seq, nt;
    while (it.next(layer, seq, nt)) {
      const PairLayer& Lr = p.L[layer];
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * P_BLOCK_N);
      const int nchunks = min(P_BLOCK_N, Lr.Cout - nt * P_BLOCK_N) >> 6;
      staged_epilogue_item(t_row, nchunks, Lr.bias + nt * P_BLOCK_N + half * 32, staging_addr, P_STAGING_BYTES,
                           R, ring, slot_ready, slot_full, &tmem_empty[acc], (layer == 0) && p.has_res,
                           Lr.relu, r, half, lane);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                 : "memory");
  }
}

int encode_act_map(EncodeTiledFn encode, CUtensorMap* tm, const void* base, int C, int W, int H, int N,
                   int cstride, int BW, int BH, int BN, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)W * cstride * 2, (cuuint64_t)H * W * cstride * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "conv1x1_pair: %s tensor map failed (CUresult %d)", what, (int)r);
  return SSDSB_OK;
}

int encode_w_map(EncodeTiledFn encode, CUtensorMap* tm, const void* w, int K, int rows, int box_rows,
                 const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SSDSB_ERR_CUDA, "conv1x1_pair: %s tensor map failed (CUresult %d)", what, (int)r);
  return SSDSB_OK;
}

}  // namespace
}  // namespace ssdsb

using namespace ssdsb;

extern "C" int ssdsb_conv1x1_pair_bf16(int N, int H, int W, int Cin, int Cmid, int Cout2, int relu1, int relu2,
                                       const void* x, const void* w1, const float* bias1,
                                       const void* residual, void* y1, const void* w2, const float* bias2,
                                       void* y2, void* stream) {
  SSDSB_REQUIRE(x && w1 && bias1 && y1 && w2 && bias2 && y2, "conv1x1_pair: NULL argument");
  SSDSB_REQUIRE(N >= 1 && H >= 1 && W >= 1, "conv1x1_pair: non-positive dimension");
  SSDSB_REQUIRE(Cin >= 64 && Cin % 64 == 0 && Cmid >= 64 && Cmid % 64 == 0 && Cout2 >= 64 && Cout2 % 64 == 0,
                "conv1x1_pair: channel counts must be multiples of 64 (Cin=%d Cmid=%d Cout2=%d)", Cin, Cmid,
                Cout2);
  SSDSB_REQUIRE((Cmid <= 256 || Cmid % 256 == 0) && (Cout2 <= 256 || Cout2 % 256 == 0),
                "conv1x1_pair: channel counts above 256 must be multiples of 256");
  SSDSB_REQUIRE((((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)y1 | (uintptr_t)y2 |
                  (uintptr_t)residual) & 15) == 0,
                "conv1x1_pair: pointers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(SSDSB_ERR_CUDA, "conv1x1_pair: cuTensorMapEncodeTiled entry point not found");

  int BW = W < 16 ? W : 16;
  if (W > 16 && W < 32) BW = W;
  int BH = BLOCK_M / BW;
  if (BH > H) BH = H;
  int BN = BLOCK_M / (BW * BH);
  if (BN > N) BN = N;
  if (BN < 1) BN = 1;

  PairParams kp;
  kp.BW = BW; kp.BH = BH; kp.BN = BN;
  kp.tiles_w = (W + BW - 1) / BW;
  kp.tiles_h = (H + BH - 1) / BH;
  kp.tiles_n = (N + BN - 1) / BN;
  kp.Ho = H; kp.Wo = W; kp.N = N;
  kp.L[0].num_k_blocks = Cin / 64;
  kp.L[0].n_tiles = (Cmid + P_BLOCK_N - 1) / P_BLOCK_N;
  kp.L[0].Cout = Cmid;
  kp.L[0].relu = relu1;
  kp.L[0].bias = bias1;
  kp.L[1].num_k_blocks = Cmid / 64;
  kp.L[1].n_tiles = (Cout2 + P_BLOCK_N - 1) / P_BLOCK_N;
  kp.L[1].Cout = Cout2;
  kp.L[1].relu = relu2;
  kp.L[1].bias = bias2;
  kp.has_res = residual ? 1 : 0;
  kp.lag = 2;
  kp.done_lag = 2;
  if (const char* e = getenv("SSDSB_PAIR_LAG")) {          // experiment knobs (profiling only)
    const int v = atoi(e);
    if (v >= 1 && v <= 8) kp.lag = v;
  }
  if (const char* e = getenv("SSDSB_PAIR_DONE_LAG")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 8) kp.done_lag = v;
  }
  // residual prefetch distance (slots - store_lag) = 4 sub-tiles as in conv_igemm.cu, and 3 operand
  // stages (r1l sweep: 6 slots / 2 stages is 10 % slower on the 64x64 stage, 4 slots 5 % slower on 128x128)
  kp.n_staging = 5;
  kp.store_lag = 1;
  if (const char* e = getenv("SSDSB_PAIR_STAGING")) {      // experiment knobs (profiling only)
    const int v = atoi(e);
    if (v >= 2 && v <= P_MAX_STAGING) kp.n_staging = v;
  }
  if (const char* e = getenv("SSDSB_PAIR_STORE_LAG")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 6) kp.store_lag = v;
  }
  if (kp.store_lag > kp.n_staging - 1) kp.store_lag = kp.n_staging - 1;
  kp.stages = (P_MAX_SMEM - 1024 - P_BAR_BYTES - kp.n_staging * P_STAGING_BYTES) / P_STAGE_BYTES;
  if (kp.stages > P_MAX_STAGES) kp.stages = P_MAX_STAGES;

  alignas(64) CUtensorMap tmA1, tmB1, tmY1, tmR1, tmB2, tmY2;
  int rc;
  if ((rc = encode_act_map(encode, &tmA1, x, Cin, W, H, N, Cin, BW, BH, BN, "input")) != SSDSB_OK) return rc;
  if ((rc = encode_act_map(encode, &tmY1, y1, Cmid, W, H, N, Cmid, BW, BH, BN, "y1")) != SSDSB_OK) return rc;
  if ((rc = encode_act_map(encode, &tmY2, y2, Cout2, W, H, N, Cout2, BW, BH, BN, "y2")) != SSDSB_OK) return rc;
  if (residual) {
    if ((rc = encode_act_map(encode, &tmR1, residual, Cmid, W, H, N, Cmid, BW, BH, BN, "residual")) != SSDSB_OK)
      return rc;
  } else {
    tmR1 = tmY1;
  }
  if ((rc = encode_w_map(encode, &tmB1, w1, Cin, Cmid, Cmid < P_BLOCK_N ? Cmid : P_BLOCK_N, "w1")) != SSDSB_OK)
    return rc;
  if ((rc = encode_w_map(encode, &tmB2, w2, Cmid, Cout2, Cout2 < P_BLOCK_N ? Cout2 : P_BLOCK_N, "w2")) != SSDSB_OK)
    return rc;

  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SSDSB_CUDA(cudaGetDevice(&dev));
    SSDSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static bool configured = false;
  if (!configured) {
    SSDSB_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_MAX_SMEM));
    configured = true;
  }
  const int m_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  const int grid = m_tiles < sms ? m_tiles : sms;
  const int smem = kp.stages * P_STAGE_BYTES + kp.n_staging * P_STAGING_BYTES + 1024 + P_BAR_BYTES;
  SSDSB_CUDA(launch_pdl(conv_pair_kernel, grid, CONV_NT, smem, (cudaStream_t)stream, tmA1, tmB1, tmY1, tmR1, tmB2,
                        tmY2, kp));
  SSDSB_LAUNCH_CHECK("conv_pair_kernel");
  return SSDSB_OK;
}
