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
// Warp roles (512 threads): w0 TMA producer, w1 expand-MMA issuer, w2 TMEM allocator + project-MMA issuer,
// w3 store / residual engine,
// w4-7 convert + epilogue (TMEM lane quarter = warp % 4), w8-15 depthwise.
// The two GEMMs are issued by two threads on two warp schedulers: a single issuing thread (~15 dependent
// instructions per tcgen05 op among busy warps) was the bottleneck of the pipeline.
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
constexpr int MB_MAX_BYTES = 232448 - 3072;   // minus the static wait-profile counters / event log

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
  int hp_pitch;            // bytes per pixel row of hpatch: HC*2 + 16 (convert writes conflict-free), dense under TMA
  int off_x, off_we, off_wp, off_hp, off_a, off_stage, off_bar;
  const float* b_exp;
  const float* b_dw;
  const float* b_proj;
  const uint2* w_dw;       // [9][hid / 4] bf16 quads
  int dbg;                 // SSDSB_MB_DEBUG bits (bisecting aid): 1 no dead-half skip, 2 always refetch dw weights
  int prof;                // 1: CTA 0 records its per-warp barrier wait cycles (ssdsb_mbconv_profile)
};

__device__ __forceinline__ uint64_t make_smem_desc_rt(uint32_t smem_addr, int row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((8u * (uint32_t)row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}
// tcgen05.mma with the two shared-memory descriptors given as (low word = address >> 4 [+ k offset], constant high
// word): the single issuing thread only does 32-bit adds per MMA
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory accesses of the depthwise stage: volatile (ordered after the barrier waits and among themselves,
// exactly as written) but without a memory clobber, so the FMAs in between schedule freely
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 o;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o.x), "=r"(o.y) : "r"(addr));
  return o;
}
__device__ __forceinline__ void sts_u2_pred(uint32_t addr, uint32_t a, uint32_t b, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.shared.v2.u32 [%0], {%1, %2};\n\t}" ::"r"(addr),
               "r"(a), "r"(b), "r"((int)pred));
}
// two fp32 lanes in one 64-bit register pair: fma.rn.f32x2 issues two IEEE FMAs as one instruction (same results
// as two fmaf, half the issue slots — the depthwise stage is issue-bound)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// clamp of two packed bf16 values: rounding to bf16 is monotone and both bounds (0, 6, +-inf) are bf16 numbers, so
// clamp(round(x)) == round(clamp(x)) — one max + one min per PAIR instead of two each per element
__device__ __forceinline__ uint32_t clamp_bf16x2(uint32_t v, uint32_t lo2, uint32_t hi2) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\tmax.bf16x2 t, %1, %2;\n\tmin.bf16x2 %0, t, %3;\n\t}" : "=r"(r) : "r"(v), "r"(lo2), "r"(hi2));
  return r;
}
__device__ __forceinline__ uint32_t relu_lo2(int relu) { return relu ? 0x00000000u : 0xff80ff80u; }        // 0 | -inf
__device__ __forceinline__ uint32_t relu_hi2(int relu) { return relu == 2 ? 0x40c040c0u : 0x7f807f80u; }   // 6 | +inf
// 4 bf16 (8 bytes) -> two packed pairs
__device__ __forceinline__ void bf4_to_f2(const uint2 v, f2_t (&f)[2]) {
  f[0] = pk2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
  f[1] = pk2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
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
// optional wait profile (ssdsb_mbconv_profile): cycles each warp spent blocked per barrier id, CTA 0 only
__device__ unsigned long long g_mb_prof[16 * 16 + 2 + 64];   // + event log: chunks 8..15 x 8 events
__shared__ unsigned long long s_mb_ev[64];
__shared__ unsigned long long s_mb_prof[16 * 16];
__shared__ int s_mb_prof_on;
// PROF is a template parameter of the kernel (0: production, none of this; 1: event stamps only — a clock read
// and a shared store per event, light enough not to change the schedule; 2: also per-barrier wait cycles)
template <int PROF>
__device__ __forceinline__ void mbw(uint64_t* bar, uint32_t parity, int id) {
  // (try_wait itself may suspend the thread up to a hardware time limit: the whole call is timed)
  long long t0 = 0;
  if (PROF == 2) t0 = clock64();
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) mb_wait_fail(id, parity);
  if (PROF == 2 && s_mb_prof_on && (threadIdx.x & 31) == 0)
    atomicAdd(&s_mb_prof[(threadIdx.x >> 5) * 16 + id], (unsigned long long)(clock64() - t0));
}
// slot 0 of a warp's row: cycles of its own work section (lane 0 only)
template <int PROF>
__device__ __forceinline__ void mb_prof_work(long long t0) {
  if (PROF == 2 && s_mb_prof_on && (threadIdx.x & 31) == 0)
    atomicAdd(&s_mb_prof[(threadIdx.x >> 5) * 16], (unsigned long long)(clock64() - t0));
}

// event log (profiling): cycle stamp, relative to the CTA's start, of event k of chunk g (8 <= g < 16)
template <int PROF>
__device__ __forceinline__ void mb_event(int g, int k, long long t0) {
  if (PROF && s_mb_prof_on && g >= 8 && g < 16) s_mb_ev[(g - 8) * 8 + k] = (unsigned long long)(clock64() - t0);
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

template <int HC, int PROF>
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
  const long long t_start = PROF ? clock64() : 0ll;
  if (PROF) {
    if (threadIdx.x < 256) s_mb_prof[threadIdx.x] = 0ull;
    if (threadIdx.x < 64) s_mb_ev[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) s_mb_prof_on = (blockIdx.x == 0) ? 1 : 0;
  }

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

  const int my_tiles = ((int)p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = my_tiles * nch;           // chunks this CTA processes; chunk g: tile g / nch, hidden chunk g % nch

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // expand weights run two chunks ahead of the project weights (the MMA thread issues expand(g+2) before
    // project(g)); with two x buffers the next tile's patch is requested a whole tile ahead
    if (lane == 0) {
      auto load_x = [&](int tc) {
        const MbTile t = mb_tile(p, (int)blockIdx.x + tc * (int)gridDim.x);
        const int xb = tc % p.n_xbuf;
        mbw<PROF>(&x_empty[xb], (uint32_t)(((tc / p.n_xbuf) & 1) ^ 1), 13);
        mbar_expect_tx(&x_full[xb], (uint32_t)(p.xkb * p.PP * rowb_x));
        for (int kb = 0; kb < p.xkb; ++kb)
          tma_load_4d(xbuf + xb * p.x_bytes + kb * p.xkb_bytes, &tmX, &x_full[xb], kb * p.blk_x, t.w0 * p.S - 1,
                      t.h0 * p.S - 1, t.n);
      };
      auto load_we = [&](int ge) {
        const int s = ge & 1;
        // (the expand of chunk ge - 2 signals ONE barrier, eacc_full: its accumulator is ready and its weight
        // stage is free — a tcgen05.commit costs the single MMA thread ~100 cycles)
        mbw<PROF>(&eacc_full[s], (uint32_t)(((ge >> 1) & 1) ^ 1), 1);
        mbar_expect_tx(&we_full[s], (uint32_t)(p.xkb * HC * rowb_x));
        for (int kb = 0; kb < p.xkb; ++kb)
          tma_load_2d(webuf + s * p.we_bytes + kb * HC * rowb_x, &tmWe, &we_full[s], kb * p.blk_x, (ge % nch) * HC);
      };
      if (p.has_expand) {
        load_x(0);
        if (G > 0) load_we(0);
        if (G > 1) load_we(1);
      }
      for (int g = 0; g < G; ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        const int hh = g % nch, tc = g / nch;
        if (p.has_expand) {
          // next tile's x patch: a tile ahead with two buffers, else as soon as this tile's last expand is issued
          if (tc + 1 < my_tiles && hh == (p.n_xbuf == 2 ? 0 : nch - 1)) load_x(tc + 1);
          if (g + 2 < G) load_we(g + 2);
        } else {
          const MbTile t = mb_tile(p, (int)blockIdx.x + tc * (int)gridDim.x);
          mbw<PROF>(&hp_empty[s], ph ^ 1, 2);
          mbar_expect_tx(&hp_full[s], (uint32_t)(p.PP * RB));
          tma_load_4d(hpbuf + s * p.hp_bytes, &tmX, &hp_full[s], hh * HC, t.w0 * p.S - 1, t.h0 * p.S - 1, t.n);
        }
        mbw<PROF>(&a_empty[s], ph ^ 1, 3);      // project(g - 2) done: A stage and project-weight stage are free
        mbar_expect_tx(&wp_full[s], (uint32_t)(p.block_n * RB));
        tma_load_2d(wpbuf + s * p.wp_bytes, &tmWp, &wp_full[s], hh * HC, 0);
      }
    }
  } else if (warp == 1) {
    // =============================== expand MMA issuer ===============================
    // runs ahead as far as the two accumulator / weight stages allow: expand(g+2) only needs convert(g) to have
    // drained its accumulator, which happens BEFORE the depthwise warps start on chunk g — its result is waiting in
    // TMEM when they are done.  Within an expand the MMAs of the PM accumulators are interleaved (back-to-back
    // MMAs on one accumulator retire only every ~146 cycles).
    if (lane == 0) {
      // A single thread runs this: every instruction costs its full latency, so the loop carries chunk / tile
      // counters incrementally (no divisions) and descriptors as 32-bit (address >> 4) values; the upper
      // descriptor word is constant per operand kind.
      const uint32_t idesc_e = make_idesc(BLOCK_M, HC);
      const int ksteps = p.blk_x / UMMA_K;
      const uint32_t hi_x = (uint32_t)(make_smem_desc_rt(0u, rowb_x) >> 32);
      auto lo = [](uint32_t addr) { return (addr & 0x3ffffu) >> 4; };
      const uint32_t x_lo = lo(smem_u32(xbuf)), we_lo = lo(smem_u32(webuf));
      const uint32_t x_buf_step = (uint32_t)p.x_bytes >> 4, x_kb_step = (uint32_t)p.xkb_bytes >> 4;
      const uint32_t x_m_step = (uint32_t)(BLOCK_M * rowb_x) >> 4;
      const uint32_t we_s_step = (uint32_t)p.we_bytes >> 4, we_kb_step = (uint32_t)(HC * rowb_x) >> 4;
      // expand sequence state (chunk ge): stage, phase, hidden chunk, x buffer / its phase
      int e_g = 0, e_hh = 0, e_xb = 0;
      uint32_t e_xph = 0;
      auto expand = [&]() {
        const int s = e_g & 1;
        const uint32_t ph = (uint32_t)((e_g >> 1) & 1);
        if (e_hh == 0) mbw<PROF>(&x_full[e_xb], e_xph, 15);
        mbw<PROF>(&we_full[s], ph, 6);
        mbw<PROF>(&eacc_empty[s], ph ^ 1, 7);
        mb_event<PROF>(e_g, 7, t_start);
        tcgen05_fence_after();
        const uint32_t xa = x_lo + (uint32_t)e_xb * x_buf_step;
        const uint32_t wa = we_lo + (uint32_t)s * we_s_step;
        const uint32_t d0 = tmem_base + eacc_col0 + (uint32_t)(s * p.PM * HC);
        for (int kb = 0; kb < p.xkb; ++kb) {
          const uint32_t xk = xa + (uint32_t)kb * x_kb_step, wk = wa + (uint32_t)kb * we_kb_step;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            uint32_t am = xk + (uint32_t)(2 * k), dm = d0;
            for (int m = 0; m < p.PM; ++m, am += x_m_step, dm += (uint32_t)HC)      // accumulators interleaved
              umma_bf16_lohi(dm, am, hi_x, wk + (uint32_t)(2 * k), hi_x, idesc_e, acc);
          }
        }
        tcgen05_commit(&eacc_full[s]);
        mb_event<PROF>(e_g, 4, t_start);
        ++e_g;
        if (++e_hh == nch) {
          e_hh = 0;
          tcgen05_commit(&x_empty[e_xb]);
          if (++e_xb == p.n_xbuf) {
            e_xb = 0;
            e_xph ^= 1u;
          }
        }
      };
      if (p.has_expand)
        for (int g = 0; g < G; ++g) expand();
    }
  } else if (warp == 2) {
    // =============================== project MMA issuer ===============================
    // (a second issuing thread on another scheduler: one thread running both GEMMs was the bottleneck of the
    // whole pipeline — ~15 dependent instructions per tcgen05.mma / commit at ~10 cycles each among busy warps)
    if (lane == 0) {
      const uint32_t idesc_p = make_idesc(BLOCK_M, p.block_n);
      const uint32_t hi_h = (uint32_t)(make_smem_desc_rt(0u, RB) >> 32);
      const uint32_t a_lo = (smem_u32(abuf) & 0x3ffffu) >> 4, wp_lo = (smem_u32(wpbuf) & 0x3ffffu) >> 4;
      const uint32_t a_s_step = (uint32_t)p.a_bytes >> 4, wp_s_step = (uint32_t)p.wp_bytes >> 4;
      int p_hh = 0;
      uint32_t p_tph = 0;           // parity of the tile count: pacc barriers
      for (int gp = 0; gp < G; ++gp) {
        const int s = gp & 1;
        const uint32_t ph = (uint32_t)((gp >> 1) & 1);
        mb_event<PROF>(gp, 2, t_start);
        mbw<PROF>(&a_full[s], ph, 4);
        mbw<PROF>(&wp_full[s], ph, 5);
        if (p_hh == 0) mbw<PROF>(pacc_empty, p_tph ^ 1u, 14);
        mb_event<PROF>(gp, 3, t_start);
        tcgen05_fence_after();
        const uint32_t aa = a_lo + (uint32_t)s * a_s_step, wb = wp_lo + (uint32_t)s * wp_s_step;
#pragma unroll
        for (int k = 0; k < HC / UMMA_K; ++k)
          umma_bf16_lohi(tmem_base, aa + (uint32_t)(2 * k), hi_h, wb + (uint32_t)(2 * k), hi_h, idesc_p,
                         (p_hh > 0 || k > 0) ? 1u : 0u);
        tcgen05_commit(&a_empty[s]);
        if (++p_hh == nch) {
          p_hh = 0;
          p_tph ^= 1u;
          tcgen05_commit(pacc_full);
        }
      }
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
          mbw<PROF>(&slot_full[sring.slot], sring.phase, 8);
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
    const uint32_t hi_e = relu_hi2(p.relu_e), lo_e = relu_lo2(p.relu_e);
    SlotRing ring = {0, 0u};
    auto epilogue = [&](int tc) {
      mbw<PROF>(pacc_full, (uint32_t)(tc & 1), 9);
      tcgen05_fence_after();
      for (int c = 0; c < nchunks_out; ++c) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          SlotRing rr = ring;
          if (p.Cout - c * 64 - half * 32 <= 0 && !(p.dbg & 1)) {
            // this warp's 32 columns lie beyond Cout (Cout % 64 == 32, e.g. every 32-channel block output): no
            // arithmetic, only the barriers' arrival counts (slot_ready first: an early slot_full arrival would
            // count towards the slot's previous use)
            if (c == nchunks_out - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(pacc_empty);
            }
            mbar_wait(&slot_ready[rr.slot], rr.phase);
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_full[rr.slot]);
            rr.advance(p.R);
          } else {
            staged_epilogue_item(t_lane + (uint32_t)(c * 64), 1, p.b_proj + c * 64 + half * 32, staging_addr,
                                 MB_STAGING_BYTES, p.R, rr, slot_ready, slot_full,
                                 c == nchunks_out - 1 ? pacc_empty : nullptr, p.has_res != 0, p.relu_p, r_tile, half,
                                 lane, p.Cout - c * 64 - half * 32);
          }
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
        mbw<PROF>(&eacc_full[s], ph, 10);
        if (lane == 0 && q == 0) mb_event<PROF>(g, 5, t_start);
        mbw<PROF>(&hp_empty[s], ph ^ 1, 2);
        tcgen05_fence_after();
        const uint32_t hp_addr = smem_u32(hpbuf + s * p.hp_bytes);
        const float* bias = p.b_exp + h * HC;
        const long long t_work = PROF == 2 ? clock64() : 0ll;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          if (m >= p.PM || m * BLOCK_M + q * 32 >= p.PP) break;        // warp-uniform
          const int r = m * BLOCK_M + r_tile;
          // pixels outside the image: clamp to [0, 0] (the depthwise conv zero-pads the expanded tensor)
          const uint32_t lo_k = inimg[m] ? lo_e : 0u, hi_k = inimg[m] ? hi_e : 0u;
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
                uint4 o;
                o.x = clamp_bf16x2(pack_bf16(f[0], f[1]), lo_k, hi_k);
                o.y = clamp_bf16x2(pack_bf16(f[2], f[3]), lo_k, hi_k);
                o.z = clamp_bf16x2(pack_bf16(f[4], f[5]), lo_k, hi_k);
                o.w = clamp_bf16x2(pack_bf16(f[6], f[7]), lo_k, hi_k);
                // 16-byte chunk (half * 4 + gq) of pixel r; the padded pitch spreads the 8 lanes of a store phase
                // over all banks
                sts_u4(hp_addr + (uint32_t)(r * p.hp_pitch + (half * 4 + gq) * 16), o);
              }
            }
          }
        }
        mb_prof_work<PROF>(t_work);
        if (lane == 0 && q == 0) mb_event<PROF>(g, 6, t_start);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&eacc_empty[s]);
          mbar_arrive(&hp_full[s]);
        }
        // the previous tile's epilogue runs after this tile's SECOND chunk has been handed to the depthwise warps
        // (convert of that chunk had to wait for dw of the previous tile's last chunk anyway, so the project MMAs
        // it needs are complete or about to be, and the depthwise warps have two chunks of work queued meanwhile)
        if (pending && h == (nch > 1 ? 1 : 0)) {
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
    const uint32_t hi_d = relu_hi2(p.relu_d), lo_d = relu_lo2(p.relu_d);
    const int n_items = p.nseg * p.BW * QD;
    const int hq = p.hid >> 2;
    const bool one_item = n_items <= MB_DW_THREADS;
    const int it_qd = tid & (QD - 1);
    const int it_seg = (tid / QD) / p.BW;
    const int it_bw = (tid / QD) - it_seg * p.BW;
    const int it_r0 = it_seg * p.rs;
    const int it_rows = min(p.rs, p.BH - it_r0);
    const bool it_active = tid < n_items && it_rows > 0;
    uint2 wraw[9];
    float4 braw;
    auto fetch_weights = [&](int hh) {
      const uint2* wp = p.w_dw + hh * QD + it_qd;
#pragma unroll
      for (int k = 0; k < 9; ++k) wraw[k] = __ldg(wp + (size_t)k * hq);
      braw = __ldg(reinterpret_cast<const float4*>(p.b_dw + hh * HC) + it_qd);
    };
    if (one_item && it_active && G > 0) fetch_weights(0);
    int g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int h = 0; h < nch; ++h, ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        mbw<PROF>(&hp_full[s], ph, 11);
        mbw<PROF>(&a_empty[s], ph ^ 1, 12);
        const uint32_t hp_addr = smem_u32(hpbuf + s * p.hp_bytes);
        const uint32_t a_addr = smem_u32(abuf + s * p.a_bytes);
        const long long t_work = PROF == 2 ? clock64() : 0ll;
        if (tid == 0) mb_event<PROF>(g, 0, t_start);
        // one item = (row segment, output column, channel quad): its 9 x 4 folded weights and bias come from global
        // memory (L1 / L2).  With at most one item per thread (the geometry search aims for exactly that) the
        // decomposition is done once per kernel and the NEXT chunk's weights are requested right after this chunk's
        // arithmetic, so their latency hides behind the fence / arrive / wait of the hand-over.
        auto process = [&](const int qd, const int bw, const int r0, const int rows, const f2_t (&wf)[9][2],
                           const f2_t (&b2)[2]) {
            const uint32_t pitch = (uint32_t)p.hp_pitch;
            const uint32_t rsb = (uint32_t)p.PW * pitch;                 // one patch row down
            struct Raw {
              uint2 l, m, r;
            };
            auto load3 = [&](uint32_t ptr) -> Raw {
              Raw v;
              v.l = lds_u2(ptr);
              v.m = lds_u2(ptr + pitch);
              v.r = lds_u2(ptr + 2 * pitch);
              return v;
            };
            auto fma_row = [&](f2_t (&a)[2], const Raw& v, int dy) {
              f2_t f[2];
              bf4_to_f2(v.l, f);
              a[0] = ffma2(f[0], wf[dy * 3][0], a[0]);
              a[1] = ffma2(f[1], wf[dy * 3][1], a[1]);
              bf4_to_f2(v.m, f);
              a[0] = ffma2(f[0], wf[dy * 3 + 1][0], a[0]);
              a[1] = ffma2(f[1], wf[dy * 3 + 1][1], a[1]);
              bf4_to_f2(v.r, f);
              a[0] = ffma2(f[0], wf[dy * 3 + 2][0], a[0]);
              a[1] = ffma2(f[1], wf[dy * 3 + 2][1], a[1]);
            };
            auto emit = [&](int oh, const f2_t (&a)[2], bool pred) {
              float o[4];
              upk2(a[0], o[0], o[1]);
              upk2(a[1], o[2], o[3]);
              sts_u2_pred(a_addr + swz_off<HC>(oh * p.BW + bw, qd), clamp_bf16x2(pack_bf16(o[0], o[1]), lo_d, hi_d),
                          clamp_bf16x2(pack_bf16(o[2], o[3]), lo_d, hi_d), pred);
            };
            if (p.S == 1) {
              // patch row r0 + i feeds outputs r0 + i (dy 0), r0 + i - 1 (dy 1), r0 + i - 2 (dy 2): three accumulators
              // whose roles rotate with period 3 (static under the unroll).  Branch-free: the next row's loads are
              // issued before this row's FMAs, rows past the last one re-read the last (their outputs are never
              // stored), stores are predicated.
              uint32_t ptr = hp_addr + (uint32_t)(r0 * p.PW + bw) * pitch + (uint32_t)(qd * 8);
              const uint32_t last = ptr + (uint32_t)(rows + 1) * rsb;
              const int n_in = rows + 2;
              f2_t acc[3][2];
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                acc[a][0] = b2[0];
                acc[a][1] = b2[1];
              }
              Raw nxt = load3(ptr);
              for (int i0 = 0; i0 < n_in; i0 += 3) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const int i = i0 + j;
                  const Raw cur = nxt;
                  ptr = min(ptr + rsb, last);
                  nxt = load3(ptr);
                  fma_row(acc[j], cur, 0);
                  fma_row(acc[(j + 2) % 3], cur, 1);
                  fma_row(acc[(j + 1) % 3], cur, 2);
                  emit(r0 + i - 2, acc[(j + 1) % 3], i >= 2 && i < n_in);
                  acc[(j + 1) % 3][0] = b2[0];
                  acc[(j + 1) % 3][1] = b2[1];
                }
              }
            } else {
              // output o = patch rows 2o (dy 0), 2o+1, 2o+2; row 2o+2 is also the top row of output o+1
              uint32_t ptr = hp_addr + (uint32_t)(2 * r0 * p.PW + 2 * bw) * pitch + (uint32_t)(qd * 8);
              const uint32_t last = ptr + (uint32_t)(2 * rows) * rsb;
              f2_t a[2] = {b2[0], b2[1]};
              {
                const Raw top = load3(ptr);
                fma_row(a, top, 0);
              }
              ptr = min(ptr + rsb, last);
              Raw n1 = load3(ptr);
              ptr = min(ptr + rsb, last);
              Raw n2 = load3(ptr);
              for (int o = 0; o < rows; ++o) {
                const Raw c1 = n1, c2 = n2;
                ptr = min(ptr + rsb, last);
                n1 = load3(ptr);
                ptr = min(ptr + rsb, last);
                n2 = load3(ptr);
                fma_row(a, c1, 1);
                fma_row(a, c2, 2);
                emit(r0 + o, a, true);
                a[0] = b2[0];
                a[1] = b2[1];
                fma_row(a, c2, 0);
              }
            }
        };
        if (one_item) {
          if (it_active) {
            f2_t wf[9][2], b2[2];
#pragma unroll
            for (int k = 0; k < 9; ++k) bf4_to_f2(wraw[k], wf[k]);
            b2[0] = pk2(braw.x, braw.y);
            b2[1] = pk2(braw.z, braw.w);
            process(it_qd, it_bw, it_r0, it_rows, wf, b2);
            if ((nch > 1 || (p.dbg & 2)) && g + 1 < G) fetch_weights(h + 1 == nch ? 0 : h + 1);   // (one chunk: same weights again)
          }
        } else {
          for (int item = tid; item < n_items; item += MB_DW_THREADS) {
            const int qd = item & (QD - 1);
            const int tt = item / QD;
            const int seg = tt / p.BW;
            const int bw = tt - seg * p.BW;
            const int r0 = seg * p.rs;
            const int rows = min(p.rs, p.BH - r0);
            if (rows <= 0) continue;
            f2_t wf[9][2], b2[2];
            const uint2* wp = p.w_dw + h * QD + qd;
#pragma unroll
            for (int k = 0; k < 9; ++k) bf4_to_f2(__ldg(wp + (size_t)k * hq), wf[k]);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b_dw + h * HC) + qd);
            b2[0] = pk2(bv.x, bv.y);
            b2[1] = pk2(bv.z, bv.w);
            process(qd, bw, r0, rows, wf, b2);
          }
        }
        mb_prof_work<PROF>(t_work);
        if (tid == 0) mb_event<PROF>(g, 1, t_start);
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
  if (PROF && s_mb_prof_on) {
    if (threadIdx.x < 256) g_mb_prof[threadIdx.x] = s_mb_prof[threadIdx.x];
    if (threadIdx.x < 64) g_mb_prof[258 + threadIdx.x] = s_mb_ev[threadIdx.x];
    if (threadIdx.x == 0) {
      g_mb_prof[256] = (unsigned long long)(clock64() - t_start);
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) ++my_tiles;
      g_mb_prof[257] = (unsigned long long)my_tiles * (unsigned long long)p.nch;
    }
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
        // stride 1 runs its rows in groups of 3 (rs + 2 input rows); ~50 issue slots per row step per warp
        const double steps = S == 1 ? 3.0 * ((rs + 2 + 2) / 3) : 2.0 * rs + 1.0;
        const double c = rounds * (steps * 50.0 + 120.0);
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
                   const cuuint64_t* strides, const cuuint32_t* box, int inner_bytes, bool swizzled = true) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle swz = !swizzled ? CU_TENSOR_MAP_SWIZZLE_NONE
                                 : inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
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

extern "C" int ssdsb_mbconv_profile(unsigned long long* out258 /* 322 entries */) {
  SSDSB_REQUIRE(out258, "mbconv_profile: NULL argument");
  SSDSB_CUDA(cudaDeviceSynchronize());
  SSDSB_CUDA(cudaMemcpyFromSymbol(out258, g_mb_prof, sizeof(unsigned long long) * 322));
  return SSDSB_OK;
}

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
  p.dbg = getenv("SSDSB_MB_DEBUG") ? atoi(getenv("SSDSB_MB_DEBUG")) : 0;
  p.prof = getenv("SSDSB_MB_PROF") ? atoi(getenv("SSDSB_MB_PROF")) : 0;      // 1: event stamps, 2: + wait cycles
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
          double best = 1e30;                    // row segmentation of the forced tile: same rule as pick_geometry
          for (int nseg = 1; nseg <= 4 && nseg <= bh; ++nseg) {
            const int rs = (bh + nseg - 1) / nseg;
            const int rounds = (nseg * bw * (hc_try / 4) + MB_DW_THREADS - 1) / MB_DW_THREADS;
            const double steps = S == 1 ? 3.0 * ((rs + 2 + 2) / 3) : 2.0 * rs + 1.0;
            const double c = rounds * (steps * 50.0 + 120.0);
            if (c < best) {
              best = c;
              gm.nseg = nseg;
              gm.rs = rs;
            }
          }
        }
      }
    }
    const int pp_pad = (gm.PP + 15) / 16 * 16;
    const int rowb_x = p.blk_x * 2;
    const int xkb_bytes = pp_pad * rowb_x;
    const int x_bytes = p.xkb * xkb_bytes;
    const int we_bytes = has_expand ? p.xkb * hc_try * rowb_x : 0;
    const int wp_bytes = block_n * hc_try * 2;
    const int hp_pitch = has_expand ? hc_try * 2 + 16 : hc_try * 2;
    const int hp_bytes = (pp_pad * hp_pitch + 1023) / 1024 * 1024;
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
        p.hp_pitch = hp_pitch;
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
    // (without an expand layer the patch lands in hpatch, which the depthwise warps read as dense linear rows)
    CUresult r = encode_tm(encode, &tmX, 4, x, dims, strides, box, inner * 2, has_expand);
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
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_MAX_BYTES);
    if (e != cudaSuccess) return e;
    kernel<<<grid, MB_NT, smem_total, st>>>(tmX, tmWe, tmWp, tmY, tmR, p);
    return cudaSuccess;
  };
  if (p.prof == 2)
    SSDSB_CUDA(HC == 64 ? launch(mbconv_kernel<64, 2>) : launch(mbconv_kernel<32, 2>));
  else if (p.prof == 1)
    SSDSB_CUDA(HC == 64 ? launch(mbconv_kernel<64, 1>) : launch(mbconv_kernel<32, 1>));
  else
    SSDSB_CUDA(HC == 64 ? launch(mbconv_kernel<64, 0>) : launch(mbconv_kernel<32, 0>));
  SSDSB_LAUNCH_CHECK("mbconv_kernel");
  return SSDSB_OK;
}
