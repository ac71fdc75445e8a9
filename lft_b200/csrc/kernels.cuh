// Kernel-side shared declarations: CTA roles, control block, setup/teardown, kernel prototypes.
#pragma once
#include "common.cuh"
#include "host.h"

namespace lft {

// CTA roles for the tcgen05 kernels: warps 0-3 own one accumulator row each (TMEM lane = thread id),
// warp 4 streams weights (one elected lane), warp 5 allocates TMEM and issues the MMAs (one lane).
constexpr int kThreads = 192;
constexpr int kWarpProducer = 4;
constexpr int kWarpMma = 5;
constexpr int kCtlBytes = 256;
// conv staging: rows <-> positions g0-kOff .. g0-kOff+kRows-1; a tap shifts by at most P + 2 positions.
//   BIG = false: patches up to 32 x 32 (|shift| <= 34), three ring stages - the default tiling of test.py
//   BIG = true : patches up to 64 x 64 (|shift| <= 66): the window needs 261 rows (66.8 KB for hi + lo), so the weight ring
//                shrinks to two stages to keep two CTAs per SM (--patch_size_for_test 64)
template <bool BIG>
struct ConvGeom {
  static constexpr int kRows = BIG ? 261 : 201;  // odd: k-chunk planes start in different banks
  static constexpr int kOff = BIG ? 66 : 36;
  static constexpr int kNST = BIG ? 2 : 3;
  static constexpr int kMaxP = BIG ? 64 : 32;
};

struct Ctl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t a_ready;
  uint64_t mma_done;
  uint64_t aux[4];
  uint32_t tmem;
};
static_assert(sizeof(Ctl) <= kCtlBytes, "control block too large");

// 2-threads-per-row layout: warps 0-7 are row owners (row m = 32*(warp&3)+lane, column half q = warp>>2;
// warps w and w+4 share a TMEM lane quarter), warp 8 streams weights, warp 9 allocates TMEM / issues MMAs.
constexpr int kThreads2 = 320;
constexpr int kRowThreads2 = 256;
constexpr int kWarpProducer2 = 8;
constexpr int kWarpMma2 = 9;

template <int NST>
LFT_DEVINL void cta_setup(Ctl* ctl, int warp, int lane, uint32_t a_ready_count, uint32_t tmem_cols,
                          int mma_warp = kWarpMma) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), 1);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->a_ready), a_ready_count);
    mbar_init(smem_u32(&ctl->mma_done), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&ctl->aux[i]), a_ready_count);
    mbar_fence_init();
  }
  if (warp == mma_warp) tmem_alloc(smem_u32(&ctl->tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

LFT_DEVINL void cta_teardown(Ctl* ctl, int warp, uint32_t tmem_cols, int mma_warp = kWarpMma) {
  tc_fence_before();
  __syncthreads();
  if (warp == mma_warp) {
    tc_fence_after();
    tmem_dealloc(ctl->tmem, tmem_cols);
  }
}

// LN-fold epilogue constants travel as __grid_constant__ kernel parameters (constant bank, uniform reads).
struct Tab512 { float v[512]; };

// Optional phase timeline of the middle CTA (debug aid, compiled in with -DLFT_TIMELINE): slot s of row
// warp 0 lane 0 -> g_tl[0][s], of the MMA thread -> g_tl[1][s].
#ifdef LFT_TIMELINE
static __device__ long long g_tl[2][32];
#define LFT_TL(s)                                                                                   \
  do {                                                                                              \
    if (blockIdx.x == gridDim.x / 2 && (threadIdx.x == 0 || threadIdx.x == 32 * kWarpMma2))          \
      g_tl[threadIdx.x == 0 ? 0 : 1][s] = clock64();                                                \
  } while (0)
static __device__ long long g_tl2[2][32];
#define LFT_TL2(s)                                                                                  \
  do {                                                                                              \
    if (blockIdx.x == gridDim.x / 2 && (threadIdx.x == 0 || threadIdx.x == 32 * kWarpMma2))          \
      g_tl2[threadIdx.x == 0 ? 0 : 1][s] = clock64();                                               \
  } while (0)
// per-row-warp marks (lane 0 of every row warp): which = 0/1 selects the table half, slot 22 + warp
#define LFT_TL2W(which, warp)                                                                       \
  do {                                                                                              \
    if (blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 0 && (warp) < 8)                       \
      g_tl2[which][22 + (warp)] = clock64();                                                        \
  } while (0)
#else
#define LFT_TL(s) do {} while (0)
#define LFT_TL2(s) do {} while (0)
#define LFT_TL2W(which, warp) do {} while (0)
#endif

// "T32" activation layout for [T, C] fp32 tensors: blocks of 32 consecutive tokens, inside a block the
// 16-byte channel chunks are the slow axis:  off(t, c) = (((t>>5)*(C/4) + c/4) * 32 + (t&31)) * 4 + c%4.
// A warp whose lanes own 32 consecutive tokens reads/writes one chunk as 512 contiguous bytes.
LFT_DEVINL long long t32_off(long long t, int chunk, int C4) {
  return ((((t >> 5) * C4 + chunk) << 5) + (t & 31)) << 2;
}

// Stage the 64-channel input window of a 3x3 conv tile: smem rows r <-> padded positions g0-kConvOff+r,
// one lane per row (T32 source: a warp reads 512 contiguous bytes per chunk), bf16 hi/lo, chunk-major.
// The position space covers the region `e` of every view: rows of e.rn + 1 positions (+ one pad row), VS = (e.rn + 1)^2;
// position (yy, xx) <-> pixel (e.r0 + yy, e.r0 + xx).
template <bool BIG>
LFT_DEVINL void conv_stage_window(const float* __restrict__ in, uint32_t a_hi, uint32_t a_lo, long long g0, long long G,
                                  long long VS, int P, Region e, int tid, bool fp32_mode) {
  constexpr int kConvRows = ConvGeom<BIG>::kRows, kConvOff = ConvGeom<BIG>::kOff;
  const int P1 = e.rn + 1;
  for (int r = tid; r < kConvRows; r += kRowThreads2) {
    const long long g = g0 - kConvOff + r;
    long long tok = -1;
    if (g >= 0 && g < G) {  // G < 2^31 (host-checked): 32-bit index math
      const unsigned gu = (unsigned)g, vsu = (unsigned)VS;
      const unsigned v = gu / vsu;
      const int qq = (int)(gu - v * vsu);
      const int y = qq / P1, x = qq - y * P1;
      if (y < e.rn && x < e.rn) tok = (long long)((v * P + e.r0 + y) * P + e.r0 + x);
    }
    float4 f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      f[i] = tok >= 0 ? __ldg(reinterpret_cast<const float4*>(in + t32_off(tok, i, 16))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      uint4 hi, lo;
      split8(reinterpret_cast<const float*>(&f[2 * kc]), hi, lo, fp32_mode);
      st_shared_v4(a_hi + kc * (kConvRows * 16) + r * 16, hi);
      if (fp32_mode) st_shared_v4(a_lo + kc * (kConvRows * 16) + r * 16, lo);
    }
  }
}

// L2 prefetch of the rows conv_stage_window will read for the tile at g0 (same row <-> thread mapping)
template <bool BIG>
LFT_DEVINL void conv_prefetch_window(const float* __restrict__ in, long long g0, long long G, long long VS, int P, Region e,
                                     int tid) {
  constexpr int kConvRows = ConvGeom<BIG>::kRows, kConvOff = ConvGeom<BIG>::kOff;
  const int P1 = e.rn + 1;
  for (int r = tid; r < kConvRows; r += kRowThreads2) {
    const long long g = g0 - kConvOff + r;
    if (g >= 0 && g < G) {
      const unsigned gu = (unsigned)g, vsu = (unsigned)VS;
      const unsigned v = gu / vsu;
      const int qq = (int)(gu - v * vsu);
      const int y = qq / P1, x = qq - y * P1;
      if (y < e.rn && x < e.rn) {
        const long long tok = (long long)((v * P + e.r0 + y) * P + e.r0 + x);
#pragma unroll
        for (int i = 0; i < 16; ++i) prefetch_l2(in + t32_off(tok, i, 16));
      }
    }
  }
}

// LayerNorm statistics of a row whose two halves live in two partner threads (same lane, warps w / w+4).
// Each thread reduces its own H values (two-pass, in registers), publishes (mean_q, M2_q) through 4 TMEM
// columns it owns, and combines with its partner's pair (Chan's parallel variance).  eps = 1e-5.
template <int H>
LFT_DEVINL void pair_ln_stats(const float* y, uint32_t mbox_own, uint32_t mbox_partner, int bar_id, float& mean,
                              float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < H; ++i) s += y[i];
  const float mq = s * (1.f / H);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < H; ++i) { const float d = y[i] - mq; m2 = fmaf(d, d, m2); }
  float v[4] = {mq, m2, 0.f, 0.f};
  tmem_st4(mbox_own, v);
  tmem_wait_st();
  tc_fence_before();
  pair_bar_sync(bar_id - 1);
  tc_fence_after();
  float o[4];
  tmem_ld4(mbox_partner, o);
  mean = 0.5f * (mq + o[0]);
  const float dm = mq - o[0];
  const float M2 = m2 + o[1] + (0.5f * H) * dm * dm;
  rstd = rsqrtf(M2 * (1.f / (2 * H)) + 1e-5f);
}

}  // namespace lft
