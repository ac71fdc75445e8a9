// The 3x3 implicit-GEMM convolution (tcgen05) with conv_init0 fused in (CUDA cores), the layout converter and the
// GEMM self-test / MMA micro-benchmark kernels.
// Reference semantics: model/LFT.py:23-33,65-66 (conv stack) and LFT.py:164-169 (SpaTrans.SAI2Token:
// unfold 3x3 + Linear == zero-padded 3x3 conv 64->128).
#include "host.h"
#include "kernels.cuh"

#include <cstring>

namespace lft {

// ------------------------------------------------------------------------------------------------
// Layout conversion for the stage-level entry points: channels-last [T][C] <-> T32.
__global__ void __launch_bounds__(256) k_layout(const float* __restrict__ in, float* __restrict__ out, long long T,
                                               int C4, int to_t32) {
  pdl_trigger();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4
  if (gid >= T * C4) return;
  long long t;
  int ch;
  if (to_t32) {  // coalesced reads of rows
    t = gid / C4;
    ch = (int)(gid - t * C4);
    *reinterpret_cast<float4*>(out + t32_off(t, ch, C4)) = __ldg(reinterpret_cast<const float4*>(in) + gid);
  } else {
    t = gid / C4;
    ch = (int)(gid - t * C4);
    reinterpret_cast<float4*>(out)[gid] = __ldg(reinterpret_cast<const float4*>(in + t32_off(t, ch, C4)));
  }
}

int launch_layout(Handle* h, const float* in, float* out, long long T, int C, int to_t32, cudaStream_t st) {
  const long long n = T * (C / 4);
  LFT_LAUNCH(h, k_layout, (unsigned)((n + 255) / 256), 256, 0, st, in, out, T, C / 4, to_t32);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LFT_ERR_CUDA, "k_layout launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// 3x3 conv, 64 -> N channels, as an implicit GEMM on tcgen05.
//
// Positions: every view is laid out as P rows of (P+1) positions (the extra one is a zero pad that
// serves as right pad of row y and left pad of row y+1) followed by one zero pad row; view stride
// VS=(P+1)^2.  In this linear space every tap (dy,dx) is the constant row shift dy*(P+1)+dx, so the
// nine taps are nine descriptor offsets into ONE staged copy of the inputs (no im2col).
// A CTA computes 128 consecutive positions; pad positions are computed and discarded (6% at P=32).
//
// Fusion of conv_init0 (LFT.py:23-25,65): with `lr` set, the input window of the FIRST conv of the stack is
// conv_init0(lr) evaluated on the fly (9 taps x 64 channels per staged position, fp32 fma chain over the 9 taps per channel), and the
// residual of the LAST conv (`buffer = conv_init(buffer) + buffer`, LFT.py:66) is recomputed the same way instead of
// being read back -- the [T,64] conv_init0 tensor never exists in HBM.
struct W0Tab { float w[64 * 9]; };  // conv_init0.0.weight [c][tap] (constant bank)

// the 9 taps of view `v` (= b*A*A + u*A + vv) at (y, x), zero padded per view
LFT_DEVINL void conv0_taps(const float* __restrict__ lr, int A, int P, unsigned v, int y, int x, float* t) {
  const unsigned NA = (unsigned)(A * A);
  const unsigned b = v / NA, a = v - b * NA;
  const int u = (int)a / A, vv = (int)a - u * A;
  const int W = A * P;
  const float* img = lr + (long long)b * W * W + (long long)(u * P) * W + vv * P;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      t[ky * 3 + kx] = (yy >= 0 && yy < P && xx >= 0 && xx < P) ? __ldg(img + (long long)yy * W + xx) : 0.f;
    }
}
template <int C0, int NC>
LFT_DEVINL void conv0_channels(const W0Tab& w0, const float* t, float* o) {  // channels C0 .. C0+NC-1 (compile-time indices)
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) s = fmaf(w0.w[(C0 + j) * 9 + k], t[k], s);
    o[j] = s;
  }
}

// conv_stage_window with conv_init0 computed in place of the loads
template <bool BIG>
LFT_DEVINL void conv_stage_window_lr(const float* __restrict__ lr, const W0Tab& w0, int A, uint32_t a_hi, uint32_t a_lo,
                                     long long g0, long long G, long long VS, int P, int tid, bool fp32_mode) {
  constexpr int kConvRows = ConvGeom<BIG>::kRows, kConvOff = ConvGeom<BIG>::kOff;
  const int P1 = P + 1;
  for (int r = tid; r < kConvRows; r += kRowThreads2) {
    const long long g = g0 - kConvOff + r;
    bool inside = false;
    unsigned v = 0;
    int y = 0, x = 0;
    if (g >= 0 && g < G) {
      const unsigned gu = (unsigned)g, vsu = (unsigned)VS;
      v = gu / vsu;
      const int qq = (int)(gu - v * vsu);
      y = qq / P1;
      x = qq - y * P1;
      inside = (y < P && x < P);
    }
    float t[9];
    if (inside) {
      conv0_taps(lr, A, P, v, y, x, t);
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) t[k] = 0.f;  // pad positions: zero feature vector
    }
    auto put = [&](int kc, const float* f) {
      uint4 hi, lo;
      split8(f, hi, lo, fp32_mode);
      st_shared_v4(a_hi + kc * (kConvRows * 16) + r * 16, hi);
      if (fp32_mode) st_shared_v4(a_lo + kc * (kConvRows * 16) + r * 16, lo);
    };
    float f[8];
    conv0_channels<0, 8>(w0, t, f);  put(0, f);
    conv0_channels<8, 8>(w0, t, f);  put(1, f);
    conv0_channels<16, 8>(w0, t, f); put(2, f);
    conv0_channels<24, 8>(w0, t, f); put(3, f);
    conv0_channels<32, 8>(w0, t, f); put(4, f);
    conv0_channels<40, 8>(w0, t, f); put(5, f);
    conv0_channels<48, 8>(w0, t, f); put(6, f);
    conv0_channels<56, 8>(w0, t, f); put(7, f);
  }
}

// Persistent: a CTA (two per SM) walks over the tiles blockIdx.x, blockIdx.x + gridDim.x, ... with TWO accumulators (tile k uses
// TMEM columns [128 (k&1), +128)): once tile k's MMAs are complete the row owners stage tile k+1's window, publish it, and run
// tile k's epilogue (accumulator loads, LeakyReLU / residual, stores) UNDER tile k+1's MMAs.  a_ready and mma_done complete one
// phase per tile (parity k&1).  The accumulator of tile k is overwritten by tile k+2, whose MMAs are gated by a_ready arrivals
// every row owner makes after its epilogue of tile k; the window is overwritten only after mma_done of the tile that read it.
template <int N, bool BIG>
__global__ void __launch_bounds__(kThreads2, 2)
k_conv3x3(const float* __restrict__ in, const uint8_t* __restrict__ wp, float* __restrict__ out,
          const float* __restrict__ res, int V, int P, int passes, int epi, const float* __restrict__ lr,
          const __grid_constant__ W0Tab w0, int A, const uint8_t* __restrict__ wst, int ntiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NST = ConvGeom<BIG>::kNST;
  constexpr int kConvRows = ConvGeom<BIG>::kRows, kConvOff = ConvGeom<BIG>::kOff;
  // fp32 mode streams STACKED slabs: per tap one [128 x 64] B operand (rows 0..63 hi, 64..127 lo), so that A_hi is read once
  // for A_hi*W_hi and A_hi*W_lo (one N = 128 MMA per k step, accumulator columns [0,64) | [64,128)) and A_lo*W_hi adds into
  // [0,64) with an N = 64 MMA on the same slab: 14 KB instead of 18 KB of operands and 112 instead of 144 pipe cycles per k step.
  constexpr uint32_t STAGE = 2 * N * 128;
  static_assert(2 * N <= 128, "one accumulator = 128 TMEM columns");
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t a_hi = s_base + kCtlBytes;
  const uint32_t a_lo = a_hi + kConvRows * 128;
  const uint32_t ring = a_lo + kConvRows * 128;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P1 = P + 1;
  const long long VS = (long long)P1 * P1;
  const long long G = (long long)V * VS;
  const int first = blockIdx.x, step = gridDim.x;
  const int ntl = first < ntiles ? (ntiles - first + step - 1) / step : 0;

  pdl_trigger();
  cta_setup<NST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;
  const bool stacked = passes == 3;

  GemmPhase ph{wp, (uint32_t)N, 9};
  if (warp == kWarpProducer2) {

    RingState<NST> rs;
    for (int k = 0; k < ntl; ++k) {
      if (stacked) {  // one 16 KB slab per tap, consecutive in memory
        for (uint32_t t = 0; t < 9; ++t) {
          mbar_wait(empty0 + 8u * rs.stage, rs.phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(full0 + 8u * rs.stage, STAGE);
            bulk_g2s(ring + rs.stage * STAGE, wst + (size_t)t * STAGE, STAGE, full0 + 8u * rs.stage);
          }
          __syncwarp();
          rs.advance();
        }
      } else {
        ring_produce<NST>(rs, ring, STAGE, full0, empty0, ph, passes);
      }
    }
  } else if (warp == kWarpMma2) {

    RingState<NST> rs;
    auto shift = [P1](uint32_t t) { return ((int)(t / 3) - 1) * P1 + ((int)(t % 3) - 1); };
    for (int k = 0; k < ntl; ++k) {
      const uint32_t acc = tmem + 128u * (uint32_t)(k & 1);
      mbar_wait(a_ready, (uint32_t)(k & 1));
      tc_fence_after();
      if (stacked) {
        const uint32_t id_wide = umma_idesc_bf16(2 * N), id_half = umma_idesc_bf16(N);
        const uint32_t a_lbo = kConvRows * 16, b_lbo = 2 * N * 16;
        const uint32_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;
        const uint32_t ahi0 = umma_desc_lo(a_hi + kConvOff * 16, a_lbo), alo0 = umma_desc_lo(a_lo + kConvOff * 16, a_lbo);
        for (uint32_t t = 0; t < 9; ++t) {
          const uint32_t ah = ahi0 + (uint32_t)shift(t), al = alo0 + (uint32_t)shift(t);
          mbar_wait(full0 + 8u * rs.stage, rs.phase);
          tc_fence_after();
          const uint32_t b0 = umma_desc_lo(ring + rs.stage * STAGE, b_lbo);
          if (elect_one()) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)   // A_hi x [W_hi; W_lo] -> columns [0,64) | [64,128)
              umma_bf16(acc, umma_desc_from(ah + j * a_step), umma_desc_from(b0 + j * b_step), id_wide, (t | j) ? 1u : 0u);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)   // A_lo x W_hi (rows 0..63 of the same slab) -> columns [0,64)
              umma_bf16(acc, umma_desc_from(al + j * a_step), umma_desc_from(b0 + j * b_step), id_half, 1u);
            umma_commit(empty0 + 8u * rs.stage);
          }
          __syncwarp();
          rs.advance();
        }
      } else {
        ring_consume_mma<NST>(rs, ring, STAGE, full0, empty0, ph, passes, a_hi + kConvOff * 16, a_lo + kConvOff * 16,
                              kConvRows * 16, 0, shift, acc, true);
      }
      umma_commit_elected(mma_done);
    }
  } else {
    // ---- stage the input window of tile k: rows r <-> positions g0 - kConvOff + r, lanes <-> rows
    auto stage = [&](int k) {
      const long long g0 = (long long)(first + k * step) * 128;
      if (N == 64 && in == nullptr)
        conv_stage_window_lr<BIG>(lr, w0, A, a_hi, a_lo, g0, G, VS, P, tid, passes == 3);
      else
        conv_stage_window<BIG>(in, a_hi, a_lo, g0, G, VS, P, Region{0, P}, tid, passes == 3);
      fence_proxy_async_smem();
      mbar_arrive(a_ready);
    };
    pdl_wait();  // the input feature map / LR patches are the previous kernel's output (weights are not: the producer runs ahead)
    if (ntl > 0) stage(0);

    // ---- epilogue: row m <-> position, column half q
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    constexpr int HC = N / 2;  // own columns
    for (int k = 0; k < ntl; ++k) {
      mbar_wait(mma_done, (uint32_t)(k & 1));
      tc_fence_after();
      if (k + 1 < ntl) stage(k + 1);  // the window is free; tile k+1's MMAs (other accumulator) run under the epilogue below
      const long long g = (long long)(first + k * step) * 128 + m;
      long long tok = -1;
      float4 r4[HC / 4];
      if (g < G) {
        const unsigned gu = (unsigned)g, vsu = (unsigned)VS;
        const unsigned v = gu / vsu;
        const int qq = (int)(gu - v * vsu);
        const int y = qq / P1, x = qq - y * P1;
        if (y < P && x < P) tok = (long long)((v * P + y) * P + x);
        if (N == 64 && (epi & 2) && res == nullptr && tok >= 0) {  // residual = conv_init0(lr), own 32 channels
          float t[9];
          conv0_taps(lr, A, P, v, y, x, t);
          float* rr = reinterpret_cast<float*>(r4);
          if (q == 0) conv0_channels<0, 32>(w0, t, rr);
          else conv0_channels<32, 32>(w0, t, rr);
        }
      }
      if ((epi & 2) && res != nullptr && tok >= 0) {
#pragma unroll
        for (int i = 0; i < HC / 4; ++i)
          r4[i] = __ldg(reinterpret_cast<const float4*>(res + t32_off(tok, q * (HC / 4) + i, N / 4)));
      }
      const uint32_t trow = tmem + 128u * (uint32_t)(k & 1) + ((uint32_t)((warp & 3) * 32) << 16);
      float v[HC];
#pragma unroll
      for (int c = 0; c < HC / 16; ++c) tmem_ld16_nowait(trow + HC * q + 16 * c, v + 16 * c);
      if (stacked) {  // + A_hi*W_lo from columns [64,128)
        float u[HC];
#pragma unroll
        for (int c = 0; c < HC / 16; ++c) tmem_ld16_nowait(trow + N + HC * q + 16 * c, u + 16 * c);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < HC; ++i) v[i] += u[i];
      } else {
        tmem_wait_ld();
      }
      if (tok >= 0) {
        if (epi & 1) {
#pragma unroll
          for (int i = 0; i < HC; ++i) v[i] = lrelu02(v[i]);
        }
#pragma unroll
        for (int i = 0; i < HC / 4; ++i) {
          float4 o4 = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (epi & 2) { o4.x += r4[i].x; o4.y += r4[i].y; o4.z += r4[i].z; o4.w += r4[i].w; }
          *reinterpret_cast<float4*>(out + t32_off(tok, q * (HC / 4) + i, N / 4)) = o4;
        }
      }
      // pull the window of tile k+2 into L2 while tile k+1's MMAs run (-4.6 %; the same hint in k_spa_embed_qkv, whose
      // staging is already hidden under its Q MMA, changed nothing)
      if (in != nullptr && k + 2 < ntl) conv_prefetch_window<BIG>(in, (long long)(first + (k + 2) * step) * 128, G, VS, P, Region{0, P}, tid);
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
}

// ------------------------------------------------------------------------------------------------
// Self-test: D[128 x N] = A[128 x K] * W[N x K]^T through exactly the machinery the real kernels use
// (chunk-major operands, weight ring, hi/lo passes, TMEM ld/st).  `variant` 1 swaps the LBO/SBO
// descriptor fields (bring-up aid).  aux[128 x 16] returns a tcgen05.st -> tcgen05.ld round trip.
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_selftest(const float* __restrict__ A, int K, const uint8_t* __restrict__ wp, int N, float* __restrict__ D,
                float* __restrict__ aux, int passes, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NST = 3;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t a_lbo = 128 * 16;
  const uint32_t a_hi = s_base + kCtlBytes;
  const uint32_t a_lo = a_hi + (K / 8) * a_lbo;
  const uint32_t ring = a_lo + (K / 8) * a_lbo;
  const uint32_t STAGE = N * 128;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cta_setup<NST>(ctl, warp, lane, 128, 512);
  const uint32_t tmem = ctl->tmem;
  A += (size_t)blockIdx.x * 128 * K;
  D += (size_t)blockIdx.x * 128 * N;
  aux += (size_t)blockIdx.x * 128 * 16;
  GemmPhase ph{wp, (uint32_t)N, (uint32_t)(K / 64)};
  if (warp == kWarpProducer) {

    RingState<NST> rs;
    ring_produce<NST>(rs, ring, STAGE, full0, empty0, ph, passes);
  } else if (warp == kWarpMma) {

    RingState<NST> rs;
    mbar_wait(a_ready, 0);
    tc_fence_after();
    if (variant != 2) {
      ring_consume_mma<NST>(rs, ring, STAGE, full0, empty0, ph, passes, a_hi, a_lo, a_lbo, 8 * a_lbo, NoShift{}, tmem, true);
    } else {
      // TS form (A operand in TMEM): hi at columns [288, 288+K/2), lo at [288+K/2, 288+K)
      const uint32_t idesc = umma_idesc_bf16(N);
      const uint32_t b_lbo = N * 16u, b_step = (2u * b_lbo) >> 4;
      const uint32_t ta_hi = tmem + 288, ta_lo = tmem + 288 + K / 2;
      uint32_t acc = 0;
      for (uint32_t ks = 0; ks < ph.kslabs; ++ks) {
        mbar_wait(full0 + 8u * rs.stage, rs.phase);
        tc_fence_after();
        uint32_t b0 = umma_desc_lo(ring + rs.stage * STAGE, b_lbo);
        if (elect_one()) {
          for (uint32_t j = 0; j < 4; ++j) {
            umma_bf16_ts(tmem, ta_hi + ks * 32 + j * 8, umma_desc_from(b0 + j * b_step), idesc, acc);
            acc = 1;
          }
          if (passes == 3)
            for (uint32_t j = 0; j < 4; ++j)
              umma_bf16_ts(tmem, ta_lo + ks * 32 + j * 8, umma_desc_from(b0 + j * b_step), idesc, 1u);
          umma_commit(empty0 + 8u * rs.stage);
        }
        __syncwarp();
        rs.advance();
        if (passes == 3) {
          mbar_wait(full0 + 8u * rs.stage, rs.phase);
          tc_fence_after();
          b0 = umma_desc_lo(ring + rs.stage * STAGE, b_lbo);
          if (elect_one()) {
            for (uint32_t j = 0; j < 4; ++j)
              umma_bf16_ts(tmem, ta_hi + ks * 32 + j * 8, umma_desc_from(b0 + j * b_step), idesc, 1u);
            umma_commit(empty0 + 8u * rs.stage);
          }
          __syncwarp();
          rs.advance();
        }
      }
    }
    umma_commit_elected(mma_done);
  } else {
    const int m = tid;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    float st[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) st[i] = 2.f * A[(size_t)m * K + i] + 1.f;
    tmem_st16(trow + 256, st);
    tmem_wait_st();
    for (int kc = 0; kc < K / 8; ++kc) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = A[(size_t)m * K + kc * 8 + i];
      uint4 hi, lo;
      split8(x, hi, lo, passes == 3);
      st_shared_v4(a_hi + kc * a_lbo + m * 16, hi);
      st_shared_v4(a_lo + kc * a_lbo + m * 16, lo);
      if (variant == 2) {  // the same packed pairs, 4 columns per 8 elements, into this thread's TMEM lane
        float fh[4] = {__uint_as_float(hi.x), __uint_as_float(hi.y), __uint_as_float(hi.z), __uint_as_float(hi.w)};
        float fl[4] = {__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w)};
        tmem_st4(trow + 288 + 4 * kc, fh);
        tmem_st4(trow + 288 + K / 2 + 4 * kc, fl);
      }
    }
    tmem_wait_st();
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);
    mbar_wait(mma_done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(trow + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[(size_t)m * N + c0 + i] = v[i];
    }
    float v[16];
    tmem_ld16(trow + 256, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) aux[m * 16 + i] = v[i];
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 512);
}

// Micro-benchmark: issue `reps` x (K/16) MMAs of shape 128 x N x 16 back to back on resident operands (no ring, no
// row owners) and report clock64 cycles per MMA.  mode 0: A and B from shared memory; mode 1: A from TMEM.
template <int mode>
__global__ void __launch_bounds__(64, 2) k_mma_bench(int N, int K, int reps, long long* out, int pad_smem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tslot), 256);
  for (int i = threadIdx.x; i < (128 * K * 2 + N * K * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 1) {
    const uint32_t a_lbo = 128 * 16, b_lbo = N * 16;
    const uint32_t a0 = umma_desc_lo(smem_u32(smem), a_lbo), b0 = umma_desc_lo(smem_u32(smem) + 128 * K * 2, b_lbo);
    const uint32_t idesc = umma_idesc_bf16(N);
    const uint32_t sw0 = smem_u32(smem) >> 4;
    (void)sw0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if constexpr (mode >= 5) {
        // the 3x3 conv issue pattern on resident operands: per tap 4 x (A_hi, W_hi), 4 x (A_lo, W_hi), 4 x (A_hi, W_lo); A rows
        // shifted per tap; mode 5: k-chunk stride 201 rows (3216 B), mode 6: 208 rows (3328 B = 26 x 128), mode 7: no shifts
        constexpr uint32_t rows = mode == 6 ? 208u : 201u;
        const uint32_t lbo = rows * 16u, astep = (2u * lbo) >> 4, bstep = (2u * b_lbo) >> 4;
        const uint32_t ahi = umma_desc_lo(smem_u32(smem) + 36 * 16, lbo), alo = umma_desc_lo(smem_u32(smem) + rows * 128 + 36 * 16, lbo);
        const uint32_t bb = umma_desc_lo(smem_u32(smem) + 2 * rows * 128, b_lbo);
        if (elect_one()) {
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            const int sh = mode == 7 ? 0 : (t / 3 - 1) * 33 + (t % 3 - 1);
            const uint32_t ah = ahi + (uint32_t)sh, al = alo + (uint32_t)sh;
            const uint32_t bh = bb + (uint32_t)((2 * t) % 3) * (N * 128u >> 4), bl = bb + (uint32_t)((2 * t + 1) % 3) * (N * 128u >> 4);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) umma_bf16(tmem, umma_desc_from(ah + j * astep), umma_desc_from(bh + j * bstep), idesc, 1u);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) umma_bf16(tmem, umma_desc_from(al + j * astep), umma_desc_from(bh + j * bstep), idesc, 1u);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) umma_bf16(tmem, umma_desc_from(ah + j * astep), umma_desc_from(bl + j * bstep), idesc, 1u);
          }
        }
        __syncwarp();
        continue;
      }
      if (elect_one()) {
        for (int j = 0; j < K / 16; ++j) {
          if (mode == 0)
            umma_bf16(tmem, umma_desc_from(a0 + j * ((2 * a_lbo) >> 4)), umma_desc_from(b0 + j * ((2 * b_lbo) >> 4)), idesc, 1u);
          else if (mode == 1)
            umma_bf16_ts(tmem, tmem + 256 - K / 2 + j * 8, umma_desc_from(b0 + j * ((2 * b_lbo) >> 4)), idesc, 1u);
          else if (mode == 2)  // SS, A start shifted by one 16-byte row (a conv tap with dx = +-1): core matrices straddle 128 B
            umma_bf16(tmem, umma_desc_from(a0 + 1 + j * ((2 * a_lbo) >> 4)), umma_desc_from(b0 + j * ((2 * b_lbo) >> 4)), idesc, 1u);
          else {               // SS, A in the SWIZZLE_128B K-major layout (rows of 128 B, 64-wide k slabs); mode 4: shifted by one row
            constexpr uint32_t sh = mode == 4 ? 1u : 0u;
            const uint32_t a_lo32 = ((sw0 + sh * 8u + (uint32_t)(j >> 2) * 1024u + (uint32_t)(j & 3) * 2u) & 0x3FFFu) | (1u << 16);
            constexpr uint32_t a_hi32 = (1024u >> 4) | (1u << 14) | (sh << 17) | (2u << 29);
            umma_bf16(tmem, ((uint64_t)a_hi32 << 32) | a_lo32, umma_desc_from(b0 + j * ((2 * b_lbo) >> 4)), idesc, 1u);
          }
        }
      }
      if (pad_smem > 0 && (r % pad_smem) == pad_smem - 1 && elect_one()) umma_commit(smem_u32(&bar2));  // extra commits
      __syncwarp();
    }
    umma_commit_elected(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  (void)pad_smem;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int launch_mma_bench(int N, int K, int reps, int mode, int grid, int smem_bytes, long long* host_out) {
  const int commit_every = mode / 16;  // mode = base + 16 * (commit after every n-th group of K/16 MMAs)
  mode %= 16;
  long long* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(long long) * grid));
#define LFT_MMA_BENCH(M)                                                                                  \
  case M:                                                                                                 \
    CUDA_TRY(cudaFuncSetAttribute(k_mma_bench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
    k_mma_bench<M><<<grid, 64, smem_bytes>>>(N, K, reps, d, commit_every);                                          \
    break;
  switch (mode) {
    LFT_MMA_BENCH(0) LFT_MMA_BENCH(1) LFT_MMA_BENCH(2) LFT_MMA_BENCH(3) LFT_MMA_BENCH(4) LFT_MMA_BENCH(5) LFT_MMA_BENCH(6)
    LFT_MMA_BENCH(7)
    default: cudaFree(d); return fail(LFT_ERR_ARG, "mma bench mode 0..7");
  }
#undef LFT_MMA_BENCH
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(host_out, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

// ------------------------------------------------------------------------------------------------ host
template <bool BIG>
constexpr size_t smem_conv64() { return kCtlBytes + 2 * ConvGeom<BIG>::kRows * 128 + ConvGeom<BIG>::kNST * 128 * 128; }

int configure_conv() {
  CUDA_TRY(cudaFuncSetAttribute(k_conv3x3<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_conv64<false>()));
  CUDA_TRY(cudaFuncSetAttribute(k_conv3x3<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_conv64<true>()));
  CUDA_TRY(cudaFuncSetAttribute(k_gemm_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return 0;
}

// in == nullptr: the input is conv_init0(lr) computed on the fly; (epi & 2) with res == nullptr: so is the residual.
int launch_conv3x3(Handle* h, int N, const float* in, const uint8_t* wp, const uint8_t* wst, float* out, const float* res,
                   int V, int P, int epi, const float* lr, cudaStream_t st) {
  const long long G = (long long)V * (P + 1) * (P + 1);
  const unsigned ntiles = (unsigned)((G + 127) / 128);
  const unsigned grid = ntiles < 2u * h->num_sms ? ntiles : 2u * h->num_sms;  // persistent: two CTAs per SM
  if ((in == nullptr || ((epi & 2) && res == nullptr)) && (N != 64 || lr == nullptr))
    return fail(LFT_ERR_ARG, "launch_conv3x3: fused conv_init0 needs N == 64 and the LR mosaic");
  W0Tab w0;
  memcpy(w0.w, h->w_conv0_host.data(), sizeof(w0.w));
  Scope sc(h, K_CONV64, st, (long long)V * P * P);
  if (N != 64) return fail(LFT_ERR_ARG, "launch_conv3x3: only the 64 -> 64 conv stack uses this kernel (the 64 -> 128 token embedding is part of k_spa_embed_qkv)");
  if (P <= ConvGeom<false>::kMaxP) {
    auto kern = k_conv3x3<64, false>;
    LFT_LAUNCH(h, kern, grid, kThreads2, smem_conv64<false>(), st, in, wp, out, res, V, P, h->passes(), epi, lr, w0, h->cfg.ang_res, wst, (int)ntiles);
  } else {
    auto kern = k_conv3x3<64, true>;
    LFT_LAUNCH(h, kern, grid, kThreads2, smem_conv64<true>(), st, in, wp, out, res, V, P, h->passes(), epi, lr, w0, h->cfg.ang_res, wst, (int)ntiles);
  }
  return sc.finish();
}

int launch_selftest(const float* dA, int K, const uint8_t* dW, int N, float* dD, float* dX, int M, int passes,
                    int variant) {
  const size_t smem = kCtlBytes + 2 * (size_t)(K / 8) * 128 * 16 + 3 * (size_t)N * 128;
  k_gemm_selftest<<<M / 128, kThreads, smem>>>(dA, K, dW, N, dD, dX, passes, variant);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace lft
