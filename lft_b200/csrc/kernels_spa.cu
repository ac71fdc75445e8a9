// SpaTrans (model/LFT.py:118-191):
//   k_spa_embed_qkv : tok = conv3x3(feat, MLP.weight) (== unfold 3x3 + Linear, LFT.py:164-169) as an implicit GEMM,
//                     then Q = LN(tok+PE_s) Wq^T, K = LN(tok+PE_s) Wk^T, V = tok Wv^T  (LFT.py:180-186), all from ONE
//                     A operand z = tok + PE_s: LayerNorm is folded into the projection epilogue,
//                       LN(z) W^T = rstd (z W'^T - mean u) + c,   W' = W diag(gamma), u = W' 1, c = W beta,
//                     and V = z Wv^T - PE_s Wv^T (constant table).
//   k_spa_attn      : per head (hd=16) softmax over the clamped 5x5 window (<=25 keys) -- the finite entries
//                     of gen_mask (LFT.py:147-162) -- never materialising the [hw,hw] mask     (CUDA cores)
//   k_spa_ffn       : Y1 = tok + O Wo^T; Y2 = Y1 + W2 relu(W1 LN2(Y1)); out = Y2 Wlin^T (1x1x1 conv 128->64)
//                     (+ the global residual of LFT.py:76 on the last block); LN2 folded the same way.
// Q/K/V/O use a planar head-major layout [view][head][y][j][x][4] (channel = head*16 + j*4 + e) so that
// both the row-owner threads of the GEMM kernels and the x-major threads of the window attention
// read/write 16-byte pieces that are contiguous across a warp.
// tcgen05 kernels: 2 threads per accumulator row (column halves), 8 row warps + producer + MMA warp.
#include "host.h"
#include "kernels.cuh"

#include <cstdlib>
#include <cstring>

namespace lft {

constexpr int kSpaNST = 3;
constexpr uint32_t kSpaStage = 128 * 128;
constexpr size_t kSmemSpa = kCtlBytes + 65536 + kSpaNST * kSpaStage;
// k_spa_embed_qkv<BIG>: conv window (hi + lo) followed by the weight ring; BIG (patches up to 64 x 64): 66.8 KB + two stages
template <bool BIG>
constexpr uint32_t embed_window_bytes() { return BIG ? 2u * ConvGeom<true>::kRows * 128u : 65536u; }
template <bool BIG>
constexpr size_t smem_embed() { return kCtlBytes + embed_window_bytes<BIG>() + ConvGeom<BIG>::kNST * kSpaStage; }
constexpr uint32_t kLbo = 128 * 16;
#ifdef LFT_X_FFN_NOLOAD
constexpr bool kFfnNoLoad = true;   // timing experiment: k_spa_ffn2 reads neither O nor tok
#else
constexpr bool kFfnNoLoad = false;
#endif
#ifdef LFT_X_NOQKV
constexpr bool kNoQKV = true;   // timing experiment: no Q / K / V stores
#else
constexpr bool kNoQKV = false;
#endif

LFT_DEVINL long long planar_off(long long v, int head, int y, int j, int x, int P) {
  return ((((v * 8 + head) * P + y) * 4 + j) * (long long)P + x) * 4;
}

// write 16 accumulator columns (= one head) of one token into the planar layout
LFT_DEVINL void planar_store16(float* base, long long v, int head, int y, int x, int P, const float* d) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_stream_v4(base + planar_off(v, head, y, j, x, P), make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]));
}

// The window attention runs on warp-level tensor-core MMAs (k_spa_attn_mma), so Q / K / V leave k_spa_embed_qkv as bf16 hi / lo
// pairs in the SAME planar layout (and the same bytes as fp32): the four 16-byte pieces of a (token, head) are
//   piece 0: hi of dims 0..7 | piece 1: hi of dims 8..15 | piece 2: lo of dims 0..7 | piece 3: lo of dims 8..15
// (x = hi + lo to 2^-16, the accuracy class of every three-pass product here).  bf16 mode: hi pieces only, two pieces per (y, x).
LFT_DEVINL void st_stream_v4u(float* p, const uint4& v) {
  st_stream_v4(p, make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)));
}
// (bf16 mode keeps no room for the lo pieces: two pieces per (y, x), rows of half the size)
LFT_DEVINL long long planar_off_np(long long v, int head, int y, int j, int x, int P, int NP) {
  return ((((v * 8 + head) * P + y) * NP + j) * (long long)P + x) * 4;
}
LFT_DEVINL void planar_store16_split(float* base, long long v, int head, int y, int x, int P, const float* d, bool fp32_mode) {
  uint4 h0, l0, h1, l1;
  split8(d, h0, l0, fp32_mode);
  split8(d + 8, h1, l1, fp32_mode);
  const int NP = fp32_mode ? 4 : 2;
  st_stream_v4u(base + planar_off_np(v, head, y, 0, x, P, NP), h0);
  st_stream_v4u(base + planar_off_np(v, head, y, 1, x, P, NP), h1);
  if (fp32_mode) {
    st_stream_v4u(base + planar_off_np(v, head, y, 2, x, P, NP), l0);
    st_stream_v4u(base + planar_off_np(v, head, y, 3, x, P, NP), l1);
  }
}
#ifdef LFT_ATTN_V1   // the CUDA-core window attention of rounds 1-2 (fp32 Q / K / V planes)
constexpr bool kAttnMma = false;
#else
constexpr bool kAttnMma = true;
#endif
// O between k_spa_attn_mma and k_spa_ffn2: operand tiles (default) or fp32 planar (-DLFT_O_PLANAR, and with the V1 kernels)
#if defined(LFT_ATTN_V1) || defined(LFT_FFN_V1) || defined(LFT_O_PLANAR)
constexpr int kOTile = 0;
#else
constexpr int kOTile = 1;
#endif

// split 16 fp32 values into two k-chunks (kc0, kc0+1) of the K=128 A operand (hi at A, lo at A+32K)
LFT_DEVINL void a_store16(uint32_t A, int kc0, int m, const float* x, bool fp32_mode) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 hi, lo;
    split8(x + 8 * j, hi, lo, fp32_mode);
    st_shared_v4(A + (kc0 + j) * kLbo + m * 16, hi);
    if (fp32_mode) st_shared_v4(A + 32768 + (kc0 + j) * kLbo + m * 16, lo);
  }
}

// ------------------------------------------------------------------------------------------------
// k_spa_embed_qkv is persistent: a CTA (two per SM) walks over the tiles blockIdx.x, blockIdx.x + gridDim.x, ... keeping
// its TMEM / barriers, and the row owners stage the NEXT tile's conv window into shared memory while the tensor core works
// on Q of the current one (z = tok + PE_s lives in TMEM, TS form - no smem operand reads for A - so the staging area is
// free as soon as the conv MMAs are done).  Accumulators are drained into registers and released before the epilogue
// arithmetic and stores, which then run under the next MMA.  Set-up / tear-down / window fill leave the per-tile chain.  Barriers: F = aux[0] (256) window of the next tile staged; a_ready (256): z ready, Q / K / V drained
// (4 arrivals per tile, each separated from the next by a wait on an MMA that needed the previous phase complete);
// mma_done: conv, Q, K, V (4 commits per tile).
template <bool BIG>
__global__ void __launch_bounds__(kThreads2, 2)
k_spa_embed_qkv(const float* __restrict__ feat, const uint8_t* __restrict__ wmlp, const float* __restrict__ pe,
                  const float* __restrict__ pev, const __grid_constant__ Tab512 tab, const uint8_t* __restrict__ wq,
                  const uint8_t* __restrict__ wk, const uint8_t* __restrict__ wv, float* __restrict__ tok,
                  float* __restrict__ Q, float* __restrict__ K, float* __restrict__ Vv, int V, int P, int passes,
                  int ntiles, Region e, Region o, Region oq) {
  // e: the pixels the position space covers (conv inputs; zero padding outside it), o: the pixels whose tok / Q / K / V are
  // stored.  Full view: e = o = {0, P}.  On the light-field path e is the region the conv inputs are valid on and o = e
  // shrunk by one pixel wherever e's border is not the view border (there the zero padding is not the true neighbour).
  // oq (inside o): the query pixels - tok and Q are only ever read there (K / V also serve as the neighbours' keys).
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NST = ConvGeom<BIG>::kNST;
  constexpr int kConvRows = ConvGeom<BIG>::kRows, kConvOff = ConvGeom<BIG>::kOff;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t U = smem_u32(smem) + kCtlBytes;
  const uint32_t c_hi = U, c_lo = U + kConvRows * 128;
  const uint32_t ring = U + embed_window_bytes<BIG>();
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const uint32_t f_ready = smem_u32(&ctl->aux[0]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P1 = e.rn + 1;
  const long long VS = (long long)P1 * P1;
  const long long G = (long long)V * VS;
  pdl_trigger();
  cta_setup<NST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_c{wmlp, 128, 9}, g_q{wq, 128, 2}, g_k{wk, 128, 2}, g_v{wv, 128, 2};
  const int first = blockIdx.x, step = gridDim.x;
  const int ntl = first < ntiles ? (ntiles - first + step - 1) / step : 0;

  if (warp == kWarpProducer2) {
    RingState<NST> rs;
    for (int k = 0; k < ntl; ++k) {
      ring_produce<NST>(rs, ring, kSpaStage, full0, empty0, g_c, passes);
#ifndef LFT_X_EMBED_ONLY   // (defined: timing experiment - token embedding only, no Q / K / V projections at all)
      ring_produce<NST>(rs, ring, kSpaStage, full0, empty0, g_q, passes);
      ring_produce<NST>(rs, ring, kSpaStage, full0, empty0, g_k, passes);
      ring_produce<NST>(rs, ring, kSpaStage, full0, empty0, g_v, passes);
#endif
    }
  } else if (warp == kWarpMma2) {
    RingState<NST> rs;
    uint32_t rpar = 0;
    auto shift = [P1](uint32_t t) { return ((int)(t / 3) - 1) * P1 + ((int)(t % 3) - 1); };
    const uint32_t ta_hi = tmem + 128, ta_lo = tmem + 192;
    for (int k = 0; k < ntl; ++k) {
      LFT_TL2(10);
      mbar_wait(f_ready, k & 1);
      LFT_TL2(11);
      if (k > 0) { mbar_wait(a_ready, rpar); rpar ^= 1; }  // V of the previous tile drained: D[0,128) is free
      tc_fence_after();
      LFT_TL2(12);
      ring_consume_mma<NST>(rs, ring, kSpaStage, full0, empty0, g_c, passes, c_hi + kConvOff * 16,
                                c_lo + kConvOff * 16, kConvRows * 16, 0, shift, tmem, true);
      umma_commit_elected(mma_done);
      LFT_TL2(13);
#ifdef LFT_X_EMBED_ONLY
      continue;
#endif
      mbar_wait(a_ready, rpar); rpar ^= 1;
      tc_fence_after();
      LFT_TL2(14);
      ring_consume_mma_ts<NST>(rs, ring, kSpaStage, full0, empty0, g_q, passes, ta_hi, ta_lo, tmem + 0, true);
      umma_commit_elected(mma_done);
      LFT_TL2(15);
      mbar_wait(a_ready, rpar); rpar ^= 1;
      tc_fence_after();
      LFT_TL2(16);
      ring_consume_mma_ts<NST>(rs, ring, kSpaStage, full0, empty0, g_k, passes, ta_hi, ta_lo, tmem + 0, true);
      umma_commit_elected(mma_done);
      LFT_TL2(17);
      mbar_wait(a_ready, rpar); rpar ^= 1;
      tc_fence_after();
      LFT_TL2(18);
      ring_consume_mma_ts<NST>(rs, ring, kSpaStage, full0, empty0, g_v, passes, ta_hi, ta_lo, tmem + 0, true);
      umma_commit_elected(mma_done);
      LFT_TL2(19);
    }
  } else {
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int PP = P * P;
    const float4* tab4 = reinterpret_cast<const float4*>(tab.v);  // [u_q | u_k | c_q | c_k] x 128 (constant bank)
    uint32_t mpar = 0;
    auto await = [&]() {
      mbar_wait(mma_done, mpar);
      mpar ^= 1;
      tc_fence_after();
    };
    auto stage = [&](int k) {
      conv_stage_window<BIG>(feat, c_hi, c_lo, (long long)(first + k * step) * 128, G, VS, P, e, tid, passes == 3);
      fence_proxy_async_smem();
      mbar_arrive(f_ready);
    };
    pdl_wait();  // `feat` is the previous kernel's output
#ifdef LFT_X_EMBED_STAGGER   // experiment: CTAs start a fraction of a tile apart, so that their store bursts do not coincide chip-wide
    {
      const long long t0 = clock64(), wait = (long long)(blockIdx.x % LFT_X_EMBED_STAGGER_N) * LFT_X_EMBED_STAGGER;
      while (clock64() - t0 < wait) {}
    }
#endif
    if (ntl > 0) stage(0);
    for (int k = 0; k < ntl; ++k) {
      LFT_TL2(0);
      const long long g = (long long)(first + k * step) * 128 + m;
      bool ok = false;
      long long v = 0;
      int y = 0, x = 0;
      if (g < G) {
        const unsigned gu = (unsigned)g, vsu = (unsigned)VS;
        v = gu / vsu;
        const int qq = (int)(gu - (unsigned)v * vsu);
        const int yy = qq / P1, xx = qq - yy * P1;
        y = e.r0 + yy;
        x = e.r0 + xx;
        ok = yy < e.rn && xx < e.rn && (unsigned)(y - o.r0) < (unsigned)o.rn && (unsigned)(x - o.r0) < (unsigned)o.rn;
      }
      bool okq = ok && (unsigned)(y - oq.r0) < (unsigned)oq.rn && (unsigned)(x - oq.r0) < (unsigned)oq.rn;
      if (!ok) { v = 0; y = 0; x = 0; }
#ifdef LFT_EXPERIMENT_NOSTORE
      ok = okq = false;  // timing experiment: no global stores at all
#endif
#ifdef LFT_EXPERIMENT_SMALLSTORE
      v = v & 7;   // timing experiment (wrong results): every store lands in the first 8 views' planes (L2-resident, 5 MB)
#endif
      const int p = y * P + x;
      const long long token = (v * P + y) * P + x;   // (SMALLSTORE experiment: v was folded above, PE index p is unaffected)
      float mean, rstd;
      {
        float z[64];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(pe) + (long long)(16 * q + i) * PP + p);
          z[4 * i] = b.x; z[4 * i + 1] = b.y; z[4 * i + 2] = b.z; z[4 * i + 3] = b.w;
        }
        await();  // conv
        LFT_TL2(1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[16];
          tmem_ld16(trow + 64 * q + 16 * c, t);
#if !defined(LFT_X_NOTOK) && !defined(LFT_X_ZTOK)
          if (okq) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_stream_v4(tok + t32_off(token, 16 * q + 4 * c + i, 32),
                           make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]));
          }
#endif
#pragma unroll
          for (int i = 0; i < 16; ++i) z[16 * c + i] += t[i];
          a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, z + 16 * c, passes == 3);
          LFT_TL2(22 + c);
        }
        pair_ln_stats<64>(z, trow + 64 * q, trow + 64 * (1 - q), 1 + (warp & 3), mean, rstd);
        LFT_TL2(26);
#if defined(LFT_X_ZTOK)   // timing experiment (ffn then reads z instead of tok): the token store leaves the z -> Q critical path
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(a_ready);
#if LFT_X_ZTOK == 2
        if (k + 1 < ntl) stage(k + 1);
#endif
        if (ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            st_stream_v4(tok + t32_off(token, 16 * q + i, 32), make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]));
        }
#endif
      }
#if !defined(LFT_X_ZTOK)
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(a_ready);             // z ready
#endif
      LFT_TL2(2);
#if !defined(LFT_X_ZTOK) || LFT_X_ZTOK != 2
      if (k + 1 < ntl) stage(k + 1);    // the staging area is free (conv MMAs of this tile are complete)
#endif
      LFT_TL2(3);
#ifdef LFT_X_EMBED_ONLY
      continue;
#endif
      const float mr = mean * rstd;
      await();  // Q
      LFT_TL2(4);
      {
        float dd[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, dd + 16 * c);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(a_ready);           // Q drained (values are in registers)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* d = dd + 16 * c;
          const int col = 64 * q + 16 * c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 uv = tab4[col / 4 + j], cv = tab4[64 + col / 4 + j];
            d[4 * j] = fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x));
            d[4 * j + 1] = fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y));
            d[4 * j + 2] = fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z));
            d[4 * j + 3] = fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w));
          }
          if (okq && !kNoQKV) {
            if (kAttnMma) planar_store16_split(Q, v, 4 * q + c, y, x, P, d, passes == 3);
            else planar_store16(Q, v, 4 * q + c, y, x, P, d);
          }
        }
      }
      LFT_TL2(5);
      await();  // K
      LFT_TL2(6);
      {
        float dd[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, dd + 16 * c);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(a_ready);           // K drained
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* d = dd + 16 * c;
          const int col = 64 * q + 16 * c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 uv = tab4[32 + col / 4 + j], cv = tab4[96 + col / 4 + j];
            d[4 * j] = fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x));
            d[4 * j + 1] = fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y));
            d[4 * j + 2] = fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z));
            d[4 * j + 3] = fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w));
          }
          if (ok && !kNoQKV) {
            if (kAttnMma) planar_store16_split(K, v, 4 * q + c, y, x, P, d, passes == 3);
            else planar_store16(K, v, 4 * q + c, y, x, P, d);
          }
        }
      }
      LFT_TL2(7);
      await();  // V
      LFT_TL2(8);
      {
        float dd[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, dd + 16 * c);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(a_ready);           // V drained: the next tile's conv may overwrite D
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* d = dd + 16 * c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 pv = __ldg(reinterpret_cast<const float4*>(pev) + (long long)(16 * q + 4 * c + j) * PP + p);
            d[4 * j] -= pv.x; d[4 * j + 1] -= pv.y; d[4 * j + 2] -= pv.z; d[4 * j + 3] -= pv.w;
          }
          if (ok && !kNoQKV) {
            if (kAttnMma) planar_store16_split(Vv, v, 4 * q + c, y, x, P, d, passes == 3);
            else planar_store16(Vv, v, 4 * q + c, y, x, P, d);
          }
        }
      }
      LFT_TL2(9);
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
}

#ifndef LFT_ATTN_RB
#define LFT_ATTN_RB 8
#endif
constexpr int kAttnRB = LFT_ATTN_RB;       // query rows per CTA (A/B: 16 stages 20 key rows for 16 instead of 12 for 8)
#ifdef LFT_ATTN_V1   // the CUDA-core window attention of round 1 is only compiled into -DLFT_ATTN_V1 builds (A/B timing)
// ------------------------------------------------------------------------------------------------
// Window attention. One thread = one head of TWO vertically adjacent queries (y0, x), (y0+1, x): the
// 6 x 5 keys their windows cover are read once (30 instead of 50 key reads).
LFT_DEVINL float dot16(const f32x2* q, const ulonglong2& a, const ulonglong2& b, const ulonglong2& c,
                        const ulonglong2& d) {
  f32x2 s0 = mul2(q[0], a.x), s1 = mul2(q[1], a.y);
  s0 = fma2(q[2], b.x, s0); s1 = fma2(q[3], b.y, s1);
  s0 = fma2(q[4], c.x, s0); s1 = fma2(q[5], c.y, s1);
  s0 = fma2(q[6], d.x, s0); s1 = fma2(q[7], d.y, s1);
  return hsum2(add2(s0, s1));
}
LFT_DEVINL void axpy16(f32x2* o, float p, const ulonglong2& a, const ulonglong2& b, const ulonglong2& c,
                       const ulonglong2& d) {
  const f32x2 pp = pack2(p, p);
  o[0] = fma2(pp, a.x, o[0]); o[1] = fma2(pp, a.y, o[1]); o[2] = fma2(pp, b.x, o[2]); o[3] = fma2(pp, b.y, o[3]);
  o[4] = fma2(pp, c.x, o[4]); o[5] = fma2(pp, c.y, o[5]); o[6] = fma2(pp, d.x, o[6]); o[7] = fma2(pp, d.y, o[7]);
}

// CTA = (view, head, block of kAttnRB query rows): the K and V planes of rows [r0-2, r0+RB+2) are contiguous
// in the planar layout and are staged in shared memory with two bulk copies (TMA engine).  Softmax is
// evaluated online, one key row at a time (5 scores per query live at once).
constexpr int kAttnThreads = kAttnRB * 16;  // 32 x-lanes x RB/2 row pairs; wider rows (P > 32) take several passes
constexpr size_t smem_attn(int P) { return 2 * (size_t)(kAttnRB + 4) * 4 * P * 16 + 16; }  // K and V rows [r0-2, r0+RB+2)

__global__ void __launch_bounds__(kAttnThreads)
k_spa_attn(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ Vv,
           float* __restrict__ O, int P, Region qr) {
  // qr: the query pixels [qr.r0, qr.r0 + qr.rn)^2 of every view (full view: {0, P}); keys come from the 5 x 5 window clamped
  // to the VIEW, so K / V must be valid on qr grown by two pixels.
  extern __shared__ __align__(128) uint8_t smem[];
  const int nblk = (qr.rn + kAttnRB - 1) / kAttnRB;
  const int rb = blockIdx.x % nblk;
  const int head = (blockIdx.x / nblk) & 7;
  const long long v = blockIdx.x / (nblk * 8);
  const int r0 = qr.r0 + rb * kAttnRB;
  const int rend = qr.r0 + qr.rn;                                  // one past the last query row / column
  const int ys = max(r0 - 2, 0), ye = min(min(r0 + kAttnRB, rend) + 2, P);   // staged key rows [ys, ye)
  const uint32_t rowbytes = (uint32_t)P * 64;                      // one (y) plane: 4 pieces x P x 16 B
  const uint32_t nbytes = (uint32_t)(ye - ys) * rowbytes;
  const uint32_t bar = smem_u32(smem);
  uint8_t* ks = smem + 16;
  uint8_t* vs = ks + (kAttnRB + 4) * rowbytes;
  pdl_trigger();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();  // Q / K / V are the previous kernel's output
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 2 * nbytes);
    const long long src = planar_off(v, head, ys, 0, 0, P);
    bulk_g2s(smem_u32(ks), K + src, nbytes, bar);
    bulk_g2s(smem_u32(vs), Vv + src, nbytes, bar);
  }
  const long long rowstride = (long long)P * 16;  // floats between consecutive y
  const int jstride = P * 4;                      // floats between the 4 pieces of one (y, x)
  const float qs = 0.25f * 1.4426950408889634f;   // log2(e)/sqrt(16): softmax through exp2
  bool staged = false;
  // query pair idx = (row pair rp, column): one pass for regions up to 32 wide, two for 64
#pragma unroll 1
  for (int idx = threadIdx.x; idx < (kAttnRB / 2) * qr.rn; idx += kAttnThreads) {
  const int rp = idx / qr.rn;
  const int x = qr.r0 + idx - rp * qr.rn;
  const int y0 = r0 + 2 * rp;
  const bool active = y0 < rend;
  const bool two = active && (y0 + 1) < rend;
  const long long base = planar_off(v, head, 0, 0, x, P);
  f32x2 q0[8], q1[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 f = active ? __ldg(reinterpret_cast<const float4*>(Q + base + y0 * rowstride + j * jstride))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    q0[2 * j] = pack2(f.x * qs, f.y * qs);
    q0[2 * j + 1] = pack2(f.z * qs, f.w * qs);
    const float4 g = two ? __ldg(reinterpret_cast<const float4*>(Q + base + (y0 + 1) * rowstride + j * jstride))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    q1[2 * j] = pack2(g.x * qs, g.y * qs);
    q1[2 * j + 1] = pack2(g.z * qs, g.w * qs);
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  f32x2 o0[8], o1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { o0[e] = 0ull; o1[e] = 0ull; }
  if (!staged) { mbar_wait(bar, 0); staged = true; }
  if (active) {
#pragma unroll
    for (int kr = 0; kr < 6; ++kr) {
      const int ky = y0 - 2 + kr;
      if (ky < 0 || ky >= P) continue;
      const float* krow = reinterpret_cast<const float*>(ks) + (size_t)(ky - ys) * (rowbytes / 4) + x * 4;
      const float* vrow = reinterpret_cast<const float*>(vs) + (size_t)(ky - ys) * (rowbytes / 4) + x * 4;
      float a0[5], a1[5];
      float n0 = m0, n1 = m1;
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int kx = x + dx;
        a0[dx + 2] = -INFINITY;
        a1[dx + 2] = -INFINITY;
        if (kx >= 0 && kx < P) {
          const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(krow + dx * 4);
          const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(krow + dx * 4 + jstride);
          const ulonglong2 k2 = *reinterpret_cast<const ulonglong2*>(krow + dx * 4 + 2 * jstride);
          const ulonglong2 k3 = *reinterpret_cast<const ulonglong2*>(krow + dx * 4 + 3 * jstride);
          if (kr <= 4) { a0[dx + 2] = dot16(q0, k0, k1, k2, k3); n0 = fmaxf(n0, a0[dx + 2]); }
          if (kr >= 1) { a1[dx + 2] = dot16(q1, k0, k1, k2, k3); n1 = fmaxf(n1, a1[dx + 2]); }
        }
      }
      if (kr <= 4) {
        const float sc = fast_exp2(m0 - n0);
        m0 = n0;
        l0 *= sc;
        const f32x2 sc2 = pack2(sc, sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) o0[e] = mul2(o0[e], sc2);
      }
      if (kr >= 1 && two) {
        const float sc = fast_exp2(m1 - n1);
        m1 = n1;
        l1 *= sc;
        const f32x2 sc2 = pack2(sc, sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) o1[e] = mul2(o1[e], sc2);
      }
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int kx = x + dx;
        if (kx >= 0 && kx < P) {
          const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(vrow + dx * 4);
          const ulonglong2 v1 = *reinterpret_cast<const ulonglong2*>(vrow + dx * 4 + jstride);
          const ulonglong2 v2 = *reinterpret_cast<const ulonglong2*>(vrow + dx * 4 + 2 * jstride);
          const ulonglong2 v3 = *reinterpret_cast<const ulonglong2*>(vrow + dx * 4 + 3 * jstride);
          if (kr <= 4) {
            const float p0 = fast_exp2(a0[dx + 2] - m0);
            l0 += p0;
            axpy16(o0, p0, v0, v1, v2, v3);
          }
          if (kr >= 1 && two) {
            const float p1 = fast_exp2(a1[dx + 2] - m1);
            l1 += p1;
            axpy16(o1, p1, v0, v1, v2, v3);
          }
        }
      }
    }
    const float i0 = 1.f / l0;
    const f32x2 i02 = pack2(i0, i0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ulonglong2 r;
      r.x = mul2(o0[2 * j], i02);
      r.y = mul2(o0[2 * j + 1], i02);
      st_stream_v4(O + base + y0 * rowstride + j * jstride, *reinterpret_cast<const float4*>(&r));
    }
    if (two) {
      const float i1 = 1.f / l1;
      const f32x2 i12 = pack2(i1, i1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ulonglong2 r;
        r.x = mul2(o1[2 * j], i12);
        r.y = mul2(o1[2 * j + 1], i12);
        st_stream_v4(O + base + (y0 + 1) * rowstride + j * jstride, *reinterpret_cast<const float4*>(&r));
      }
    }
  }
  }  // query pairs
  if (!staged) mbar_wait(bar, 0);  // threads without a query still wait for the copies before the CTA may exit
}

#endif  // LFT_ATTN_V1

// ------------------------------------------------------------------------------------------------
// Window attention on tensor cores (round 2, the default): S = Q K^T and O = P V of the 5 x 5 window as warp-level
// mma.sync.m16n8k16 (bf16 operands, fp32 accumulate) with the three-term hi / lo split of every fp32-grade product.  The
// window attention is ragged and tiny (<= 25 keys of 16 dims per query), which tcgen05's 128-row tiles cannot hold without
// ~10x padding; a register-fragment MMA can: a warp takes a 4 x 4 block of queries (one m16 tile) of one head, whose windows lie
// inside the 8 x 8 block of keys around it = 64 keys = eight n8 tiles (one key row each) for S and four k16 steps for P V.
// 39 % of the S / P entries are inside a window, the rest is masked to -inf / 0 - still 2.2x fewer issue slots and 3.7x fewer
// shared-memory wavefronts per query than the FFMA2 formulation above, which was bound by getting every key to its 25 queries
// through the LSU.  CTA = (view, head, 8 query rows) as before: K / V rows [r0-2, r0+10) arrive by two bulk copies; B fragments
// come from them by ldmatrix (K) / ldmatrix.trans (V) - the 16-byte pieces of the planar layout ARE the 8 x 8 matrix rows, eight
// neighbouring keys are 128 contiguous bytes (conflict-free); Q fragments are read straight from global memory (32-bit words).
// Keys outside a query's window are masked after the MMA; their addresses are clamped to the rows / columns this launch's
// k_spa_embed_qkv has written (masked keys lie at most 3 pixels outside the K / V region), so no stale workspace bits ever reach a
// multiplier.  Soft-max in fp32 on the accumulator fragments (row statistics by quad shuffles), P re-split into hi / lo.
LFT_DEVINL void ldsm4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
LFT_DEVINL void ldsm4t(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
LFT_DEVINL void hmma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two probabilities -> packed bf16 hi (truncation) and lo (rounded remainder); bf16 mode: hi = rounded, no lo
LFT_DEVINL void split_pair(float a, float b, bool fp32_mode, uint32_t& hi, uint32_t& lo) {
  if (!fp32_mode) { hi = pack_bf16(a, b); lo = 0u; return; }
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  lo = pack_bf16(a - __uint_as_float(ua & 0xffff0000u), b - __uint_as_float(ub & 0xffff0000u));
}

#ifndef LFT_ATTN_SKEW
#define LFT_ATTN_SKEW 500
#endif
constexpr int kAttnMmaThreads = 256;  // 8 warps: warp w takes block row w & 1 of the item's 8 rows and every fourth block column
// Query blocks sit on the ABSOLUTE 4 x 4 grid of the view (not the region's), so a query meets its keys in the same order
// whatever region a launch computes: the light-field path's sub-region results stay bit-identical to the full forward's.
LFT_DEVINL int attn_mma_nblk(Region qr) { return ((qr.r0 + qr.rn) - (qr.r0 & ~3) + kAttnRB - 1) / kAttnRB; }
constexpr size_t smem_attn_mma(int P) { return 32 + 2 * 2 * (size_t)(kAttnRB + 4) * 4 * P * 16; }  // two (K rows | V rows) buffers

// Persistent: a CTA (two per SM) walks over the work items (view, head, block of 8 query rows) i = blockIdx.x + k gridDim.x with
// two staging buffers - the bulk copies of item k + 1 run under the MMAs of item k (the one-item-per-CTA form spent half of a
// CTA's life waiting for its 49 KB).  The host makes gridDim.x a multiple of the row blocks per view, so a CTA keeps its row
// block: key-row offsets and row masks are loop invariants.
template <bool FP32, bool OTILE>
__global__ void __launch_bounds__(kAttnMmaThreads, 2)
k_spa_attn_mma(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ Vv,
               float* __restrict__ O, int P, Region qr, int nitems) {
  constexpr bool otile = OTILE;
  // otile != 0: O leaves as the bf16 hi / lo A operand of k_spa_ffn2's output projection, tile by tile in the operand's own
  // shared-memory layout [tile of 128 compacted tokens][hi | lo][k chunk 16][row 128][8 bf16] (64 KB per tile, channel =
  // head * 16 + dim), which k_spa_ffn2 fetches with two bulk copies; otile == 0: fp32 planar (the layout of Q / K / V).
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nblk = attn_mma_nblk(qr);
  const int rb = blockIdx.x % nblk;
  const int rend = qr.r0 + qr.rn, cend = rend;                     // one past the last query row / column
  const int rowbase = (qr.r0 & ~3) + rb * kAttnRB;                 // first row of the CTA's blocks (multiple of 4)
  const int q0 = max(rowbase, qr.r0), q1 = min(rowbase + kAttnRB, rend);     // query rows [q0, q1) of this CTA
  const int ys = max(q0 - 2, 0), ye = min(q1 + 2, P);              // staged key rows [ys, ye)
  const int kxlo = max(qr.r0 - 2, 0), kxhi = min(cend + 2, P) - 1; // key columns k_spa_embed_qkv has written
  constexpr int NP = FP32 ? 4 : 2;                                 // 16-byte pieces per (y, x): hi, hi | lo, lo
  const uint32_t rowbytes = (uint32_t)P * 16 * NP;                 // one (y) plane: NP pieces x P x 16 B
  const uint32_t piecebytes = (uint32_t)P * 16;
  const uint32_t nbytes = (uint32_t)(ye - ys) * rowbytes;
  const uint32_t bufbytes = 2u * (kAttnRB + 4) * rowbytes;
  const uint32_t bar0 = smem_u32(smem);
  const uint32_t buf0 = smem_u32(smem) + 32;
  unsigned* cnt = reinterpret_cast<unsigned*>(smem + 16);      // warps that have finished with buffer 0 / 1 (monotonic)
  const long long PP16 = (long long)P * P * 4 * NP;                // floats per (view, head) plane
  const int nk = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const unsigned vh0 = blockIdx.x / (unsigned)nblk, dvh = gridDim.x / (unsigned)nblk;   // gridDim.x % nblk == 0 (host)
  auto plane_of = [&](int k) { return (long long)(vh0 + (unsigned)k * dvh) * PP16; };
  pdl_trigger();
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    cnt[0] = cnt[1] = 0u;
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();  // Q / K / V are the previous kernel's output
  auto issue = [&](int k) {   // thread 0: stage item k's K / V rows into buffer k & 1
    const long long src = plane_of(k) + (long long)ys * (rowbytes / 4);
    const uint32_t bar = bar0 + 8u * (k & 1), dst = buf0 + (uint32_t)(k & 1) * bufbytes;
    mbar_arrive_expect_tx(bar, 2 * nbytes);
    bulk_g2s(dst, K + src, nbytes, bar);
    bulk_g2s(dst + (kAttnRB + 4) * rowbytes, Vv + src, nbytes, bar);
  };
  if (threadIdx.x == 0) {
    if (nk > 0) issue(0);
    if (nk > 1) issue(1);
  }
  constexpr bool fp32m = FP32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;      // accumulator fragment: rows g, g + 8; columns 2c, 2c + 1
  const int mat = lane >> 3, mr = lane & 7;   // ldmatrix: this lane addresses row mr of matrix mat
  constexpr int kBR = kAttnRB / 4;            // block rows per item
  constexpr int kWC = kAttnMmaThreads / 32 / kBR;   // warps per block row = block-column stride of a warp
  const int bx0 = qr.r0 & ~3;
  const int ncb = (cend - bx0 + 3) >> 2;
  const float qs = 0.25f * 1.4426950408889634f;   // log2(e)/sqrt(16): softmax through exp2
  const float NINF = -INFINITY;

  // ---- per-warp invariants: the block row by, its key rows and row masks.
  // m16 tile rows: A = g -> query (by + iA, bx + g%4), B = g + 8 -> (by + 2 + iA, .), iA = g / 4; n8 tile j = key row by - 2 + j.
  const int by = rowbase + 4 * (warp % kBR);
  const bool row_active = by < q1 && by + 4 > q0;
  const int iA = g >> 2;
  uint32_t koff[8], voff[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) koff[j] = (uint32_t)(min(max(by - 2 + j, ys), ye - 1) - ys) * rowbytes + (uint32_t)mat * piecebytes;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    voff[t] = (uint32_t)(min(max(by - 2 + 2 * t + (mat & 1), ys), ye - 1) - ys) * rowbytes + (uint32_t)(mat >> 1) * piecebytes;
  // additive row masks (0 / -inf) as accumulator initial values: row A meets key rows j = iA .. iA + 4 (<= 5), row B
  // j = iA + 2 .. iA + 6 (>= 2), both only inside the view; the (row half, j) pairs outside those ranges are masked for every lane
  float rbA[6], rbB[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int kyA = by - 2 + j, kyB = by + j;          // rbB[j] belongs to n8 tile j + 2
    rbA[j] = (j >= iA && j <= iA + 4 && kyA >= 0 && kyA < P) ? 0.f : NINF;
    rbB[j] = (j + 2 >= iA + 2 && j + 2 <= iA + 6 && kyB >= 0 && kyB < P) ? 0.f : NINF;
  }
  bool dxok[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int dx = 2 * c + e - 2 - (g & 3);
    dxok[e] = dx >= -2 && dx <= 2;
  }
  const int qyA = min(max(by + iA, qr.r0), rend - 1), qyB = min(max(by + 2 + iA, qr.r0), rend - 1);
  const uint32_t qoffA = (uint32_t)qyA * (4u * NP) * P + c, qoffB = (uint32_t)qyB * (4u * NP) * P + c;   // 32-bit words inside the plane
  const int ps = P * 4;                       // words between the pieces of one (y, x)
  const bool odd = c & 1;
  const int qyo = by + iA + (odd ? 2 : 0);    // the row this lane stores (even lanes: row A's piece, odd lanes: row B's)
  const bool okrow = qyo >= qr.r0 && qyo < rend;
  const uint32_t ooff = (uint32_t)(qyo * 4 + (c >> 1)) * P * 4;
  // tile-format output: compacted token of (view, y, x) = (view * rn + y - r0) * rn + x - r0
  const int tyA = by + iA - qr.r0, tyB = tyA + 2;
  const bool okA = tyA >= 0 && tyA < qr.rn, okB = tyB >= 0 && tyB < qr.rn;
  const unsigned tyrn[2] = {(unsigned)(tyA * qr.rn), (unsigned)(tyB * qr.rn)};
  const unsigned rn2 = (unsigned)(qr.rn * qr.rn);
  uint32_t* Ot = reinterpret_cast<uint32_t*>(O) + c;

  // Q fragments of block column cb of the item whose plane starts at Qw; queries outside the region shadow the nearest valid
  // one (their results are not stored).  Row B and the pieces are fixed 64-bit strides from row A's first piece.
  const long long dAB = (long long)qoffB - (long long)qoffA, ps1 = ps, ps2 = 2 * ps, ps3 = 3 * ps;
  auto load_q = [&](const uint32_t* Qw, int cb, uint32_t* qh, uint32_t* ql) {
    const int qx = min(bx0 + 4 * cb + (g & 3), P - 1);   // (columns left / right of the region: any written or stale data will do)
    const uint32_t* a = Qw + (qoffA + (uint32_t)qx * 4u);
    const uint32_t* b = a + dAB;
    qh[0] = __ldg(a); qh[1] = __ldg(b); qh[2] = __ldg(a + ps1); qh[3] = __ldg(b + ps1);
    if (fp32m) { ql[0] = __ldg(a + ps2); ql[1] = __ldg(b + ps2); ql[2] = __ldg(a + ps3); ql[3] = __ldg(b + ps3); }
  };
  const uint32_t* Qall = reinterpret_cast<const uint32_t*>(Q);
  uint32_t qh[4] = {0u, 0u, 0u, 0u}, ql[4] = {0u, 0u, 0u, 0u};
  const int cb0 = warp / kBR;
  const bool has_work = row_active && cb0 < ncb;
  if (has_work && nk > 0) load_q(Qall + plane_of(0), cb0, qh, ql);
#ifndef LFT_ATTN_NOSKEW
  if (warp >= 4) {   // the two warps of a scheduler (w, w + 4) start half a block apart
    const long long t0 = clock64();
    while (clock64() - t0 < LFT_ATTN_SKEW) {}
  }
#endif
#pragma unroll 1
  for (int k = 0; k < nk; ++k) {
    const uint32_t ks = buf0 + (uint32_t)(k & 1) * bufbytes, vs = ks + (kAttnRB + 4) * rowbytes;
    const long long plane = plane_of(k);
    const unsigned vhk = vh0 + (unsigned)k * dvh;
    const unsigned tvq = (vhk >> 3) * rn2;                          // first compacted token of the item's view
    uint32_t* Oh = Ot + (vhk & 7u) * 1024u;                         // + head * 2 k chunks * 512 words
    const uint32_t* Qw = Qall + plane;
    const uint32_t* Qn = Qall + plane_of(k + 1);
    float* Ob = O + (long long)vhk * ((long long)P * P * 16);   // (fp32 planar O: four pieces per (y, x) in both modes)
    mbar_wait(bar0 + 8u * (k & 1), (uint32_t)(k >> 1) & 1u);   // item k's rows have landed
    if (has_work) {
#pragma unroll 1
      for (int cb = cb0; cb < ncb; cb += kWC) {
        const int bx = bx0 + 4 * cb;
        // ---- S = Q K^T on top of the masks: n8 tile j = key row by - 2 + j, keys bx - 2 .. bx + 5 (column = 2c + e of the tile)
        const uint32_t kcol = (uint32_t)min(max(bx - 2 + mr, kxlo), kxhi) * 16;
        float cb_[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) cb_[e] = (dxok[e] && (unsigned)(bx - 2 + 2 * c + e) < (unsigned)P) ? 0.f : NINF;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
                s[j][e] = j < 6 ? fminf(rbA[j], cb_[e]) : 0.f;          // (the row halves no lane's window reaches are never read)
            s[j][2 + e] = j >= 2 ? fminf(rbB[j - 2], cb_[e]) : 0.f;
          }
          uint32_t kb[4];
          ldsm4(ks + koff[j] + kcol, kb);
          hmma16816(s[j], qh, kb[0], kb[1]);
          if (fp32m) {
            hmma16816(s[j], ql, kb[0], kb[1]);
            hmma16816(s[j], qh, kb[2], kb[3]);
          }
        }
        // the Q fragments are dead: the next block's queries (of this item or of the next one) are loaded straight into their
        // registers and have the soft-max and P V of this block to arrive
        {
          const bool same = cb + kWC < ncb;
          if (same || k + 1 < nk) load_q(same ? Qw : Qn, same ? cb + kWC : cb0, qh, ql);
        }
        // ---- soft-max (masked entries are -inf; every query sees itself, so the row maxima are finite).  Packed f32x2 math on
        // the accumulator pairs, two partial sums per row half (short dependency chains: four warps per scheduler is all the
        // latency hiding there is)
        float mA0 = fmaxf(s[0][0], s[0][1]), mA1 = fmaxf(s[1][0], s[1][1]);
        float mB0 = fmaxf(s[2][2], s[2][3]), mB1 = fmaxf(s[3][2], s[3][3]);
#pragma unroll
        for (int j = 2; j < 6; j += 2) {
          mA0 = fmaxf(mA0, fmaxf(s[j][0], s[j][1]));
          mA1 = fmaxf(mA1, fmaxf(s[j + 1][0], s[j + 1][1]));
          mB0 = fmaxf(mB0, fmaxf(s[j + 2][2], s[j + 2][3]));
          mB1 = fmaxf(mB1, fmaxf(s[j + 3][2], s[j + 3][3]));
        }
        float mA = fmaxf(mA0, mA1), mB = fmaxf(mB0, mB1);
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
        const f32x2 qs2 = pack2(qs, qs), nA2 = pack2(-mA * qs, -mA * qs), nB2 = pack2(-mB * qs, -mB * qs);
        const f32x2 neg1 = pack2(-1.f, -1.f);
        f32x2 lA2[2] = {0ull, 0ull}, lB2[2] = {0ull, 0ull};
        // P as the A operands of the four k16 steps: pa[t] = {row A tile 2t, row B tile 2t, row A tile 2t+1, row B tile 2t+1}
        uint32_t pa[4][4], pb[4][4];
        auto prob = [&](float x0, float x1, f32x2 n2, f32x2& l2, uint32_t& hi, uint32_t& lo) {
          float e0, e1;
          unpack2(fma2(pack2(x0, x1), qs2, n2), e0, e1);
          const float p0 = fast_exp2(e0), p1 = fast_exp2(e1);
          const f32x2 pp = pack2(p0, p1);
          l2 = add2(l2, pp);
          if (fp32m) {
            const uint32_t u0 = __float_as_uint(p0), u1 = __float_as_uint(p1);
            hi = __byte_perm(u0, u1, 0x7632);
            float r0, r1;   // p - trunc_bf16(p), exact
            unpack2(fma2(pack2(__uint_as_float(u0 & 0xffff0000u), __uint_as_float(u1 & 0xffff0000u)), neg1, pp), r0, r1);
            lo = pack_bf16(r0, r1);
          } else {
            hi = pack_bf16(p0, p1);
            lo = 0u;
          }
        };
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = j >> 1, h = (j & 1) * 2;
          if (j < 6) prob(s[j][0], s[j][1], nA2, lA2[j & 1], pa[t][h], pb[t][h]);
          else pa[t][h] = pb[t][h] = 0u;
          if (j >= 2) prob(s[j][2], s[j][3], nB2, lB2[j & 1], pa[t][h + 1], pb[t][h + 1]);
          else pa[t][h + 1] = pb[t][h + 1] = 0u;
        }
        float lA = hsum2(add2(lA2[0], lA2[1])), lB = hsum2(add2(lB2[0], lB2[1]));
        lA += __shfl_xor_sync(0xffffffffu, lA, 1);
        lB += __shfl_xor_sync(0xffffffffu, lB, 1);
        lA += __shfl_xor_sync(0xffffffffu, lA, 2);
        lB += __shfl_xor_sync(0xffffffffu, lB, 2);
        // ---- O = P V: k16 step t = key rows by - 2 + 2t, + 1; n8 tile d = dims 8d .. 8d + 7
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float o2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // odd k steps: two more independent accumulation chains
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t va = vs + voff[t] + kcol;
          float (*oo)[4] = (t & 1) ? o2 : o;
          uint32_t vh[4];
          ldsm4t(va, vh);   // {keys 0-7, keys 8-15} x {dims 0-7, dims 8-15} of the hi pieces
          hmma16816(oo[0], pa[t], vh[0], vh[1]);
          hmma16816(oo[1], pa[t], vh[2], vh[3]);
          if (fp32m) {
            hmma16816(oo[0], pb[t], vh[0], vh[1]);
            hmma16816(oo[1], pb[t], vh[2], vh[3]);
            uint32_t vl[4];
            ldsm4t(va + 2 * piecebytes, vl);
            hmma16816(oo[0], pa[t], vl[0], vl[1]);
            hmma16816(oo[1], pa[t], vl[2], vl[3]);
          }
        }
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            float x0, x1;
            unpack2(add2(pack2(o[d][i], o[d][i + 1]), pack2(o2[d][i], o2[d][i + 1])), x0, x1);
            o[d][i] = x0; o[d][i + 1] = x1;
          }
        // ---- normalise and store (fp32 planar, as k_spa_ffn reads it): the lanes of a pair (c, c ^ 1) hold the two halves of
        // a 16-byte piece; the even lane stores row A's piece, the odd lane row B's
        const float iA_ = fast_rcp(lA), iB_ = fast_rcp(lB);   // l in [1, 25]: rcp.approx is within 1 ulp
        const int qx = bx + (g & 3);
        if (otile) {
          const unsigned tx = (unsigned)(qx - qr.r0);
          if (tx < (unsigned)qr.rn) {
            const unsigned tv = tvq + tx;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              if (hf ? okB : okA) {
                const unsigned t = tv + tyrn[hf];
                // 32-bit words: tile * 16384 + plane * 8192 + k chunk * 512 + row * 4 + c
                uint32_t* dst = Oh + (((t >> 7) << 14) | ((t & 127u) << 2));
                const f32x2 inv2 = hf ? pack2(iB_, iB_) : pack2(iA_, iA_);
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                  const f32x2 v2 = mul2(pack2(o[d][2 * hf], o[d][2 * hf + 1]), inv2);
                  float x0, x1;
                  unpack2(v2, x0, x1);
                  if (fp32m) {
                    const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
                    float r0_, r1_;
                    unpack2(fma2(pack2(__uint_as_float(u0 & 0xffff0000u), __uint_as_float(u1 & 0xffff0000u)), neg1, v2), r0_, r1_);
                    dst[d * 512] = __byte_perm(u0, u1, 0x7632);
                    dst[8192 + d * 512] = pack_bf16(r0_, r1_);
                  } else {
                    dst[d * 512] = pack_bf16(x0, x1);
                  }
                }
              }
            }
          }
        } else {
        const bool okq = okrow && qx >= qr.r0 && qx < cend;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const float a0 = o[d][0] * iA_, a1 = o[d][1] * iA_, b0 = o[d][2] * iB_, b1 = o[d][3] * iB_;
          const float r0_ = __shfl_xor_sync(0xffffffffu, odd ? a0 : b0, 1);
          const float r1_ = __shfl_xor_sync(0xffffffffu, odd ? a1 : b1, 1);
          const float4 out = odd ? make_float4(r0_, r1_, b0, b1) : make_float4(a0, a1, r0_, r1_);
          if (okq) st_stream_v4(Ob + ooff + (uint32_t)(2 * d * P + qx) * 4, out);
        }
        }
      }
    }
    // Buffer release without a CTA-wide barrier (the warps of a scheduler should drift apart: in lockstep they all want the
    // tensor pipe, then all the ALUs): every warp counts itself off item k; the last one refills the buffer with item k + 2.
    // A warp's ldmatrix reads have completed when it gets here (their HMMAs have been issued).
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();   // release: this warp's reads of the buffer are ordered before the count ...
      const unsigned done = atomicAdd(&cnt[k & 1], 1u) + 1u;
      __threadfence_block();   // ... acquire: and every counted warp's before the refill issued below
      if (done == (unsigned)(kAttnMmaThreads / 32) * (unsigned)((k >> 1) + 1) && k + 2 < nk) issue(k + 2);
    }
  }
}

#ifdef LFT_FFN_V1   // the first formulation of k_spa_ffn is only compiled into -DLFT_FFN_V1 builds (A/B timing)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads2, 2)
k_spa_ffn(const float* __restrict__ O, float* __restrict__ tok, const __grid_constant__ Tab512 tab,
          const uint8_t* __restrict__ wo, const uint8_t* __restrict__ w1a, const uint8_t* __restrict__ w1b,
          const uint8_t* __restrict__ w2a, const uint8_t* __restrict__ w2b, const uint8_t* __restrict__ wlin,
          float* __restrict__ out, const float* __restrict__ final_res, long long T, int P, int passes, Region fr) {
  // fr: the pixels of every view this launch computes (full view: {0, P}); T = views * fr.rn^2 compacted tokens, 128 per CTA
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t A = smem_u32(smem) + kCtlBytes;
  const uint32_t ring = A + 65536;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const uint32_t hb_ready = smem_u32(&ctl->aux[0]);  // 256 arrivals, used once per CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  LFT_TL(30);
  pdl_trigger();
  cta_setup<kSpaNST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_o{wo, 128, 2}, g_1a{w1a, 128, 2}, g_2a{w2a, 128, 2}, g_1b{w1b, 128, 2}, g_2b{w2b, 128, 2},
      g_l{wlin, 64, 2};

  if (warp == kWarpProducer2) {

    RingState<kSpaNST> rs;
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_o, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1a, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1b, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2a, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2b, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_l, passes);
  } else if (warp == kWarpMma2) {

    RingState<kSpaNST> rs;
    uint32_t par = 0;
    int tl = 0;
    auto wait_a = [&]() {
      mbar_wait(a_ready, par);
      par ^= 1;
      tc_fence_after();
      LFT_TL(tl); ++tl;
    };
    auto gemm = [&](const GemmPhase& g, uint32_t dcol, bool fresh) {
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + dcol, fresh);
    };
    auto done = [&]() { umma_commit_elected(mma_done); LFT_TL(tl); ++tl; };
    // TS form (A operand in TMEM columns [128,192) hi | [192,256) lo) wherever those columns are free
    auto gemm_ts = [&](const GemmPhase& g, uint32_t dcol, bool fresh) {
      ring_consume_mma_ts<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g, passes, tmem + 128, tmem + 192, tmem + dcol, fresh);
    };
    wait_a(); gemm_ts(g_o, 0, true); done();                               // D[0,128)   = O Wo^T                    (TS)
    wait_a(); gemm(g_1a, 0, true); gemm(g_1b, 128, true); done();          // D[0,256)   = Y1 W'1^T  (both halves)   (SS)
    wait_a(); gemm(g_2a, 0, true);                                         // D[0,128)   = relu(.)[:, :128] W2a^T    (SS)
    // hidden[:, 128:] is published on its OWN barrier: the row owners reach that arrival without an intervening wait, so on
    // a_ready a fast warp's second arrival would be counted into the phase slower warps have not finished (FFN2a would
    // start on a half-written operand)
    mbar_wait(hb_ready, 0);
    tc_fence_after();
    LFT_TL(tl); ++tl;
    gemm_ts(g_2b, 0, false); done();                                       // D[0,128)  += relu(.)[:, 128:] W2b^T    (TS); ONE commit for FFN2a+b
    wait_a(); gemm_ts(g_l, 0, true); done();                               // D[0,64)    = Y2 Wlin^T                 (TS)
  } else {
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    const long long t = (long long)blockIdx.x * 128 + m;
    const bool ok = t < T;
    const unsigned PP = (unsigned)(P * P), RR = (unsigned)(fr.rn * fr.rn);
    // compacted token index -> (view, y, x) -> token of the full view; T < 2^31 (checked on the host): 32-bit index math
    auto locate = [&](unsigned tc, unsigned& vu, int& y, int& x) {
      vu = tc / RR;
      const int rem = (int)(tc - vu * RR);
      const int yy = rem / fr.rn;
      y = fr.r0 + yy;
      x = fr.r0 + rem - yy * fr.rn;
      return vu * PP + (unsigned)(y * P + x);
    };
    unsigned vu;
    int y, x;
    const unsigned tt = locate(ok ? (unsigned)t : 0u, vu, y, x);
    const long long v = vu;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* trow_g = tok + t32_off(tt, 16 * q, 32);  // own half of the token row, chunk stride 128 floats (Y1 is spilled here)
    uint32_t par = 0;
    int tl = 1;
    auto publish = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
      LFT_TL(tl); ++tl;
    };
    auto await = [&]() {
      mbar_wait(mma_done, par);
      par ^= 1;
      tc_fence_after();
      LFT_TL(tl); ++tl;
    };
    auto publish_tmem = [&]() {  // A operand written with tcgen05.st
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(a_ready);
      LFT_TL(tl); ++tl;
    };
    const bool fp32m = passes == 3;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    LFT_TL(0);
    pdl_wait();  // O / tok are the previous kernels' output

    // phase 0: A <- O (planar gather of own heads 4q..4q+3)
    {
      float4 f[16];
      const float* ob = O + planar_off(v, 4 * q, y, 0, x, P);
      const long long hs = (long long)PP * 16, js = (long long)P * 4;  // head / piece strides in floats
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[4 * c + j] = ok ? __ldg(reinterpret_cast<const float4*>(ob + c * hs + j * js)) : zero4;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, reinterpret_cast<const float*>(&f[4 * c]), fp32m);
    }
    publish_tmem();

    // phase 1: Y1 = tok + D (own half) -> global spill + raw A operand; LN2 statistics (folded into FFN1's epilogue)
    float mean, rstd;
    {
      float4 tk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) tk[i] = ok ? *reinterpret_cast<const float4*>(trow_g + 128 * i) : zero4;  // in flight
      await();
      float yv[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, yv + 16 * c);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        yv[4 * i] += tk[i].x; yv[4 * i + 1] += tk[i].y; yv[4 * i + 2] += tk[i].z; yv[4 * i + 3] += tk[i].w;
      }
      if (ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          *reinterpret_cast<float4*>(trow_g + 128 * i) = make_float4(yv[4 * i], yv[4 * i + 1], yv[4 * i + 2], yv[4 * i + 3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) a_store16(A, 8 * q + 2 * c, m, yv + 16 * c, passes == 3);
      pair_ln_stats<64>(yv, trow + 64 * q, trow + 64 * (1 - q), 1 + (warp & 3), mean, rstd);
    }
    publish();
    const float mr = mean * rstd;
    const float4* u1 = reinterpret_cast<const float4*>(tab.v);        // [u_1 256 | c_1 256] (constant bank)
    const float4* c1 = reinterpret_cast<const float4*>(tab.v + 256);

    {  // speculative L2 prefetch of the O pieces and token row of the tile 2 CTAs x 148 SMs ahead, i.e. of the CTA that follows
       // on some SM about one tile time from now (k_spa_ffn -1 %; a persistent version of this kernel was 7 % slower)
      const long long tn = t + 296ll * 128;
      if (tn < T) {
        unsigned vn;
        int yn, xn;
        const unsigned tnu = locate((unsigned)tn, vn, yn, xn);
        const float* obn = O + planar_off(vn, 4 * q, yn, 0, xn, P);
        const long long hs = (long long)PP * 16, js = (long long)P * 4;
        const float* tgn = tok + t32_off(tnu, 16 * q, 32);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) prefetch_l2(obn + c * hs + j * js);
#pragma unroll
        for (int i = 0; i < 16; ++i) prefetch_l2(tgn + 128 * i);
      }
    }
    // phase 2: hidden[:, :128] (own 64 columns of D[0,128)) -> A
    await();
    float dh[64];  // all four accumulator loads in flight, one wait
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, dh + 16 * c);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float* d = dh + 16 * c;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 uv = u1[16 * q + 4 * c + j], cv = c1[16 * q + 4 * c + j];
        d[4 * j] = fmaxf(fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x)), 0.f);
        d[4 * j + 1] = fmaxf(fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y)), 0.f);
        d[4 * j + 2] = fmaxf(fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z)), 0.f);
        d[4 * j + 3] = fmaxf(fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w)), 0.f);
      }
      a_store16(A, 8 * q + 2 * c, m, d, passes == 3);
    }
    publish();

    // phase 3: hidden[:, 128:] from D[128,256) -> TMEM A operand (the same columns, once both partner threads have
    // read their halves); it does not touch the smem operand FFN2a is still reading, so no wait is needed here
    {
      float hb[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 128 + 64 * q + 16 * c, hb + 16 * c);
      tmem_wait_ld();
      tc_fence_before();
      pair_bar_sync(warp & 3);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 uv = u1[32 + 16 * q + j], cv = c1[32 + 16 * q + j];
        hb[4 * j] = fmaxf(fmaf(rstd, hb[4 * j], fmaf(-mr, uv.x, cv.x)), 0.f);
        hb[4 * j + 1] = fmaxf(fmaf(rstd, hb[4 * j + 1], fmaf(-mr, uv.y, cv.y)), 0.f);
        hb[4 * j + 2] = fmaxf(fmaf(rstd, hb[4 * j + 2], fmaf(-mr, uv.z, cv.z)), 0.f);
        hb[4 * j + 3] = fmaxf(fmaf(rstd, hb[4 * j + 3], fmaf(-mr, uv.w, cv.w)), 0.f);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, hb + 16 * c, fp32m);
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(hb_ready);  // NOT a_ready: no wait separates this arrival from the previous one (see the MMA warp)
    LFT_TL(tl); ++tl;

    // phase 4: Y2 = Y1 (spilled row, prefetched) + D[0,128) -> A
    {
      float4 y1[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) y1[i] = ok ? *reinterpret_cast<const float4*>(trow_g + 128 * i) : zero4;
      await();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float d[16];
        tmem_ld16(trow + 64 * q + 16 * c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d[4 * j] += y1[4 * c + j].x; d[4 * j + 1] += y1[4 * c + j].y;
          d[4 * j + 2] += y1[4 * c + j].z; d[4 * j + 3] += y1[4 * c + j].w;
        }
        a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, d, fp32m);
      }
    }
    publish_tmem();

    // phase 5: out = D[0,64) (+ global residual), own 32 columns
    float4 r4[8];
    if (final_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r4[i] = ok ? __ldg(reinterpret_cast<const float4*>(final_res + t32_off(tt, 8 * q + i, 16))) : zero4;
    }
    await();
    {
      float d[32];
      tmem_ld16_nowait(trow + 32 * q, d);
      tmem_ld16_nowait(trow + 32 * q + 16, d + 16);
      tmem_wait_ld();
      if (ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o4 = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
          if (final_res) { o4.x += r4[i].x; o4.y += r4[i].y; o4.z += r4[i].z; o4.w += r4[i].w; }
          *reinterpret_cast<float4*>(out + t32_off(tt, 8 * q + i, 16)) = o4;
        }
      }
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
  LFT_TL(31);
}

#endif  // LFT_FFN_V1

// ------------------------------------------------------------------------------------------------
// k_spa_ffn, second formulation (round 2; the default, -DLFT_FFN_V1 selects the one above).  Same arithmetic, different residency:
//   * the FFN is evaluated half by half - FFN1a -> hidden_a -> FFN2a, then FFN1b -> hidden_b -> FFN2b - so that BOTH hidden
//     halves become tcgen05 TS operands in TMEM columns [128,256) (the first formulation wrote hidden_a to shared memory: 64
//     st.shared.v4 per thread through the port the MMA operands are read from, 5.6 K of its 33.7 K cycles per tile), and
//   * Y1 (the residual stream after the attention) is never spilled to global memory: its bf16 hi/lo operand stays in shared
//     memory from FFN1a to the end (nothing overwrites it any more) and Y2 = (Y1_hi + Y1_lo) + FFN2 reconstructs it - the
//     pair carries Y1 to 2^-17 relative, the accuracy class of every three-pass product here (in bf16 mode the hi part alone
//     would be a 2^-9 error on the residual, so that mode keeps the spill).  1 KB / token less HBM traffic.
//   TMEM: D = [0,128) accumulator, T = [128,256) TS operand (hi [128,192) | lo [192,256)) and, for FFN1b, accumulator.
//   Row phases / a_ready arrivals (each separated from the next by a wait on an MMA that needed the previous one complete):
//     O -> T | Y1 -> smem A (+ LN2 statistics) | hidden_a -> T | hidden_b -> T (in place over the FFN1b accumulator) | Y2 -> T
//   MMA phases / mma_done commits: D = O Wo^T | D = Y1 W1a^T | D = hidden_a W2a^T, T = Y1 W1b^T | D += hidden_b W2b^T | D = Y2 Wlin^T
//   (FFN1b's accumulator is the region FFN2a reads its operand from: the MMA thread waits for FFN2a's own commit - aux[1] - first.)
__global__ void __launch_bounds__(kThreads2, 2)
k_spa_ffn2(const float* __restrict__ O, float* __restrict__ tok, const __grid_constant__ Tab512 tab,
           const uint8_t* __restrict__ wo, const uint8_t* __restrict__ w1a, const uint8_t* __restrict__ w1b,
           const uint8_t* __restrict__ w2a, const uint8_t* __restrict__ w2b, const uint8_t* __restrict__ wlin,
           float* __restrict__ out, const float* __restrict__ final_res, long long T, int P, int passes, Region fr,
           int otile) {
  // otile != 0: `O` holds the attention output as ready-made bf16 hi / lo operand tiles (k_spa_attn_mma): the producer warp
  // fetches this CTA's tile into the operand area with two bulk copies and the output projection runs in SS form - the row
  // owners have no phase 0 (no gather, no split, no operand stores), so the first GEMM starts as soon as the copy lands.
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t A = smem_u32(smem) + kCtlBytes;
  const uint32_t ring = A + 65536;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const uint32_t f2a_done = smem_u32(&ctl->aux[1]);  // completed by one tcgen05.commit (re-initialised below)
  const uint32_t o_full = smem_u32(&ctl->aux[2]);    // completed by the O tile's bulk copies (re-initialised below)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  cta_setup<kSpaNST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  if (tid == 0) {
    mbar_init(f2a_done, 1);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_o{wo, 128, 2}, g_1a{w1a, 128, 2}, g_2a{w2a, 128, 2}, g_1b{w1b, 128, 2}, g_2b{w2b, 128, 2},
      g_l{wlin, 64, 2};
  const bool fp32m = passes == 3;

  if (warp == kWarpProducer2) {
    RingState<kSpaNST> rs;  // consumption order
    if (otile) {
      pdl_wait();  // the tile is the previous kernel's output
      if (elect_one()) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(O) + (size_t)blockIdx.x * 65536u;
        mbar_arrive_expect_tx(o_full, fp32m ? 65536u : 32768u);
        bulk_g2s(A, src, 32768u, o_full);
        if (fp32m) bulk_g2s(A + 32768u, src + 32768u, 32768u, o_full);
      }
      __syncwarp();
    }
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_o, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1a, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2a, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1b, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2b, passes);
    ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_l, passes);
  } else if (warp == kWarpMma2) {
    RingState<kSpaNST> rs;
    uint32_t par = 0;
    auto wait_a = [&]() {
      mbar_wait(a_ready, par);
      par ^= 1;
      tc_fence_after();
    };
    auto gemm_ss = [&](const GemmPhase& g, uint32_t dcol) {   // A = Y1 operand in shared memory
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + dcol, true);
    };
    auto gemm_ts = [&](const GemmPhase& g, bool fresh) {      // A = TS operand in T, accumulator D
      ring_consume_mma_ts<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g, passes, tmem + 128, tmem + 192, tmem + 0, fresh);
    };
    auto done = [&]() { umma_commit_elected(mma_done); };
    if (otile) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      gemm_ss(g_o, 0); done();                       // D = O Wo^T, operand tile fetched by the producer
    } else {
      wait_a(); gemm_ts(g_o, true); done();          // D = O Wo^T
    }
    wait_a(); gemm_ss(g_1a, 0); done();              // D = Y1 W'1a^T
    wait_a(); gemm_ts(g_2a, true);                   // D = hidden_a W2a^T
    umma_commit_elected(f2a_done);
    mbar_wait(f2a_done, 0);                          // FFN2a has read its operand: T may become FFN1b's accumulator
    tc_fence_after();
    gemm_ss(g_1b, 128); done();                      // T = Y1 W'1b^T   (one commit covers FFN2a as well)
    wait_a(); gemm_ts(g_2b, false); done();          // D += hidden_b W2b^T
    wait_a(); gemm_ts(g_l, true); done();            // D[0,64) = Y2 Wlin^T
  } else {
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    const long long t = (long long)blockIdx.x * 128 + m;
    const bool ok = t < T;
    const unsigned PP = (unsigned)(P * P), RR = (unsigned)(fr.rn * fr.rn);
    auto locate = [&](unsigned tc, unsigned& vu, int& y, int& x) {
      vu = tc / RR;
      const int rem = (int)(tc - vu * RR);
      const int yy = rem / fr.rn;
      y = fr.r0 + yy;
      x = fr.r0 + rem - yy * fr.rn;
      return vu * PP + (unsigned)(y * P + x);
    };
    unsigned vu;
    int y, x;
    const unsigned tt = locate(ok ? (unsigned)t : 0u, vu, y, x);
    const long long v = vu;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* trow_g = tok + t32_off(tt, 16 * q, 32);  // own half of the token row, chunk stride 128 floats
    uint32_t par = 0;
    auto await = [&]() {
      mbar_wait(mma_done, par);
      par ^= 1;
      tc_fence_after();
    };
    auto publish_tmem = [&]() {  // operand written with tcgen05.st
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(a_ready);
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait();  // O / tok are the previous kernels' output
    LFT_TL(0);

    // phase 0: T <- O (planar gather of own heads 4q..4q+3)
    if (!otile) {
      float4 f[16];
      const float* ob = O + planar_off(v, 4 * q, y, 0, x, P);
      const long long hs = (long long)PP * 16, js = (long long)P * 4;  // head / piece strides in floats
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[4 * c + j] = (ok && !kFfnNoLoad) ? __ldg(reinterpret_cast<const float4*>(ob + c * hs + j * js)) : zero4;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, reinterpret_cast<const float*>(&f[4 * c]), fp32m);
      publish_tmem();
    }

    // phase 1: Y1 = tok + D (own half) -> bf16 hi/lo operand in shared memory (kept until the end); LN2 statistics
    float mean, rstd;
    {
      float4 tk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) tk[i] = (ok && !kFfnNoLoad) ? __ldg(reinterpret_cast<const float4*>(trow_g + 128 * i)) : zero4;  // in flight
      await();
      float yv[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, yv + 16 * c);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        yv[4 * i] += tk[i].x; yv[4 * i + 1] += tk[i].y; yv[4 * i + 2] += tk[i].z; yv[4 * i + 3] += tk[i].w;
      }
      if (!fp32m && ok) {  // bf16 mode: the hi part alone cannot carry the residual, keep the fp32 copy in `tok`
#pragma unroll
        for (int i = 0; i < 16; ++i)
          *reinterpret_cast<float4*>(trow_g + 128 * i) = make_float4(yv[4 * i], yv[4 * i + 1], yv[4 * i + 2], yv[4 * i + 3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) a_store16(A, 8 * q + 2 * c, m, yv + 16 * c, fp32m);
      pair_ln_stats<64>(yv, trow + 64 * q, trow + 64 * (1 - q), 1 + (warp & 3), mean, rstd);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);
    LFT_TL(2);
    const float mr = mean * rstd;
    const float4* u1 = reinterpret_cast<const float4*>(tab.v);        // [u_1 256 | c_1 256] (constant bank)
    const float4* c1 = reinterpret_cast<const float4*>(tab.v + 256);

    {  // speculative L2 prefetch of the O pieces and token row of the tile 2 CTAs x 148 SMs ahead
      const long long tn = t + 296ll * 128;
      if (tn < T) {
        unsigned vn;
        int yn, xn;
        const unsigned tnu = locate((unsigned)tn, vn, yn, xn);
        const float* tgn = tok + t32_off(tnu, 16 * q, 32);
        if (otile) {   // that CTA's operand tile: 512 lines of 128 bytes, two per row-owner thread
          const uint8_t* on = reinterpret_cast<const uint8_t*>(O) + ((size_t)blockIdx.x + 296u) * 65536u + (size_t)tid * 256u;
          prefetch_l2(on);
          prefetch_l2(on + 128);
        } else {
          const float* obn = O + planar_off(vn, 4 * q, yn, 0, xn, P);
          const long long hs = (long long)PP * 16, js = (long long)P * 4;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) prefetch_l2(obn + c * hs + j * js);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) prefetch_l2(tgn + 128 * i);
      }
    }

    // phases 2 / 3: hidden half (own 64 columns of the accumulator at column `acc`) -> relu(LN2-folded) -> T
    auto hidden = [&](uint32_t acc, int tab_off, bool in_place) {
      float hv[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + acc + 64 * q + 16 * c, hv + 16 * c);
      tmem_wait_ld();
      if (in_place) {  // the operand overwrites the accumulator: the partner thread must have read its half first
        tc_fence_before();
        pair_bar_sync(warp & 3);
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 uv = u1[tab_off + 16 * q + j], cv = c1[tab_off + 16 * q + j];
        hv[4 * j] = fmaxf(fmaf(rstd, hv[4 * j], fmaf(-mr, uv.x, cv.x)), 0.f);
        hv[4 * j + 1] = fmaxf(fmaf(rstd, hv[4 * j + 1], fmaf(-mr, uv.y, cv.y)), 0.f);
        hv[4 * j + 2] = fmaxf(fmaf(rstd, hv[4 * j + 2], fmaf(-mr, uv.z, cv.z)), 0.f);
        hv[4 * j + 3] = fmaxf(fmaf(rstd, hv[4 * j + 3], fmaf(-mr, uv.w, cv.w)), 0.f);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, hv + 16 * c, fp32m);
    };
    await();                 // FFN1a
    LFT_TL(3);
    hidden(0, 0, false);     // T held the O operand, consumed by the out-projection (complete: FFN1a was issued after it)
    publish_tmem();
    LFT_TL(4);
    await();                 // FFN2a (D) and FFN1b (T)
    LFT_TL(5);
    hidden(128, 32, true);
    publish_tmem();
    LFT_TL(6);

    // phase 4: Y2 = Y1 + D -> T.  Y1 from its hi/lo operand in shared memory (fp32 mode) or from the spilled row (bf16 mode).
    {
      float y1[64];
      if (fp32m) {
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          uint4 hi, lo;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(A + (8 * q + kc) * kLbo + m * 16));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(A + 32768 + (8 * q + kc) * kLbo + m * 16));
          const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // bf16 pair -> two fp32 (a bf16 is the top half of an fp32)
            y1[8 * kc + 2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
            y1[8 * kc + 2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 f = ok ? *reinterpret_cast<const float4*>(trow_g + 128 * i) : zero4;
          y1[4 * i] = f.x; y1[4 * i + 1] = f.y; y1[4 * i + 2] = f.z; y1[4 * i + 3] = f.w;
        }
      }
      await();               // FFN2b
      LFT_TL(7);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float d[16];
        tmem_ld16(trow + 64 * q + 16 * c, d);
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] += y1[16 * c + j];
        a_tmem_store16(trow + 128, trow + 192, 64 * q + 16 * c, d, fp32m);
      }
    }
    publish_tmem();
    LFT_TL(8);

    // phase 5: out = D[0,64) (+ global residual), own 32 columns
    float4 r4[8];
    if (final_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r4[i] = ok ? __ldg(reinterpret_cast<const float4*>(final_res + t32_off(tt, 8 * q + i, 16))) : zero4;
    }
    await();
    LFT_TL(9);
    {
      float d[32];
      tmem_ld16_nowait(trow + 32 * q, d);
      tmem_ld16_nowait(trow + 32 * q + 16, d + 16);
      tmem_wait_ld();
      if (ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o4 = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
          if (final_res) { o4.x += r4[i].x; o4.y += r4[i].y; o4.z += r4[i].z; o4.w += r4[i].w; }
          *reinterpret_cast<float4*>(out + t32_off(tt, 8 * q + i, 16)) = o4;
        }
      }
    }
    LFT_TL(10);
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
}

int debug_timeline_ring_embed(long long* out) {  // out[40*3]
#ifdef LFT_TIMELINE
  CUDA_TRY(cudaMemcpyFromSymbol(out, g_ring_tl, sizeof(long long) * 120));
  return 0;
#else
  (void)out;
  return fail(LFT_ERR_STATE, "library built without -DLFT_TIMELINE");
#endif
}

int debug_timeline_embed(long long* out) {
#ifdef LFT_TIMELINE
  CUDA_TRY(cudaMemcpyFromSymbol(out, g_tl2, sizeof(long long) * 64));
  return 0;
#else
  (void)out;
  return fail(LFT_ERR_STATE, "library built without -DLFT_TIMELINE");
#endif
}

int debug_timeline_spa(long long* out) {
#ifdef LFT_TIMELINE
  CUDA_TRY(cudaMemcpyFromSymbol(out, g_tl, sizeof(long long) * 64));
  return 0;
#else
  (void)out;
  return fail(LFT_ERR_STATE, "library built without -DLFT_TIMELINE");
#endif
}

int configure_spa() {
  CUDA_TRY(cudaFuncSetAttribute(k_spa_embed_qkv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_embed<false>()));
  CUDA_TRY(cudaFuncSetAttribute(k_spa_embed_qkv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_embed<true>()));
#ifdef LFT_FFN_V1
  CUDA_TRY(cudaFuncSetAttribute(k_spa_ffn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
#endif
  CUDA_TRY(cudaFuncSetAttribute(k_spa_ffn2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
#ifdef LFT_ATTN_V1
  CUDA_TRY(cudaFuncSetAttribute(k_spa_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attn(64)));
#endif
  CUDA_TRY(cudaFuncSetAttribute(k_spa_attn_mma<true, kOTile != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attn_mma(64)));
  CUDA_TRY(cudaFuncSetAttribute(k_spa_attn_mma<false, kOTile != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_attn_mma(64)));
  return 0;
}

// altblock[layer].spa_trans: in [T,64] -> out [T,64].  `need`: the pixels of every view the caller needs `out` on (full view:
// {0, P}); `in` must be valid on `need` grown by 3 pixels (2 for the attention window + 1 for the 3x3 token embedding), clipped
// to the view.
static Region grow(Region r, int k, int P) {
  const int lo = r.r0 - k < 0 ? 0 : r.r0 - k;
  const int hi = r.r0 + r.rn + k > P ? P : r.r0 + r.rn + k;
  return Region{lo, hi - lo};
}

int run_spa(Handle* h, int layer, const float* in, float* out, const float* final_res, Workspace& w, int B, int P,
            Region need, cudaStream_t st) {
  const int A = h->cfg.ang_res;
  const long long V = (long long)B * A * A;
  const Layer& L = h->layer[layer];
  const Region kv = grow(need, 2, P);   // K / V (and tok, Q) are stored here
  const Region e = grow(need, 3, P);    // conv inputs
  int rc;
  {
    const long long G = V * (e.rn + 1) * (e.rn + 1);
    Tab512 tq;
    memcpy(tq.v, L.s_tab[h->mode()].data(), sizeof(tq.v));  // [u_q | u_k | c_q | c_k]
    Scope sc(h, K_SPA_QKV, st, V * kv.rn * kv.rn);
    const unsigned ntiles = (unsigned)((G + 127) / 128);
    const unsigned pg = ntiles < 2u * h->num_sms ? ntiles : 2u * h->num_sms;  // persistent: two CTAs per SM
    if (P <= ConvGeom<false>::kMaxP) {
      auto kern = k_spa_embed_qkv<false>;
      LFT_LAUNCH(h, kern, pg, kThreads2, smem_embed<false>(), st, in, L.s_wmlp, L.s_pe, L.s_pev[h->mode()], tq, L.s_wq, L.s_wk,
                 L.s_wv, w.tok, w.q, w.k, w.v, (int)V, P, h->passes(), (int)ntiles, e, kv, need);
    } else {
      auto kern = k_spa_embed_qkv<true>;
      LFT_LAUNCH(h, kern, pg, kThreads2, smem_embed<true>(), st, in, L.s_wmlp, L.s_pe, L.s_pev[h->mode()], tq, L.s_wq, L.s_wk,
                 L.s_wv, w.tok, w.q, w.k, w.v, (int)V, P, h->passes(), (int)ntiles, e, kv, need);
    }
    if ((rc = sc.finish())) return rc;
  }
  {
    Scope sc(h, K_SPA_ATTN, st, V * need.rn * need.rn);
#ifdef LFT_ATTN_V1
    const int nblk = (need.rn + kAttnRB - 1) / kAttnRB;
    LFT_LAUNCH(h, k_spa_attn, (unsigned)(V * 8 * nblk), kAttnThreads, smem_attn(P), st, (const float*)w.q, (const float*)w.k,
               (const float*)w.v, w.o, P, need);
#else
    {
      const int nb = ((need.r0 + need.rn) - (need.r0 & ~3) + kAttnRB - 1) / kAttnRB;  // attn_mma_nblk
      const long long items = V * 8 * nb;
      if (items >= (1ll << 31)) return fail(LFT_ERR_ARG, "k_spa_attn_mma: too many work items");
      const long long maxg = 2ll * h->num_sms;   // persistent: two CTAs per SM; a multiple of nb so that a CTA keeps its row block
      const unsigned pg = (unsigned)(items <= maxg ? items : (maxg / nb) * nb);
      auto launch = [&](auto kern) {
        LFT_LAUNCH(h, kern, pg, kAttnMmaThreads, smem_attn_mma(P), st, (const float*)w.q, (const float*)w.k, (const float*)w.v,
                   w.o, P, need, (int)items);
      };
      if (h->passes() == 3) launch(k_spa_attn_mma<true, kOTile != 0>);
      else launch(k_spa_attn_mma<false, kOTile != 0>);
    }
#endif
    if ((rc = sc.finish())) return rc;
  }
  {
    Tab512 tf;
    memcpy(tf.v, L.s_tab[h->mode()].data() + 512, sizeof(tf.v));  // [u_1 256 | c_1 256]
    const long long T = V * need.rn * need.rn;
    Scope sc(h, K_SPA_FFN, st, T);
#ifdef LFT_FFN_V1
    LFT_LAUNCH(h, k_spa_ffn, (unsigned)((T + 127) / 128), kThreads2, kSmemSpa, st, (const float*)w.o, w.tok, tf, L.s_wo, L.s_w1a,
               L.s_w1b, L.s_w2a, L.s_w2b, L.s_wlin, out, final_res, T, P, h->passes(), need);
#else
    LFT_LAUNCH(h, k_spa_ffn2, (unsigned)((T + 127) / 128), kThreads2, kSmemSpa, st, (const float*)w.o, w.tok, tf, L.s_wo, L.s_w1a,
               L.s_w1b, L.s_w2a, L.s_w2b, L.s_wlin, out, final_res, T, P, h->passes(), need, kOTile);
#endif
    if ((rc = sc.finish())) return rc;
  }
  return 0;
}

}  // namespace lft
