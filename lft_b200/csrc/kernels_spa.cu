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

namespace lft {

constexpr int kSpaNST = 3;
constexpr uint32_t kSpaStage = 128 * 128;
constexpr size_t kSmemSpa = kCtlBytes + 65536 + kSpaNST * kSpaStage;
constexpr uint32_t kLbo = 128 * 16;

LFT_DEVINL long long planar_off(long long v, int head, int y, int j, int x, int P) {
  return ((((v * 8 + head) * P + y) * 4 + j) * (long long)P + x) * 4;
}

// write 16 accumulator columns (= one head) of one token into the planar layout
LFT_DEVINL void planar_store16(float* base, long long v, int head, int y, int x, int P, const float* d) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(base + planar_off(v, head, y, j, x, P)) =
        make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
}

// split 16 fp32 values into two k-chunks (kc0, kc0+1) of the K=128 A operand (hi at A, lo at A+32K)
LFT_DEVINL void a_store16(uint32_t A, int kc0, int m, const float* x) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 hi, lo;
    split8(x + 8 * j, hi, lo);
    st_shared_v4(A + (kc0 + j) * kLbo + m * 16, hi);
    st_shared_v4(A + 32768 + (kc0 + j) * kLbo + m * 16, lo);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads2, 2)
k_spa_embed_qkv(const float* __restrict__ feat, const uint8_t* __restrict__ wmlp, const float* __restrict__ pe,
                const float* __restrict__ pev, const float* __restrict__ tab, const uint8_t* __restrict__ wq,
                const uint8_t* __restrict__ wk, const uint8_t* __restrict__ wv, float* __restrict__ tok,
                float* __restrict__ Q, float* __restrict__ K, float* __restrict__ Vv, int V, int P, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t U = smem_u32(smem) + kCtlBytes;
  const uint32_t c_hi = U, c_lo = U + kConvRows * 128;  // conv staging (51.5 KB), dead after the conv MMAs
  const uint32_t A = U;                                  // K=128 operand: hi [0,32K), lo [32K,64K)
  const uint32_t ring = U + 65536;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P1 = P + 1;
  const long long VS = (long long)P1 * P1;
  const long long G = (long long)V * VS;
  const long long g0 = (long long)blockIdx.x * 128;
  cta_setup<kSpaNST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_c{wmlp, 128, 9}, g_q{wq, 128, 2}, g_k{wk, 128, 2}, g_v{wv, 128, 2};

  if (warp == kWarpProducer2) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_c, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_q, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_k, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_v, passes);
    }
  } else if (warp == kWarpMma2) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      mbar_wait(a_ready, 0);
      tc_fence_after();
      auto shift = [P1](uint32_t t) { return ((int)(t / 3) - 1) * P1 + ((int)(t % 3) - 1); };
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_c, passes, c_hi + kConvOff * 16,
                                c_lo + kConvOff * 16, kConvRows * 16, 0, shift, tmem, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 1);
      tc_fence_after();
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_q, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 0, true);
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_k, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 128, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 0);
      tc_fence_after();
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_v, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 0, true);
      umma_commit(mma_done);
    }
  } else {
    conv_stage_window(feat, c_hi, c_lo, g0, G, VS, P, tid);
    fence_proxy_async_smem();
    mbar_arrive(a_ready);

    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    const long long g = g0 + m;
    bool ok = false;
    long long v = 0;
    int y = 0, x = 0;
    if (g < G) {
      v = g / VS;
      const int qq = (int)(g - v * VS);
      y = qq / P1;
      x = qq - y * P1;
      ok = (y < P && x < P);
    }
    if (!ok) { v = 0; y = 0; x = 0; }
    const int PP = P * P;
    const int p = y * P + x;
    const long long token = (v * P + y) * P + x;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);

    // ---- phase 1: tok (own 64 columns) -> global; z = tok + PE_s -> A operand and LN statistics.
    // (Q, K = LN(z) W^T via the folded epilogue; V = tok Wv^T = z Wv^T - PE_s Wv^T.)  PE is fetched before the wait.
    float mean, rstd;
    {
      float z[64];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(pe) + (long long)(16 * q + i) * PP + p);  // [chunk][p][4]
        z[4 * i] = b.x; z[4 * i + 1] = b.y; z[4 * i + 2] = b.z; z[4 * i + 3] = b.w;
      }
      mbar_wait(mma_done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[16];
        tmem_ld16(trow + 64 * q + 16 * c, t);
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(tok + t32_off(token, 16 * q + 4 * c + i, 32)) =
                make_float4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) z[16 * c + i] += t[i];
        a_store16(A, 8 * q + 2 * c, m, z + 16 * c);
      }
      pair_ln_stats<64>(z, trow + 128 + 4 * q, trow + 128 + 4 * (1 - q), 1 + (warp & 3), mean, rstd);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);

    // ---- phase 2: Q, K epilogues (affine LN correction), V MMAs start as soon as Q has been read
    const float4* tab4 = reinterpret_cast<const float4*>(tab);  // [u_q | u_k | c_q | c_k] x 128
    const float mr = mean * rstd;
    mbar_wait(mma_done, 1);
    tc_fence_after();
#pragma unroll 2
    for (int c = 0; c < 4; ++c) {
      float d[16];
      const int col = 64 * q + 16 * c;
      tmem_ld16(trow + col, d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 uv = __ldg(tab4 + col / 4 + j), cv = __ldg(tab4 + 64 + col / 4 + j);
        d[4 * j] = fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x));
        d[4 * j + 1] = fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y));
        d[4 * j + 2] = fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z));
        d[4 * j + 3] = fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w));
      }
      if (ok) planar_store16(Q, v, 4 * q + c, y, x, P, d);
    }
    tc_fence_before();
    mbar_arrive(a_ready);
#pragma unroll 2
    for (int c = 0; c < 4; ++c) {
      float d[16];
      const int col = 64 * q + 16 * c;
      tmem_ld16(trow + 128 + col, d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 uv = __ldg(tab4 + 32 + col / 4 + j), cv = __ldg(tab4 + 96 + col / 4 + j);
        d[4 * j] = fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x));
        d[4 * j + 1] = fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y));
        d[4 * j + 2] = fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z));
        d[4 * j + 3] = fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w));
      }
      if (ok) planar_store16(K, v, 4 * q + c, y, x, P, d);
    }
    // ---- phase 3: V = D - PE_s Wv^T  (table prefetched before the wait)
    {
      float4 pv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pv[i] = __ldg(reinterpret_cast<const float4*>(pev) + (long long)(16 * q + i) * PP + p);
      mbar_wait(mma_done, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float d[16];
        tmem_ld16(trow + 64 * q + 16 * c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d[4 * j] -= pv[4 * c + j].x; d[4 * j + 1] -= pv[4 * c + j].y;
          d[4 * j + 2] -= pv[4 * c + j].z; d[4 * j + 3] -= pv[4 * c + j].w;
        }
        if (ok) planar_store16(Vv, v, 4 * q + c, y, x, P, d);
      }
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
}

// ------------------------------------------------------------------------------------------------
// Window attention. One thread = one head of TWO vertically adjacent queries (y0, x), (y0+1, x): the
// 6 x 5 keys their windows cover are read once (30 instead of 50 key reads).
LFT_DEVINL float dot16(const float* q, const float4& a, const float4& b, const float4& c, const float4& d) {
  float s0 = q[0] * a.x, s1 = q[4] * b.x, s2 = q[8] * c.x, s3 = q[12] * d.x;
  s0 = fmaf(q[1], a.y, s0); s1 = fmaf(q[5], b.y, s1); s2 = fmaf(q[9], c.y, s2); s3 = fmaf(q[13], d.y, s3);
  s0 = fmaf(q[2], a.z, s0); s1 = fmaf(q[6], b.z, s1); s2 = fmaf(q[10], c.z, s2); s3 = fmaf(q[14], d.z, s3);
  s0 = fmaf(q[3], a.w, s0); s1 = fmaf(q[7], b.w, s1); s2 = fmaf(q[11], c.w, s2); s3 = fmaf(q[15], d.w, s3);
  return (s0 + s1) + (s2 + s3);
}
LFT_DEVINL void axpy16(float* o, float p, const float4& a, const float4& b, const float4& c, const float4& d) {
  o[0] = fmaf(p, a.x, o[0]); o[1] = fmaf(p, a.y, o[1]); o[2] = fmaf(p, a.z, o[2]); o[3] = fmaf(p, a.w, o[3]);
  o[4] = fmaf(p, b.x, o[4]); o[5] = fmaf(p, b.y, o[5]); o[6] = fmaf(p, b.z, o[6]); o[7] = fmaf(p, b.w, o[7]);
  o[8] = fmaf(p, c.x, o[8]); o[9] = fmaf(p, c.y, o[9]); o[10] = fmaf(p, c.z, o[10]); o[11] = fmaf(p, c.w, o[11]);
  o[12] = fmaf(p, d.x, o[12]); o[13] = fmaf(p, d.y, o[13]); o[14] = fmaf(p, d.z, o[14]); o[15] = fmaf(p, d.w, o[15]);
}

// CTA = (view, head, block of kAttnRB query rows): the K and V planes of rows [r0-2, r0+RB+2) are contiguous
// in the planar layout and are staged in shared memory with two bulk copies (TMA engine).  Softmax is
// evaluated online, one key row at a time (5 scores per query live at once).
constexpr int kAttnRB = 8;
constexpr size_t kSmemAttn = 2 * (kAttnRB + 4) * 4 * 32 * 16 + 16;

__global__ void __launch_bounds__(128)
k_spa_attn(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ Vv,
           float* __restrict__ O, int P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nblk = (P + kAttnRB - 1) / kAttnRB;
  const int rb = blockIdx.x % nblk;
  const int head = (blockIdx.x / nblk) & 7;
  const long long v = blockIdx.x / (nblk * 8);
  const int r0 = rb * kAttnRB;
  const int ys = max(r0 - 2, 0), ye = min(r0 + kAttnRB + 2, P);   // staged key rows [ys, ye)
  const uint32_t rowbytes = (uint32_t)P * 64;                      // one (y) plane: 4 pieces x P x 16 B
  const uint32_t nbytes = (uint32_t)(ye - ys) * rowbytes;
  const uint32_t bar = smem_u32(smem);
  uint8_t* ks = smem + 16;
  uint8_t* vs = ks + (kAttnRB + 4) * rowbytes;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 2 * nbytes);
    const long long src = planar_off(v, head, ys, 0, 0, P);
    bulk_g2s(smem_u32(ks), K + src, nbytes, bar);
    bulk_g2s(smem_u32(vs), Vv + src, nbytes, bar);
  }
  const int x = threadIdx.x % P;
  const int rp = threadIdx.x / P;
  const int y0 = r0 + 2 * rp;
  const bool active = (rp < kAttnRB / 2) && (y0 < P);
  const bool two = active && (y0 + 1) < P;
  const long long rowstride = (long long)P * 16;  // floats between consecutive y
  const int jstride = P * 4;                      // floats between the 4 pieces of one (y, x)
  const long long base = planar_off(v, head, 0, 0, x, P);
  const float qs = 0.25f * 1.4426950408889634f;   // log2(e)/sqrt(16): softmax through exp2
  float q0[16], q1[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 f = active ? __ldg(reinterpret_cast<const float4*>(Q + base + y0 * rowstride + j * jstride))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    q0[4 * j] = f.x * qs; q0[4 * j + 1] = f.y * qs; q0[4 * j + 2] = f.z * qs; q0[4 * j + 3] = f.w * qs;
    const float4 g = two ? __ldg(reinterpret_cast<const float4*>(Q + base + (y0 + 1) * rowstride + j * jstride))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    q1[4 * j] = g.x * qs; q1[4 * j + 1] = g.y * qs; q1[4 * j + 2] = g.z * qs; q1[4 * j + 3] = g.w * qs;
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o0[16], o1[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) { o0[e] = 0.f; o1[e] = 0.f; }
  mbar_wait(bar, 0);
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
          const float4 k0 = *reinterpret_cast<const float4*>(krow + dx * 4);
          const float4 k1 = *reinterpret_cast<const float4*>(krow + dx * 4 + jstride);
          const float4 k2 = *reinterpret_cast<const float4*>(krow + dx * 4 + 2 * jstride);
          const float4 k3 = *reinterpret_cast<const float4*>(krow + dx * 4 + 3 * jstride);
          if (kr <= 4) { a0[dx + 2] = dot16(q0, k0, k1, k2, k3); n0 = fmaxf(n0, a0[dx + 2]); }
          if (kr >= 1) { a1[dx + 2] = dot16(q1, k0, k1, k2, k3); n1 = fmaxf(n1, a1[dx + 2]); }
        }
      }
      if (kr <= 4) {
        const float sc = fast_exp2(m0 - n0);
        m0 = n0;
        l0 *= sc;
#pragma unroll
        for (int e = 0; e < 16; ++e) o0[e] *= sc;
      }
      if (kr >= 1 && two) {
        const float sc = fast_exp2(m1 - n1);
        m1 = n1;
        l1 *= sc;
#pragma unroll
        for (int e = 0; e < 16; ++e) o1[e] *= sc;
      }
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int kx = x + dx;
        if (kx >= 0 && kx < P) {
          const float4 v0 = *reinterpret_cast<const float4*>(vrow + dx * 4);
          const float4 v1 = *reinterpret_cast<const float4*>(vrow + dx * 4 + jstride);
          const float4 v2 = *reinterpret_cast<const float4*>(vrow + dx * 4 + 2 * jstride);
          const float4 v3 = *reinterpret_cast<const float4*>(vrow + dx * 4 + 3 * jstride);
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
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(O + base + y0 * rowstride + j * jstride) =
          make_float4(o0[4 * j] * i0, o0[4 * j + 1] * i0, o0[4 * j + 2] * i0, o0[4 * j + 3] * i0);
    if (two) {
      const float i1 = 1.f / l1;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(O + base + (y0 + 1) * rowstride + j * jstride) =
            make_float4(o1[4 * j] * i1, o1[4 * j + 1] * i1, o1[4 * j + 2] * i1, o1[4 * j + 3] * i1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads2, 2)
k_spa_ffn(const float* __restrict__ O, float* __restrict__ tok, const float* __restrict__ tab,
          const uint8_t* __restrict__ wo, const uint8_t* __restrict__ w1a, const uint8_t* __restrict__ w1b,
          const uint8_t* __restrict__ w2a, const uint8_t* __restrict__ w2b, const uint8_t* __restrict__ wlin,
          float* __restrict__ out, const float* __restrict__ final_res, long long T, int P, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t A = smem_u32(smem) + kCtlBytes;
  const uint32_t ring = A + 65536;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cta_setup<kSpaNST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_o{wo, 128, 2}, g_1a{w1a, 128, 2}, g_2a{w2a, 128, 2}, g_1b{w1b, 128, 2}, g_2b{w2b, 128, 2},
      g_l{wlin, 64, 2};

  if (warp == kWarpProducer2) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_o, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1a, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2a, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1b, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2b, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_l, passes);
    }
  } else if (warp == kWarpMma2) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      uint32_t par = 0;
      auto step = [&](const GemmPhase& g, uint32_t dcol, bool fresh) {
        mbar_wait(a_ready, par);
        par ^= 1;
        tc_fence_after();
        ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                  tmem + dcol, fresh);
        umma_commit(mma_done);
      };
      step(g_o, 0, true);      // D[0,128)    = O Wo^T
      step(g_1a, 0, true);     // D[0,128)    = Y1 W'1[0:128]^T
      step(g_2a, 128, false);  // S[128,256) += relu(.) W2[:,0:128]^T     (S was initialised to Y1)
      step(g_1b, 0, true);     // D[0,128)    = Y1 W'1[128:256]^T
      step(g_2b, 128, false);  // S          += relu(.) W2[:,128:256]^T   -> S = Y2
      step(g_l, 0, true);      // D[0,64)     = Y2 Wlin^T
    }
  } else {
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    const long long t = (long long)blockIdx.x * 128 + m;
    const bool ok = t < T;
    const long long tt = ok ? t : 0;
    const int PP = P * P;
    const long long v = tt / PP;
    const int p = (int)(tt - v * PP);
    const int y = p / P, x = p - y * P;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* trow_g = tok + t32_off(tt, 16 * q, 32);  // own half of the token row, chunk stride 128 floats (Y1 is spilled here)
    uint32_t par = 0;
    auto publish = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    };
    auto await = [&]() {
      mbar_wait(mma_done, par);
      par ^= 1;
      tc_fence_after();
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // phase 0: A <- O (planar gather of own heads 4q..4q+3)
    {
      float4 f[16];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          f[4 * c + j] = ok ? __ldg(reinterpret_cast<const float4*>(O + planar_off(v, 4 * q + c, y, j, x, P))) : zero4;
#pragma unroll
      for (int c = 0; c < 4; ++c) a_store16(A, 8 * q + 2 * c, m, reinterpret_cast<const float*>(&f[4 * c]));
    }
    publish();

    // phase 1: Y1 = tok + D (own half): -> global (spill), TMEM S, raw A; LN2 statistics
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
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(trow + 128 + 64 * q + 16 * c, yv + 16 * c);
      if (ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          *reinterpret_cast<float4*>(trow_g + 128 * i) = make_float4(yv[4 * i], yv[4 * i + 1], yv[4 * i + 2], yv[4 * i + 3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) a_store16(A, 8 * q + 2 * c, m, yv + 16 * c);
      pair_ln_stats<64>(yv, trow + 64 * q, trow + 64 * (1 - q), 1 + (warp & 3), mean, rstd);
    }
    publish();
    const float mr = mean * rstd;

    // phases 2..5: hidden halves (LN2 folded: hidden = relu(rstd*D - rstd*mean*u1 + c1))
    for (int half = 0; half < 2; ++half) {
      float4 y1[16];
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) y1[i] = ok ? *reinterpret_cast<const float4*>(trow_g + 128 * i) : zero4;  // prefetch Y1
      }
      await();  // FFN1 half done: D[0,128)
      const float4* u1 = reinterpret_cast<const float4*>(tab + 512 + 128 * half + 64 * q);
      const float4* c1 = reinterpret_cast<const float4*>(tab + 768 + 128 * half + 64 * q);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float d[16];
        tmem_ld16(trow + 64 * q + 16 * c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 uv = __ldg(u1 + 4 * c + j), cv = __ldg(c1 + 4 * c + j);
          d[4 * j] = fmaxf(fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x)), 0.f);
          d[4 * j + 1] = fmaxf(fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y)), 0.f);
          d[4 * j + 2] = fmaxf(fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z)), 0.f);
          d[4 * j + 3] = fmaxf(fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w)), 0.f);
        }
        a_store16(A, 8 * q + 2 * c, m, d);
      }
      publish();
      await();  // FFN2 half accumulated into S
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) a_store16(A, 8 * q + 2 * c, m, reinterpret_cast<const float*>(&y1[4 * c]));
        publish();
      }
    }
    // phase 6: A <- Y2 = S (own half)
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 64 * q + 16 * c, z);
      a_store16(A, 8 * q + 2 * c, m, z);
    }
    publish();
    // phase 7: out = D[0,64) (+ global residual), own 32 columns
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
}

int configure_spa() {
  CUDA_TRY(cudaFuncSetAttribute(k_spa_embed_qkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
  CUDA_TRY(cudaFuncSetAttribute(k_spa_ffn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
  CUDA_TRY(cudaFuncSetAttribute(k_spa_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAttn));
  return 0;
}

// altblock[layer].spa_trans: in [T,64] -> out [T,64]
int run_spa(Handle* h, int layer, const float* in, float* out, const float* final_res, Workspace& w, int B, int P,
            cudaStream_t st) {
  const int A = h->cfg.ang_res;
  const long long V = (long long)B * A * A;
  const long long T = V * P * P;
  const Layer& L = h->layer[layer];
  int rc;
  {
    const long long G = V * (P + 1) * (P + 1);
    Scope sc(h, K_SPA_QKV, st);
    k_spa_embed_qkv<<<(unsigned)((G + 127) / 128), kThreads2, kSmemSpa, st>>>(
        in, L.s_wmlp, L.s_pe, L.s_pev, L.s_tab, L.s_wq, L.s_wk, L.s_wv, w.tok, w.q, w.k, w.v, (int)V, P, h->passes());
    if ((rc = sc.finish())) return rc;
  }
  {
    Scope sc(h, K_SPA_ATTN, st);
    const int nblk = (P + kAttnRB - 1) / kAttnRB;
    k_spa_attn<<<(unsigned)(V * 8 * nblk), 128, kSmemAttn, st>>>(w.q, w.k, w.v, w.o, P);
    if ((rc = sc.finish())) return rc;
  }
  {
    Scope sc(h, K_SPA_FFN, st);
    k_spa_ffn<<<(unsigned)((T + 127) / 128), kThreads2, kSmemSpa, st>>>(w.o, w.tok, L.s_tab, L.s_wo, L.s_w1a, L.s_w1b,
                                                                        L.s_w2a, L.s_w2b, L.s_wlin, out, final_res, T, P,
                                                                        h->passes());
    if ((rc = sc.finish())) return rc;
  }
  return 0;
}

}  // namespace lft
