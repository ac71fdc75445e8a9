// Fused AngTrans (model/LFT.py:194-238): for every pixel, MHSA over its N=A*A angular tokens + FFN.
//   Xn = LN(X + PE_a); Q,K = Xn Wq^T, Xn Wk^T; V = X Wv^T (raw tokens!); per head (hd=8) softmax(QK^T/sqrt 8) V;
//   X1 = X + O Wo^T;  X2 = X1 + W2 relu(W1 LN2(X1)).
// One CTA = 128 accumulator rows = floor(128/N) pixels x N views.  The (b c a h w) <-> (a, b h w, c)
// permutes of the reference (LFT.py:216-223) are address arithmetic in the row loads/stores.
// Projections/FFN run on tcgen05 (A operand built in smem by the row-owner threads, weights streamed
// through the ring); the 25x25 (hd 8) attention core is CUDA-core work with K/V shared through smem.
#include "host.h"
#include "kernels.cuh"

namespace lft {

constexpr int kAngNST = 3;
constexpr uint32_t kAngStage = 128 * 128;  // largest slab: N=128 rows
constexpr size_t kSmemAng = kCtlBytes + 65536 + kAngNST * kAngStage;

__global__ void __launch_bounds__(kThreads, 2)
k_ang(const float* __restrict__ in, float* __restrict__ out, const uint8_t* __restrict__ wqk,
      const uint8_t* __restrict__ wv, const uint8_t* __restrict__ wo, const uint8_t* __restrict__ w1,
      const uint8_t* __restrict__ w2, const float* __restrict__ ln, const float* __restrict__ pe, int N, int PP,
      long long npix, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t R1 = s_base + kCtlBytes;  // 32 KB
  const uint32_t R2 = R1 + 32768;          // 32 KB
  const uint32_t ring = R2 + 32768;
  uint8_t* r2_ptr = smem + kCtlBytes + 32768;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t LBO = 128 * 16;

  cta_setup<kAngNST>(ctl, warp, lane, 128, 256);
  const uint32_t tmem = ctl->tmem;

  const GemmPhase g_qk{wqk, 128, 1}, g_v{wv, 64, 1}, g_o{wo, 64, 1}, g_1{w1, 128, 1}, g_2{w2, 64, 2};

  if (warp == kWarpProducer) {
    if (lane == 0) {
      RingState<kAngNST> rs;
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_qk, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_v, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_o, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_1, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_2, passes);
    }
  } else if (warp == kWarpMma) {
    if (lane == 0) {
      RingState<kAngNST> rs;
      mbar_wait(a_ready, 0);
      tc_fence_after();
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_qk, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                                tmem + 0, true);
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_v, passes, R2, R2 + 16384, LBO, 0, NoShift{},
                                tmem + 128, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 1);
      tc_fence_after();
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_o, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                                tmem + 0, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 0);
      tc_fence_after();
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_1, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                                tmem + 0, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 1);
      tc_fence_after();
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_2, passes, R1, R2, LBO, 8 * LBO, NoShift{},
                                tmem + 128, true);
      umma_commit(mma_done);
    }
  } else {
    // ------------------------------------------------------------ row owner: m = tid
    const int m = tid;
    const int PPT = 128 / N;
    const int pl = m / N, a = m - pl * N;
    const long long gp = (long long)blockIdx.x * PPT + pl;
    const bool rowok = (pl < PPT) && (gp < npix);
    long long tok = 0;
    if (rowok) {
      const long long b = gp / PP;
      const int p = (int)(gp - b * PP);
      tok = (b * N + a) * PP + p;
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);

    // ---- phase 0: load X, stash, LN1(X+PE) -> R1, X -> R2
    {
      float x[64];
      const float4* src = reinterpret_cast<const float4*>(in + tok * 64);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 f = rowok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(trow + 192 + 16 * c, x + 16 * c);
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        uint4 hi, lo;
        split8(x + 8 * kc, hi, lo);
        st_shared_v4(R2 + kc * LBO + m * 16, hi);
        st_shared_v4(R2 + 16384 + kc * LBO + m * 16, lo);
      }
      const float4* pep = reinterpret_cast<const float4*>(pe + (rowok ? a : 0) * 64);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 f = __ldg(pep + i);
        x[4 * i] += f.x; x[4 * i + 1] += f.y; x[4 * i + 2] += f.z; x[4 * i + 3] += f.w;
        sum += (x[4 * i] + x[4 * i + 1]) + (x[4 * i + 2] + x[4 * i + 3]);
      }
      const float mean = sum * (1.f / 64.f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { const float d = x[i] - mean; var = fmaf(d, d, var); }
      const float rstd = rsqrtf(var * (1.f / 64.f) + 1e-5f);
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (x[8 * kc + e] - mean) * rstd * __ldg(ln + 8 * kc + e) + __ldg(ln + 64 + 8 * kc + e);
        uint4 hi, lo;
        split8(y, hi, lo);
        st_shared_v4(R1 + kc * LBO + m * 16, hi);
        st_shared_v4(R1 + 16384 + kc * LBO + m * 16, lo);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    // ---- phase 1: attention (two head-halves; K,V of 4 heads at a time in R2)
    mbar_wait(mma_done, 0);
    tc_fence_after();
    const float scale = 0.35355339059327373f;  // 1/sqrt(8)
    for (int half = 0; half < 2; ++half) {
      {
        float kv[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // K cols 64+32*half .. +32
          tmem_ld16(trow + 64 + 32 * half + 16 * c, kv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = 4 * c + j;
            *reinterpret_cast<float4*>(r2_ptr + m * 128 + ((ch ^ (m & 7)) * 16)) =
                make_float4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // V cols 128+32*half .. +32
          tmem_ld16(trow + 128 + 32 * half + 16 * c, kv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = 4 * c + j;
            *reinterpret_cast<float4*>(r2_ptr + 16384 + m * 128 + ((ch ^ (m & 7)) * 16)) =
                make_float4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
          }
        }
      }
      named_bar_sync(1, 128);
      float q[32];
      tmem_ld16(trow + 32 * half, q);
      tmem_ld16(trow + 32 * half + 16, q + 16);
      float o[32];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        float l = 1.f;
        if (pl < PPT) {
          const int r0 = pl * N;
          float mx = -INFINITY;
          for (int t = 0; t < N; ++t) {
            const int r = r0 + t;
            const float4 k0 = *reinterpret_cast<const float4*>(r2_ptr + r * 128 + (((2 * hh) ^ (r & 7)) * 16));
            const float4 k1 = *reinterpret_cast<const float4*>(r2_ptr + r * 128 + (((2 * hh + 1) ^ (r & 7)) * 16));
            float s = q[8 * hh] * k0.x;
            s = fmaf(q[8 * hh + 1], k0.y, s); s = fmaf(q[8 * hh + 2], k0.z, s); s = fmaf(q[8 * hh + 3], k0.w, s);
            s = fmaf(q[8 * hh + 4], k1.x, s); s = fmaf(q[8 * hh + 5], k1.y, s); s = fmaf(q[8 * hh + 6], k1.z, s);
            s = fmaf(q[8 * hh + 7], k1.w, s);
            mx = fmaxf(mx, s);
          }
          l = 0.f;
          for (int t = 0; t < N; ++t) {
            const int r = r0 + t;
            const float4 k0 = *reinterpret_cast<const float4*>(r2_ptr + r * 128 + (((2 * hh) ^ (r & 7)) * 16));
            const float4 k1 = *reinterpret_cast<const float4*>(r2_ptr + r * 128 + (((2 * hh + 1) ^ (r & 7)) * 16));
            float s = q[8 * hh] * k0.x;
            s = fmaf(q[8 * hh + 1], k0.y, s); s = fmaf(q[8 * hh + 2], k0.z, s); s = fmaf(q[8 * hh + 3], k0.w, s);
            s = fmaf(q[8 * hh + 4], k1.x, s); s = fmaf(q[8 * hh + 5], k1.y, s); s = fmaf(q[8 * hh + 6], k1.z, s);
            s = fmaf(q[8 * hh + 7], k1.w, s);
            const float p = __expf((s - mx) * scale);
            l += p;
            const float4 v0 = *reinterpret_cast<const float4*>(r2_ptr + 16384 + r * 128 + (((2 * hh) ^ (r & 7)) * 16));
            const float4 v1 =
                *reinterpret_cast<const float4*>(r2_ptr + 16384 + r * 128 + (((2 * hh + 1) ^ (r & 7)) * 16));
            acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v0.z, acc[2]);
            acc[3] = fmaf(p, v0.w, acc[3]); acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]);
            acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
          }
        }
        const float inv = 1.f / l;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[8 * hh + e] = acc[e] * inv;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        split8(o + 8 * c, hi, lo);
        st_shared_v4(R1 + (4 * half + c) * LBO + m * 16, hi);
        st_shared_v4(R1 + 16384 + (4 * half + c) * LBO + m * 16, lo);
      }
      named_bar_sync(1, 128);  // everyone done with this half's K/V before it is overwritten / reused
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);

    // ---- phase 2: X1 = X + O Wo^T (stash), LN2(X1) -> R1
    mbar_wait(mma_done, 1);
    tc_fence_after();
    {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float d[16], x[16];
        tmem_ld16(trow + 16 * c, d);
        tmem_ld16(trow + 192 + 16 * c, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] += d[i]; sum += x[i]; }
        tmem_st16(trow + 192 + 16 * c, x);
      }
      tmem_wait_st();
      const float mean = sum * (1.f / 64.f);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[16];
        tmem_ld16(trow + 192 + 16 * c, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float dd = x[i] - mean; var = fmaf(dd, dd, var); }
      }
      const float rstd = rsqrtf(var * (1.f / 64.f) + 1e-5f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[16];
        tmem_ld16(trow + 192 + 16 * c, x);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          x[i] = (x[i] - mean) * rstd * __ldg(ln + 128 + 16 * c + i) + __ldg(ln + 192 + 16 * c + i);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 hi, lo;
          split8(x + 8 * j, hi, lo);
          st_shared_v4(R1 + (2 * c + j) * LBO + m * 16, hi);
          st_shared_v4(R1 + 16384 + (2 * c + j) * LBO + m * 16, lo);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    // ---- phase 3: hidden = relu(D[0,128)) -> K=128 operand (hi in R1, lo in R2)
    mbar_wait(mma_done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float d[16];
      tmem_ld16(trow + 16 * c, d);
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = fmaxf(d[i], 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 hi, lo;
        split8(d + 8 * j, hi, lo);
        st_shared_v4(R1 + (2 * c + j) * LBO + m * 16, hi);
        st_shared_v4(R2 + (2 * c + j) * LBO + m * 16, lo);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);

    // ---- phase 4: X2 = X1 + D[128,192) -> global
    mbar_wait(mma_done, 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float d[16], x[16];
      tmem_ld16(trow + 128 + 16 * c, d);
      tmem_ld16(trow + 192 + 16 * c, x);
      if (rowok) {
        float4* op = reinterpret_cast<float4*>(out + tok * 64 + 16 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          op[i] = make_float4(x[4 * i] + d[4 * i], x[4 * i + 1] + d[4 * i + 1], x[4 * i + 2] + d[4 * i + 2],
                              x[4 * i + 3] + d[4 * i + 3]);
      }
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256);
}

int configure_ang() {
  CUDA_TRY(cudaFuncSetAttribute(k_ang, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  return 0;
}

int run_ang(Handle* h, int layer, const float* in, float* out, int B, int P, cudaStream_t st) {
  const int N = h->cfg.ang_res * h->cfg.ang_res;
  const long long npix = (long long)B * P * P;
  const int PPT = 128 / N;
  const unsigned grid = (unsigned)((npix + PPT - 1) / PPT);
  const Layer& L = h->layer[layer];
  Scope sc(h, K_ANG, st);
  k_ang<<<grid, kThreads, kSmemAng, st>>>(in, out, L.a_wqk, L.a_wv, L.a_wo, L.a_w1, L.a_w2, L.a_ln, h->pe_ang, N, P * P,
                                         npix, h->passes());
  return sc.finish();
}

}  // namespace lft
