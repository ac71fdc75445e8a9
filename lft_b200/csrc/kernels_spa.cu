// SpaTrans (model/LFT.py:118-191) after the 3x3 token embedding (k_conv3x3<128>, kernels_conv.cu):
//   k_spa_qkv  : Yn = LN(tok + PE_s);  Q = Yn Wq^T, K = Yn Wk^T, V = tok Wv^T          (tcgen05)
//   k_spa_attn : per head (hd=16) softmax over the clamped 5x5 window (<=25 keys) -- the finite entries
//                of gen_mask (LFT.py:147-162) -- never materialising the [hw,hw] mask     (CUDA cores)
//   k_spa_ffn  : Y1 = tok + O Wo^T; Y2 = Y1 + W2 relu(W1 LN2(Y1)); out = Y2 Wlin^T (1x1x1 conv 128->64)
//                (+ the global residual of LFT.py:76 on the last block)                    (tcgen05)
// Q/K/V/O use a planar head-major layout [view][head][y][j][x][4] (channel = head*16 + j*4 + e) so that
// both the row-owner threads of the GEMM kernels and the x-major threads of the window attention
// read/write 16-byte pieces that are contiguous across a warp.
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

// write 16 accumulator columns [c0, c0+16) (= head c0/16) of one token into the planar layout
LFT_DEVINL void planar_store16(float* base, long long v, int c0, int y, int x, int P, const float* d) {
  const int head = c0 >> 4;
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
__global__ void __launch_bounds__(kThreads, 2)
k_spa_qkv(const float* __restrict__ tok, const float* __restrict__ pe, const float* __restrict__ ln,
          const uint8_t* __restrict__ wq, const uint8_t* __restrict__ wk, const uint8_t* __restrict__ wv,
          float* __restrict__ Q, float* __restrict__ K, float* __restrict__ Vv, long long T, int P, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t A = smem_u32(smem) + kCtlBytes;
  const uint32_t ring = A + 65536;
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cta_setup<kSpaNST>(ctl, warp, lane, 128, 256);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_q{wq, 128, 2}, g_k{wk, 128, 2}, g_v{wv, 128, 2};

  if (warp == kWarpProducer) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_q, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_k, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_v, passes);
    }
  } else if (warp == kWarpMma) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      mbar_wait(a_ready, 0);
      tc_fence_after();
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_q, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 0, true);
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_k, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 128, true);
      umma_commit(mma_done);
      mbar_wait(a_ready, 1);
      tc_fence_after();
      ring_consume_mma<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_v, passes, A, A + 32768, kLbo, 8 * kLbo, NoShift{},
                                tmem + 0, true);
      umma_commit(mma_done);
    }
  } else {
    const int m = tid;
    const long long t = (long long)blockIdx.x * 128 + m;
    const bool ok = t < T;
    const long long tt = ok ? t : 0;
    const int PP = P * P;
    const long long v = tt / PP;
    const int p = (int)(tt - v * PP);
    const int y = p / P, x = p - y * P;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const float4* trp = reinterpret_cast<const float4*>(tok + tt * 128);
    const float4* pep = reinterpret_cast<const float4*>(pe + (long long)p * 128);

    // phase 0: z = tok + PE (stash in TMEM [128,256)), LN -> A
    float sum = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = ok ? __ldg(trp + 4 * c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = __ldg(pep + 4 * c + i);
        z[4 * i] = a.x + b.x; z[4 * i + 1] = a.y + b.y; z[4 * i + 2] = a.z + b.z; z[4 * i + 3] = a.w + b.w;
        sum += (z[4 * i] + z[4 * i + 1]) + (z[4 * i + 2] + z[4 * i + 3]);
      }
      tmem_st16(trow + 128 + 16 * c, z);
    }
    tmem_wait_st();
    const float mean = sum * (1.f / 128.f);
    float var = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 16 * c, z);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = z[i] - mean; var = fmaf(d, d, var); }
    }
    const float rstd = rsqrtf(var * (1.f / 128.f) + 1e-5f);
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 16 * c, z);
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = (z[i] - mean) * rstd * __ldg(ln + 16 * c + i) + __ldg(ln + 128 + 16 * c + i);
      a_store16(A, 2 * c, m, z);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);

    // phase 1: Q out, then refill A with raw tok (V operand), then K out while the V MMAs run
    mbar_wait(mma_done, 0);
    tc_fence_after();
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float d[16];
      tmem_ld16(trow + 16 * c, d);
      if (ok) planar_store16(Q, v, 16 * c, y, x, P, d);
    }
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = ok ? __ldg(trp + 4 * c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        z[4 * i] = a.x; z[4 * i + 1] = a.y; z[4 * i + 2] = a.z; z[4 * i + 3] = a.w;
      }
      a_store16(A, 2 * c, m, z);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(a_ready);
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float d[16];
      tmem_ld16(trow + 128 + 16 * c, d);
      if (ok) planar_store16(K, v, 16 * c, y, x, P, d);
    }
    // phase 2: V out
    mbar_wait(mma_done, 1);
    tc_fence_after();
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float d[16];
      tmem_ld16(trow + 16 * c, d);
      if (ok) planar_store16(Vv, v, 16 * c, y, x, P, d);
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256);
}

// ------------------------------------------------------------------------------------------------
// Window attention: one thread = one (token, head). gid -> x fastest, then y, head, view.
__global__ void __launch_bounds__(128)
k_spa_attn(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ Vv,
           float* __restrict__ O, long long nviews, int P) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = nviews * 8 * P * P;
  if (gid >= total) return;
  const int x = (int)(gid % P);
  const int y = (int)((gid / P) % P);
  const int head = (int)((gid / ((long long)P * P)) & 7);
  const long long v = gid / ((long long)P * P * 8);
  float q[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(Q + planar_off(v, head, y, j, x, P)));
    q[4 * j] = f.x; q[4 * j + 1] = f.y; q[4 * j + 2] = f.z; q[4 * j + 3] = f.w;
  }
  float s[25];
  float mx = -INFINITY;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int ky = y + dy, kx = x + dx;
      const int i = (dy + 2) * 5 + dx + 2;
      if (ky >= 0 && ky < P && kx >= 0 && kx < P) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(K + planar_off(v, head, ky, j, kx, P)));
          acc = fmaf(q[4 * j], f.x, acc); acc = fmaf(q[4 * j + 1], f.y, acc);
          acc = fmaf(q[4 * j + 2], f.z, acc); acc = fmaf(q[4 * j + 3], f.w, acc);
        }
        s[i] = acc * 0.25f;  // 1/sqrt(16)
        mx = fmaxf(mx, s[i]);
      } else {
        s[i] = -INFINITY;
      }
    }
  float l = 0.f;
  float o[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) o[e] = 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int ky = y + dy, kx = x + dx;
      const int i = (dy + 2) * 5 + dx + 2;
      if (ky >= 0 && ky < P && kx >= 0 && kx < P) {
        const float p = __expf(s[i] - mx);
        l += p;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(Vv + planar_off(v, head, ky, j, kx, P)));
          o[4 * j] = fmaf(p, f.x, o[4 * j]); o[4 * j + 1] = fmaf(p, f.y, o[4 * j + 1]);
          o[4 * j + 2] = fmaf(p, f.z, o[4 * j + 2]); o[4 * j + 3] = fmaf(p, f.w, o[4 * j + 3]);
        }
      }
    }
  const float inv = 1.f / l;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(O + planar_off(v, head, y, j, x, P)) =
        make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
k_spa_ffn(const float* __restrict__ O, float* __restrict__ tok, const float* __restrict__ ln,
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
  cta_setup<kSpaNST>(ctl, warp, lane, 128, 256);
  const uint32_t tmem = ctl->tmem;
  const GemmPhase g_o{wo, 128, 2}, g_1a{w1a, 128, 2}, g_2a{w2a, 128, 2}, g_1b{w1b, 128, 2}, g_2b{w2b, 128, 2},
      g_l{wlin, 64, 2};

  if (warp == kWarpProducer) {
    if (lane == 0) {
      RingState<kSpaNST> rs;
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_o, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1a, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2a, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_1b, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_2b, passes);
      ring_produce<kSpaNST>(rs, ring, kSpaStage, full0, empty0, g_l, passes);
    }
  } else if (warp == kWarpMma) {
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
      step(g_1a, 0, true);     // D[0,128)    = LN2(Y1) W1[0:128]^T
      step(g_2a, 128, false);  // S[128,256) += relu(.) W2[:,0:128]^T     (S was initialised to Y1)
      step(g_1b, 0, true);     // D[0,128)    = LN2(Y1) W1[128:256]^T
      step(g_2b, 128, false);  // S          += relu(.) W2[:,128:256]^T   -> S = Y2
      step(g_l, 0, true);      // D[0,64)     = Y2 Wlin^T
    }
  } else {
    const int m = tid;
    const long long t = (long long)blockIdx.x * 128 + m;
    const bool ok = t < T;
    const long long tt = ok ? t : 0;
    const int PP = P * P;
    const long long v = tt / PP;
    const int p = (int)(tt - v * PP);
    const int y = p / P, x = p - y * P;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    float* trow_g = tok + tt * 128;
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

    // phase 0: A <- O (planar gather)
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 f = ok ? __ldg(reinterpret_cast<const float4*>(O + planar_off(v, c, y, j, x, P)))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        z[4 * j] = f.x; z[4 * j + 1] = f.y; z[4 * j + 2] = f.z; z[4 * j + 3] = f.w;
      }
      a_store16(A, 2 * c, m, z);
    }
    publish();

    // phase 1: Y1 = tok + D -> global (in place) and TMEM S; LN2 -> A
    await();
    float sum = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float d[16];
      tmem_ld16(trow + 16 * c, d);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = ok ? *reinterpret_cast<const float4*>(trow_g + 16 * c + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[4 * i] += a.x; d[4 * i + 1] += a.y; d[4 * i + 2] += a.z; d[4 * i + 3] += a.w;
        sum += (d[4 * i] + d[4 * i + 1]) + (d[4 * i + 2] + d[4 * i + 3]);
      }
      tmem_st16(trow + 128 + 16 * c, d);
      if (ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(trow_g + 16 * c + 4 * i) = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
      }
    }
    tmem_wait_st();
    const float mean = sum * (1.f / 128.f);
    float var = 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 16 * c, z);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = z[i] - mean; var = fmaf(d, d, var); }
    }
    const float rstd = rsqrtf(var * (1.f / 128.f) + 1e-5f);
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 16 * c, z);
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = (z[i] - mean) * rstd * __ldg(ln + 256 + 16 * c + i) + __ldg(ln + 384 + 16 * c + i);
      a_store16(A, 2 * c, m, z);
    }
    publish();

    // phases 2..5: hidden halves
    for (int half = 0; half < 2; ++half) {
      await();  // FFN1 half done: D[0,128)
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        float d[16];
        tmem_ld16(trow + 16 * c, d);
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = fmaxf(d[i], 0.f);
        a_store16(A, 2 * c, m, d);
      }
      publish();
      await();  // FFN2 half accumulated into S
      if (half == 0) {
        // rebuild LN2(Y1) from the Y1 row written in phase 1 (same thread wrote it)
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
          float z[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = ok ? *reinterpret_cast<const float4*>(trow_g + 16 * c + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            z[4 * i] = a.x; z[4 * i + 1] = a.y; z[4 * i + 2] = a.z; z[4 * i + 3] = a.w;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = (z[i] - mean) * rstd * __ldg(ln + 256 + 16 * c + i) + __ldg(ln + 384 + 16 * c + i);
          a_store16(A, 2 * c, m, z);
        }
        publish();
      }
    }
    // phase 6: A <- Y2 = S
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      float z[16];
      tmem_ld16(trow + 128 + 16 * c, z);
      a_store16(A, 2 * c, m, z);
    }
    publish();
    // phase 7: out = D[0,64) (+ global residual)
    await();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float d[16];
      tmem_ld16(trow + 16 * c, d);
      if (ok) {
        float4* op = reinterpret_cast<float4*>(out + t * 64 + 16 * c);
        if (final_res) {
          const float4* rp = reinterpret_cast<const float4*>(final_res + t * 64 + 16 * c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 r = __ldg(rp + i);
            d[4 * i] += r.x; d[4 * i + 1] += r.y; d[4 * i + 2] += r.z; d[4 * i + 3] += r.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) op[i] = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
      }
    }
    tc_fence_before();
  }
  cta_teardown(ctl, warp, 256);
}

int configure_spa() {
  CUDA_TRY(cudaFuncSetAttribute(k_spa_qkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
  CUDA_TRY(cudaFuncSetAttribute(k_spa_ffn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSpa));
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
  if ((rc = launch_conv3x3(h, 128, in, L.s_wmlp, w.tok, nullptr, (int)V, P, 0, st))) return rc;
  const unsigned grid = (unsigned)((T + 127) / 128);
  {
    Scope sc(h, K_SPA_QKV, st);
    k_spa_qkv<<<grid, kThreads, kSmemSpa, st>>>(w.tok, L.s_pe, L.s_ln, L.s_wq, L.s_wk, L.s_wv, w.q, w.k, w.v, T, P,
                                               h->passes());
    if ((rc = sc.finish())) return rc;
  }
  {
    Scope sc(h, K_SPA_ATTN, st);
    const long long total = T * 8;
    k_spa_attn<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(w.q, w.k, w.v, w.o, V, P);
    if ((rc = sc.finish())) return rc;
  }
  {
    Scope sc(h, K_SPA_FFN, st);
    k_spa_ffn<<<grid, kThreads, kSmemSpa, st>>>(w.o, w.tok, L.s_ln, L.s_wo, L.s_w1a, L.s_w1b, L.s_w2a, L.s_w2b, L.s_wlin,
                                               out, final_res, T, P, h->passes());
    if ((rc = sc.finish())) return rc;
  }
  return 0;
}

}  // namespace lft
