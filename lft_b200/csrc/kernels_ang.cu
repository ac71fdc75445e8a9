// Fused AngTrans (model/LFT.py:194-238): for every pixel, MHSA over its N=A*A angular tokens + FFN.
//   Xn = LN(X + PE_a); Q,K = Xn Wq^T, Xn Wk^T; V = X Wv^T (raw tokens!); per head (hd=8) softmax(QK^T/sqrt 8) V;
//   X1 = X + O Wo^T;  X2 = X1 + W2 relu(W1 LN2(X1)).
// One CTA = 128 accumulator rows = floor(128/N) pixels x N views, 2 threads per row (channel halves = heads
// 4q..4q+3).  The (b c a h w) <-> (a, b h w, c) permutes of the reference (LFT.py:216-223) are address
// arithmetic in the row loads/stores.  Both LayerNorms are folded into the following projection
// (LN(z) W^T = rstd (x W'^T + PE W'^T - mean u) + c), so Q, K and V come from ONE raw operand.
// Projections/FFN run on tcgen05; the N x N (hd 8) attention core is CUDA-core work with K/V shared
// through shared memory (fp32, XOR-swizzled 16-byte chunks).
#include "host.h"
#include "kernels.cuh"

#include <cstring>

namespace lft {

constexpr int kAngNST = 3;
constexpr uint32_t kAngStage = 128 * 128;  // largest slab: N=128 rows
constexpr size_t kSmemAng = kCtlBytes + 65536 + kAngNST * kAngStage;

// q . k over 8 dims with packed FFMA2: q as 4 f32x2, k row as two 16-byte loads
LFT_DEVINL float dot8(const f32x2* q, const ulonglong2& k0, const ulonglong2& k1) {
  f32x2 a = mul2(q[0], k0.x);
  f32x2 b = mul2(q[1], k0.y);
  a = fma2(q[2], k1.x, a);
  b = fma2(q[3], k1.y, b);
  return hsum2(add2(a, b));
}

// Attention of ONE head for NQ queries of the same pixel held by one thread (A=5 path): the 25 keys/values of that
// pixel and head are read ONCE from shared memory for all NQ queries.  Softmax is evaluated online in chunks of 5
// keys (scores of one chunk live in registers; the accumulators are rescaled once per chunk).  q is pre-scaled by
// log2(e)/sqrt(hd).  kb/vb: this pixel's first key in the [head][half][kv row][4 floats] planes (kv row = t*5 + pixel).
// Generalised: NK keys in chunks of CH (NK % CH == 0), key t at kv row t*KS + pixel (KS = pixels per tile).
template <int NQ, int NK, int KS, int CH>
LFT_DEVINL void ang_attn_head(const f32x2 (*q)[4], const ulonglong2* __restrict__ kb, const ulonglong2* __restrict__ vb,
                              float (*o)[8]) {
  static_assert(NK % CH == 0, "chunking");
  float mx[NQ], l[NQ];
  f32x2 acc[NQ][4];
#pragma unroll
  for (int x = 0; x < NQ; ++x) {
    mx[x] = -INFINITY;
    l[x] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[x][e] = 0ull;
  }
#pragma unroll(NK / CH <= 5 ? NK / CH : 1)
  for (int c = 0; c < NK / CH; ++c) {
    float sc[NQ][CH];
    const ulonglong2* kc = kb + KS * CH * c;
    const ulonglong2* vc = vb + KS * CH * c;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const ulonglong2 k0 = kc[KS * j], k1 = kc[KS * j + 128];
#pragma unroll
      for (int x = 0; x < NQ; ++x) sc[x][j] = dot8(q[x], k0, k1);
    }
#pragma unroll
    for (int x = 0; x < NQ; ++x) {
      float cm = sc[x][0];
#pragma unroll
      for (int j = 1; j < CH; ++j) cm = fmaxf(cm, sc[x][j]);
      const float mn = fmaxf(mx[x], cm);
      const float corr = fast_exp2(mx[x] - mn);  // first chunk: exp2(-inf) = 0 on zero accumulators
      mx[x] = mn;
      l[x] *= corr;
      const f32x2 c2 = pack2(corr, corr);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[x][e] = mul2(acc[x][e], c2);
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const ulonglong2 v0 = vc[KS * j], v1 = vc[KS * j + 128];
#pragma unroll
      for (int x = 0; x < NQ; ++x) {
        const float pw = fast_exp2(sc[x][j] - mx[x]);
        l[x] += pw;
        const f32x2 pp = pack2(pw, pw);
        acc[x][0] = fma2(pp, v0.x, acc[x][0]); acc[x][1] = fma2(pp, v0.y, acc[x][1]);
        acc[x][2] = fma2(pp, v1.x, acc[x][2]); acc[x][3] = fma2(pp, v1.y, acc[x][3]);
      }
    }
  }
#pragma unroll
  for (int x = 0; x < NQ; ++x) {
    const float inv = 1.f / l[x];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a, b;
      unpack2(acc[x][e], a, b);
      o[x][2 * e] = a * inv;
      o[x][2 * e + 1] = b * inv;
    }
  }
}

// One head pair (relative heads h2 and 2+h2 of this thread's channel half) of the paired-view attention for a lane pair:
// the lower lane (of l, l+16) computes head h2, the upper lane head 2+h2, each for its own query and its partner's; queries and
// results travel by warp shuffles.  Q is read from the fp32 stash in TMEM columns [tq, tq+32); results are written as the bf16
// hi/lo TS-form A operand of the output projection.
template <int NK, int KS, int CH>
LFT_DEVINL void ang_pair_heads(uint32_t tq, uint32_t to_hi, uint32_t to_lo, int h2, bool up, int partner,
                               const uint8_t* ks_q, const uint8_t* vs_q, int pl, bool fp32_mode) {
  float qa[8], qb[8];
  tmem_ld8(tq + 8 * h2, qa);
  tmem_ld8(tq + 16 + 8 * h2, qb);
  f32x2 qq[2][4];
  {
    float mine[8], part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = up ? qb[i] : qa[i];
      part[i] = __shfl_sync(0xffffffffu, up ? qa[i] : qb[i], partner);  // the partner's query, MY head
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      qq[0][e] = pack2(mine[2 * e], mine[2 * e + 1]);
      qq[1][e] = pack2(part[2 * e], part[2 * e + 1]);
    }
  }
  const int head = (up ? 2 : 0) + h2;
  const ulonglong2* kb = reinterpret_cast<const ulonglong2*>(ks_q + head * 4096) + pl;
  const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(vs_q + head * 4096) + pl;
  float o[2][8];
  ang_attn_head<2, NK, KS, CH>(qq, kb, vb, o);
  float oa[8], ob[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float recv = __shfl_sync(0xffffffffu, o[1][i], partner);  // my query, the partner's head
    oa[i] = up ? recv : o[0][i];
    ob[i] = up ? o[0][i] : recv;
  }
  uint4 hi, lo;
  split8(oa, hi, lo, fp32_mode);
  tmem_st4u(to_hi + 4 * h2, hi);
  if (fp32_mode) tmem_st4u(to_lo + 4 * h2, lo);
  split8(ob, hi, lo, fp32_mode);
  tmem_st4u(to_hi + 8 + 4 * h2, hi);
  if (fp32_mode) tmem_st4u(to_lo + 8 + 4 * h2, lo);
}

// A = 5: the five rows of view 24 (one per pixel) have no partner view.  After the pair pass the last row warp of each
// channel half runs this pass: unit (pixel s, relative head hh) goes to lane 8*hh + s (a quarter-warp reads one head: no bank
// conflicts), the query comes from the row's own lane by shuffle (lane 12+s, or 28 for s = 4), the result goes back the
// same way and overwrites the place-holder the pair pass wrote.  One head x one query per lane: ~1/4 of the pair pass.
LFT_DEVINL void ang_singles25(uint32_t tq, uint32_t to_hi, uint32_t to_lo, int lane, const uint8_t* ks_q, const uint8_t* vs_q,
                              bool fp32_mode) {
  const int hh = lane >> 3, s = lane & 7;                 // unit of this lane (valid for s < 5)
  const int sv = s < 5 ? s : 0;
  const int qsrc = sv < 4 ? 12 + sv : 28;                 // lane that owns the row of pixel sv
  const int my_s = lane == 28 ? 4 : lane - 12;            // pixel of the row this lane owns (meaningful on lanes 12..15, 28)
  const bool owner = (lane >= 12 && lane < 16) || lane == 28;
  f32x2 qq[1][4];
  {
    float mine[8];
#pragma unroll
    for (int h = 0; h < 4; ++h) {                         // every lane reads ITS row's Q of head h; the owners' values are picked up
      float qh[8];
      tmem_ld8(tq + 8 * h, qh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = __shfl_sync(0xffffffffu, qh[i], qsrc);
        if (h == hh) mine[i] = v;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) qq[0][e] = pack2(mine[2 * e], mine[2 * e + 1]);
  }
  const ulonglong2* kb = reinterpret_cast<const ulonglong2*>(ks_q + hh * 4096) + sv;
  const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(vs_q + hh * 4096) + sv;
  float o[1][8];
  ang_attn_head<1, 25, 5, 5>(qq, kb, vb, o);
  tmem_wait_st();                                         // the pair pass's operand stores, re-read below
#pragma unroll
  for (int h = 0; h < 4; ++h) {                           // owners collect head h from lane 8*h + their pixel
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __shfl_sync(0xffffffffu, o[0][i], 8 * h + (owner ? my_s : 0));
    // tcgen05.ld/st are warp-collective (.sync.aligned): every lane executes them; the other lanes write their own
    // (pair-pass) operand words back unchanged
    uint4 hi, lo;
    split8(r, hi, lo, fp32_mode);
    const uint4 oh = tmem_ld4u(to_hi + 4 * h);
    tmem_st4u(to_hi + 4 * h, owner ? hi : oh);
    if (fp32_mode) {
      const uint4 ol = tmem_ld4u(to_lo + 4 * h);
      tmem_st4u(to_lo + 4 * h, owner ? lo : ol);
    }
  }
}

// One (pixel, head) item of the A = 5 attention for one query: all 25 scores first (independent dot products), one maximum,
// then the weighted sum - no online rescaling (25 scores fit in registers).  kb / vb: the pixel's first key in the planes,
// key t at kb[5 t] (first 4 dims) and kb[5 t + 128] (last 4 dims).
LFT_DEVINL void ang_attn_item25(const f32x2* q, const ulonglong2* __restrict__ kb, const ulonglong2* __restrict__ vb, float* o) {
  float sc[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) sc[t] = dot8(q, kb[5 * t], kb[5 * t + 128]);
  float mx = sc[0];
#pragma unroll
  for (int t = 1; t < 25; ++t) mx = fmaxf(mx, sc[t]);
  float l = 0.f;
  f32x2 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int t = 0; t < 25; ++t) {
    const float pw = fast_exp2(sc[t] - mx);
    l += pw;
    const f32x2 pp = pack2(pw, pw);
    const ulonglong2 v0 = vb[5 * t], v1 = vb[5 * t + 128];
    acc[0] = fma2(pp, v0.x, acc[0]); acc[1] = fma2(pp, v0.y, acc[1]);
    acc[2] = fma2(pp, v1.x, acc[2]); acc[3] = fma2(pp, v1.y, acc[3]);
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float a, b;
    unpack2(acc[e], a, b);
    o[2 * e] = a * inv;
    o[2 * e + 1] = b * inv;
  }
}

// A = 5 (25 views per pixel, 5 pixels per tile): the attention of a tile is re-mapped from the row owners to
// (pixel, head) work items: a WARP works on ONE pixel and one head with lane = query view, so all lanes read the same key /
// value at the same time (uniform-address LDS.128) and no query pairing, partner shuffles or "single view" pass are needed.
// Measured against the paired row-owner formulation of round 1 (profiles/r02_counters.md): the same time per launch within
// 3 % (1.11 vs 1.08 ms under ncu), 156 M instead of 126 M shared-memory wavefronts (a uniform 16-byte load still costs two
// wavefronts; 25 of 32 lanes are used and 20 items leave 8 warps a third, partly filled round), L1/shared pipe 69 %, issue
// 44 %: the per-pixel 25 x 25 x 8 attention stays bound by the shared-memory pipe whichever way the work is mapped - each
// key / value row must reach 25 queries through it.  Kept because it is the simpler formulation (and the basis for a
// tensor-core S = Q K^T, which needs the same data movement).
// Rows stay view-major (global accesses of the owners stay coalesced); Q, K and V leave the accumulators through shared
// memory, one head half (4 heads) at a time:
//   R2: K [rel head 4][half 2][kv row 128][16 B] 16 KB | V 16 KB         (kv row = view * 5 + pixel)
//   R1: Q / O of head half 0 (16 KB, same layout, O overwrites Q in place) | of head half 1 (16 KB)
// per half g: export (thread 0 of a row: K, thread 1: V and Q, both with the LayerNorm-fold correction) | barrier | 20 items
// (pixel, rel head) over the 8 warps | barrier.  The O operand of the output projection (TS form, TMEM columns [64,128))
// overlaps K accumulator columns of the other half, so half 0's results stay in shared memory until half 1 has been
// exported.
// Generalised to the other odd angular resolutions that fill a 128-row tile with whole pixels: NV = 49 (A = 7, 2 pixels per
// tile) and NV = 81 (A = 9, 1 pixel per tile).  More than 32 views -> a (pixel, head) item is split into chunks of 32 queries
// (one per lane); the soft-max is evaluated online in chunks of CH keys (81 scores do not fit in registers).  For A = 9 the
// paired row-owner attention was bound by shared-memory wavefronts (all 81 rows of a tile read the same 81 keys, but every
// quarter-warp fetched them separately: ~41 K wavefronts per tile); the uniform loads of this mapping need ~16 K.
template <int N, int PPT, int CH>
LFT_DEVINL void ang_attention_items(uint32_t trow, int warp, int lane, int q, int kvrow, float rstd, float mr,
                                    const float4* __restrict__ pq4, const float4* tab4, uint8_t* planes, bool fp32_mode) {
  constexpr int QCH = (N + 31) / 32;           // query chunks per (pixel, head)
  constexpr int NITEMS = 4 * PPT * QCH;        // items per head half
  const int srow = kvrow >= 0 ? kvrow : 0;     // idle rows (A = 7, 9) stage nothing and convert row 0's values (their output is discarded)
  uint8_t* qo_ptr = planes;           // R1
  uint8_t* ks_ptr = planes + 32768;   // R2
  uint8_t* vs_ptr = ks_ptr + 16384;
  const float scale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e)/sqrt(8), folded into Q (softmax via exp2)
  // results of head half g: shared memory -> bf16 hi/lo TS-form operand; the two threads of a row take two heads each
  auto convert = [&](int g) {
    const uint8_t* src = qo_ptr + g * 16384 + srow * 16;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int rh = 2 * q + hh;
      float o[8];
      *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(src + (rh * 2) * 2048);
      *reinterpret_cast<float4*>(o + 4) = *reinterpret_cast<const float4*>(src + (rh * 2 + 1) * 2048);
      uint4 hi, lo;
      split8(o, hi, lo, fp32_mode);
      tmem_st4u(trow + 64 + 16 * g + 4 * rh, hi);
      if (fp32_mode) tmem_st4u(trow + 96 + 16 * g + 4 * rh, lo);
    }
  };
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    float kv[16];
    LFT_TL(12 + 4 * g);
    if (q == 0) {  // K of heads 4g..4g+3 (accumulator columns 64 + 32g ..), corrected
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = 64 + 32 * g + 16 * c;
        tmem_ld16(trow + col, kv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pv = __ldg(pq4 + (col / 4 + j) * N);
          const float4 uv = tab4[col / 4 + j], cv = tab4[32 + col / 4 + j];
          const float4 r = make_float4(fmaf(rstd, kv[4 * j] + pv.x, fmaf(-mr, uv.x, cv.x)),
                                       fmaf(rstd, kv[4 * j + 1] + pv.y, fmaf(-mr, uv.y, cv.y)),
                                       fmaf(rstd, kv[4 * j + 2] + pv.z, fmaf(-mr, uv.z, cv.z)),
                                       fmaf(rstd, kv[4 * j + 3] + pv.w, fmaf(-mr, uv.w, cv.w)));
          if (kvrow >= 0) *reinterpret_cast<float4*>(ks_ptr + ((2 * c + (j >> 1)) * 2 + (j & 1)) * 2048 + kvrow * 16) = r;
        }
      }
    } else {       // V (raw) and Q (corrected, pre-scaled) of the same heads
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld16(trow + 128 + 32 * g + 16 * c, kv);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (kvrow >= 0)
            *reinterpret_cast<float4*>(vs_ptr + ((2 * c + (j >> 1)) * 2 + (j & 1)) * 2048 + kvrow * 16) =
                make_float4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = 32 * g + 16 * c;
        tmem_ld16(trow + col, kv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pv = __ldg(pq4 + (col / 4 + j) * N);
          const float4 uv = tab4[col / 4 + j], cv = tab4[32 + col / 4 + j];
          const float4 r = make_float4(scale * fmaf(rstd, kv[4 * j] + pv.x, fmaf(-mr, uv.x, cv.x)),
                                       scale * fmaf(rstd, kv[4 * j + 1] + pv.y, fmaf(-mr, uv.y, cv.y)),
                                       scale * fmaf(rstd, kv[4 * j + 2] + pv.z, fmaf(-mr, uv.z, cv.z)),
                                       scale * fmaf(rstd, kv[4 * j + 3] + pv.w, fmaf(-mr, uv.w, cv.w)));
          if (kvrow >= 0) *reinterpret_cast<float4*>(qo_ptr + g * 16384 + ((2 * c + (j >> 1)) * 2 + (j & 1)) * 2048 + kvrow * 16) = r;
        }
      }
    }
    LFT_TL(13 + 4 * g);
    tc_fence_before();
    rows_bar_sync256();  // planes of half g complete; for g = 1: every Q / K / V accumulator column has been consumed
    tc_fence_after();
    LFT_TL(14 + 4 * g);
    if (g == 1) convert(0);
    const uint8_t* qo_g = qo_ptr + g * 16384;
#pragma unroll 1
#ifdef LFT_X_ANG_NOATTN   // timing experiment (wrong results): the attention items are skipped
    for (int it = NITEMS; it < NITEMS; it += 8) {
#else
    for (int it = warp; it < NITEMS; it += 8) {  // items (rel head, pixel, query chunk); lane = query view of the chunk
#endif
      const int rh = it / (PPT * QCH);
      const int rem = it - rh * (PPT * QCH);
      const int p = rem / QCH, qc = rem - p * QCH;
      const int aq = 32 * qc + lane;
      const int a = aq < N ? aq : N - 1;         // lanes beyond the last view shadow it (no store)
      uint8_t* qrow = const_cast<uint8_t*>(qo_g) + (rh * 2) * 2048 + (a * PPT + p) * 16;
      f32x2 qq[1][4];
      {
        const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(qrow);
        const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(qrow + 2048);
        qq[0][0] = q0.x; qq[0][1] = q0.y; qq[0][2] = q1.x; qq[0][3] = q1.y;
      }
      const ulonglong2* kb = reinterpret_cast<const ulonglong2*>(ks_ptr + rh * 4096) + p;
      const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(vs_ptr + rh * 4096) + p;
      float o[1][8];
      if constexpr (N == 25) {
#ifdef LFT_ANG_ONLINE
        ang_attn_head<1, 25, 5, 5>(qq, kb, vb, o);
#else
        ang_attn_item25(qq[0], kb, vb, o[0]);
#endif
      } else {
        ang_attn_head<1, N, PPT, CH>(qq, kb, vb, o);
      }
      if (aq < N) {  // O overwrites Q in place (only this lane ever read it)
        *reinterpret_cast<float4*>(qrow) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        *reinterpret_cast<float4*>(qrow + 2048) = make_float4(o[0][4], o[0][5], o[0][6], o[0][7]);
      }
    }
    LFT_TL(15 + 4 * g);
    rows_bar_sync256();  // results of half g complete; K / V planes free for the next half
  }
  LFT_TL(20);
  convert(1);
}

// A = 5 on tensor cores (round 2, the default for NV = 25; -DLFT_ANG_ITEMS_V1 selects the FFMA2 items above): the same
// (pixel, head) work items, but S = Q K^T and O = P V are warp-level mma.sync.m16n8k16 (bf16 hi / lo, fp32 accumulate).
//   queries 25 -> two m16 tiles (views 0..15, 16..31: the rows beyond view 24 shadow it and are not written back)
//   keys    25 -> four n8 tiles (keys 25..31 are masked: -inf as the accumulator's initial value)
//   S       : the head dimension is 8, so hi and lo ride side by side along k = 16: A = [Q_hi | Q_lo],
//             B = [K_hi | K_hi] gives Q_hi K_hi + Q_lo K_hi in ONE instruction, B = [K_lo | 0] adds Q_hi K_lo  (16 HMMA)
//   O = P V : k = keys (two k16 steps), n = the 8 dims; P_hi V_hi + P_lo V_hi + P_hi V_lo                      (12 HMMA)
// The planes keep their geometry ([rel head 4][2][kv row 128][16 B], kv row = view * 5 + pixel) but hold bf16: slot 0 = hi of
// the head's 8 dims, slot 1 = lo - exactly the 8 x 8 b16 matrices ldmatrix wants (rows of one pixel are 80 bytes apart: the
// eight 16-byte rows of a matrix fall into eight different bank groups).  O (fp32) overwrites the item's own Q rows: dims
// 0..3 in slot 0, dims 4..7 in slot 1, which is what convert() reads.
LFT_DEVINL void ang_ldsm4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
LFT_DEVINL void ang_ldsm4t(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
LFT_DEVINL void ang_hmma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One work item: NMT m16 tiles (views 16 mt0 .. of (rel head, pixel)) against the pixel's 25 keys.
template <bool FP32, int NMT>
LFT_DEVINL void ang_mma_item(uint32_t qo_u, uint32_t ks_u, uint32_t vs_u, uint8_t* qo_half, int rh, int p, int mt0,
                             const uint32_t* aoff, uint32_t koff, float b3, int g, int c) {
  constexpr int N = 25, PPT = 5;
  const float ninf = -INFINITY;
  const uint32_t ioff = (uint32_t)rh * 4096u + (uint32_t)p * 16u;
  uint32_t a[NMT][4], kh[4], kl[4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) ang_ldsm4(qo_u + ioff + aoff[mt0 + mt], a[mt]);
  ang_ldsm4(ks_u + ioff + koff, kh);
  if (FP32) ang_ldsm4(ks_u + ioff + 2048u + koff, kl);
  float s[NMT][4][4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[mt][n][0] = s[mt][n][2] = n == 3 ? b3 : 0.f;
      s[mt][n][1] = s[mt][n][3] = n == 3 ? ninf : 0.f;
    }
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
      ang_hmma(s[mt][n], a[mt], kh[n], kh[n]);
      if (FP32) ang_hmma(s[mt][n], a[mt], kl[n], 0u);
    }
  // soft-max over the 25 keys of each row this thread holds a slice of (rows g, g + 8 of every m tile): all row maxima first,
  // then all exponentials, then all sums - the shuffles of the 2 NMT rows overlap
  float mx[NMT][2], l[NMT][2];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float m = fmaxf(fmaxf(s[mt][0][2 * hf], s[mt][0][2 * hf + 1]), fmaxf(s[mt][1][2 * hf], s[mt][1][2 * hf + 1]));
      mx[mt][hf] = fmaxf(m, fmaxf(fmaxf(s[mt][2][2 * hf], s[mt][2][2 * hf + 1]), s[mt][3][2 * hf]));
    }
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) mx[mt][hf] = fmaxf(mx[mt][hf], __shfl_xor_sync(0xffffffffu, mx[mt][hf], 1));
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) mx[mt][hf] = fmaxf(mx[mt][hf], __shfl_xor_sync(0xffffffffu, mx[mt][hf], 2));
  uint32_t ph[NMT][2][4], pl_[NMT][2][4];   // [m tile][k16 step] A operands
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float ls = 0.f;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float p0 = fast_exp2(s[mt][n][2 * hf] - mx[mt][hf]);
        const float p1 = n == 3 ? 0.f : fast_exp2(s[mt][n][2 * hf + 1] - mx[mt][hf]);
        ls += p0 + p1;
        uint32_t hi, lo;
        if (FP32) {
          const uint32_t u0 = __float_as_uint(p0), u1 = __float_as_uint(p1);
          hi = __byte_perm(u0, u1, 0x7632);
          lo = pack_bf16(p0 - __uint_as_float(u0 & 0xffff0000u), p1 - __uint_as_float(u1 & 0xffff0000u));
        } else {
          hi = pack_bf16(p0, p1);
          lo = 0u;
        }
        ph[mt][n >> 1][2 * (n & 1) + hf] = hi;
        pl_[mt][n >> 1][2 * (n & 1) + hf] = lo;
      }
      l[mt][hf] = ls;
    }
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) l[mt][hf] += __shfl_xor_sync(0xffffffffu, l[mt][hf], 1);
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) l[mt][hf] += __shfl_xor_sync(0xffffffffu, l[mt][hf], 2);
  // O = P V
  uint32_t vh[4], vl[4];
  ang_ldsm4t(vs_u + ioff + koff, vh);
  if (FP32) ang_ldsm4t(vs_u + ioff + 2048u + koff, vl);
  float o[NMT][4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
      ang_hmma(o[mt], ph[mt][t], vh[2 * t], vh[2 * t + 1]);
      if (FP32) {
        ang_hmma(o[mt], pl_[mt][t], vh[2 * t], vh[2 * t + 1]);
        ang_hmma(o[mt], ph[mt][t], vl[2 * t], vl[2 * t + 1]);
      }
    }
  // write O (fp32) over the item's own Q rows: dims 2c, 2c + 1 -> slot c / 2, bytes 8 (c & 1)
  uint8_t* obase = qo_half + rh * 4096 + (c >> 1) * 2048 + p * 16 + (c & 1) * 8;
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int view = 16 * (mt0 + mt) + 8 * hf + g;
      const float inv = fast_rcp(l[mt][hf]);   // l in [1, 25]
      if (view < N)
        *reinterpret_cast<float2*>(obase + view * (PPT * 16)) = make_float2(o[mt][2 * hf] * inv, o[mt][2 * hf + 1] * inv);
    }
}

template <bool FP32>
LFT_DEVINL void ang_attention_mma25(uint32_t trow, int warp, int lane, int q, int kvrow, float rstd, float mr,
                                    const float4* __restrict__ pq4, const float4* tab4, uint8_t* planes) {
  constexpr int N = 25, PPT = 5;
  constexpr bool fp32_mode = FP32;
  uint8_t* qo_ptr = planes;           // R1: Q / O of head half 0 (16 KB) | of head half 1 (16 KB)
  uint8_t* ks_ptr = planes + 32768;   // R2: K 16 KB | V 16 KB
  uint8_t* vs_ptr = ks_ptr + 16384;
  const float scale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e)/sqrt(8), folded into Q (softmax via exp2)
  const int g = lane >> 2, c = lane & 3;      // accumulator fragment: rows g, g + 8; columns 2c, 2c + 1
  const int mat = lane >> 3, mrow = lane & 7; // ldmatrix: this lane addresses row mrow of matrix mat
  // 8 values of one head -> the hi and the lo piece of this row
  auto put = [&](uint8_t* plane, int rh, const float* x) {
    uint4 hi, lo;
    split8(x, hi, lo, fp32_mode);
    *reinterpret_cast<uint4*>(plane + (rh * 2) * 2048 + kvrow * 16) = hi;
    *reinterpret_cast<uint4*>(plane + (rh * 2 + 1) * 2048 + kvrow * 16) = lo;   // (bf16 mode: zeros - the Q_lo half of A is always read)
  };
  // corrected Q (pre-scaled) / K of accumulator columns col .. col + 15 (two heads) -> planes
  auto put_qk = [&](uint8_t* plane, int col, int rh0, float sc) {
    float kv[16];
    float4 pv[4];   // PE_a W'^T of this view: issued before the accumulator load (whose wait the compiler cannot move loads across)
#pragma unroll
    for (int j = 0; j < 4; ++j) pv[j] = __ldg(pq4 + (col / 4 + j) * N);
    tmem_ld16(trow + col, kv);
    const f32x2 nmr2 = pack2(-mr, -mr), rstd2 = pack2(rstd, rstd), sc2 = pack2(sc, sc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // sc * (rstd * (acc + PE W') - mr * u + c), two columns per packed instruction (same roundings)
      const float4 uv = tab4[col / 4 + j], cv = tab4[32 + col / 4 + j];
      const f32x2 a = mul2(sc2, fma2(rstd2, add2(pack2(kv[4 * j], kv[4 * j + 1]), pack2(pv[j].x, pv[j].y)),
                                     fma2(nmr2, pack2(uv.x, uv.y), pack2(cv.x, cv.y))));
      const f32x2 b = mul2(sc2, fma2(rstd2, add2(pack2(kv[4 * j + 2], kv[4 * j + 3]), pack2(pv[j].z, pv[j].w)),
                                     fma2(nmr2, pack2(uv.z, uv.w), pack2(cv.z, cv.w))));
      unpack2(a, kv[4 * j], kv[4 * j + 1]);
      unpack2(b, kv[4 * j + 2], kv[4 * j + 3]);
    }
    put(plane, rh0, kv);
    put(plane, rh0 + 1, kv + 8);
  };
  // results of head half hg: shared memory (fp32) -> bf16 hi/lo TS-form operand; the two threads of a row take two heads each
  auto convert = [&](int hg) {
    const uint8_t* src = qo_ptr + hg * 16384 + kvrow * 16;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int rh = 2 * q + hh;
      float o[8];
      *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(src + (rh * 2) * 2048);
      *reinterpret_cast<float4*>(o + 4) = *reinterpret_cast<const float4*>(src + (rh * 2 + 1) * 2048);
      uint4 hi, lo;
      split8(o, hi, lo, fp32_mode);
      tmem_st4u(trow + 64 + 16 * hg + 4 * rh, hi);
      if (fp32_mode) tmem_st4u(trow + 96 + 16 * hg + 4 * rh, lo);
    }
  };
  // per-lane ldmatrix row offsets inside a (rel head, pixel) item: view = 8 * matrix-in-group + mrow, rows beyond view 24 shadow it
  uint32_t aoff[2], koff;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)   // A = [Q_hi rows 0-7, Q_hi rows 8-15, Q_lo rows 0-7, Q_lo rows 8-15] of m16 tile mt
    aoff[mt] = (uint32_t)(mat >> 1) * 2048u + (uint32_t)min(16 * mt + 8 * (mat & 1) + mrow, N - 1) * (PPT * 16);
  koff = (uint32_t)min(8 * mat + mrow, N - 1) * (PPT * 16);   // K / V: matrix = keys 8 mat .. 8 mat + 7 (slot 0: hi; + 2048: lo)
  const float b3 = c == 0 ? 0.f : -INFINITY;   // n8 tile 3 = keys 24 + 2c + e: only key 24 exists
#pragma unroll 1
  for (int hg = 0; hg < 2; ++hg) {
    LFT_TL(12 + 4 * hg);
    // exports of heads 4hg .. 4hg + 3, three 16-column chunks per thread: thread 0 of a row K (corrected) and the first half of
    // Q, thread 1 V (raw) and the second half of Q (corrected, pre-scaled)
    if (q == 0) {
      put_qk(ks_ptr, 64 + 32 * hg, 0, 1.f);
      put_qk(ks_ptr, 64 + 32 * hg + 16, 2, 1.f);
      put_qk(qo_ptr + hg * 16384, 32 * hg, 0, scale);
    } else {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float kv[16];
        tmem_ld16(trow + 128 + 32 * hg + 16 * cc, kv);
        put(vs_ptr, 2 * cc, kv);
        put(vs_ptr, 2 * cc + 1, kv + 8);
      }
      put_qk(qo_ptr + hg * 16384, 32 * hg + 16, 2, scale);
    }
    LFT_TL(13 + 4 * hg);
    tc_fence_before();
    rows_bar_sync256();  // planes of half hg complete; for hg = 1: every Q / K / V accumulator column has been consumed
    tc_fence_after();
    LFT_TL(14 + 4 * hg);
    if (hg == 1) convert(0);
    uint8_t* qo_half = qo_ptr + hg * 16384;
    const uint32_t qo_u = smem_u32(qo_half), ks_u = smem_u32(ks_ptr), vs_u = smem_u32(vs_ptr);
    // 20 (rel head, pixel) items: two full rounds of whole items over the 8 warps, the last four items as eight half items
    // (one m16 tile each) - 2.5 item times per warp instead of 3
#ifndef LFT_X_ANG_NOITEMS   // (defined: timing experiment, wrong results)
#pragma unroll 1
    for (int it = warp; it < 16; it += 8)
      ang_mma_item<FP32, 2>(qo_u, ks_u, vs_u, qo_half, it / PPT, it % PPT, 0, aoff, koff, b3, g, c);
    {
      const int it = 16 + (warp >> 1);
      ang_mma_item<FP32, 1>(qo_u, ks_u, vs_u, qo_half, it / PPT, it % PPT, warp & 1, aoff, koff, b3, g, c);
    }
#endif
    LFT_TL(15 + 4 * hg);
    rows_bar_sync256();  // results of half hg complete; K / V planes free for the next half
  }
  LFT_TL(20);
  convert(1);
}

// The same formulation for the larger odd angular resolutions that use the item attention (NV = 49: A = 7, two pixels per tile;
// NV = 81: A = 9, one pixel per tile): an item is ONE m16 tile of queries of a (rel head, pixel) against all NV keys -
// NT = ceil(NV / 8) n8 tiles (padded to an even count for the k16 steps of P V; padding keys are masked), all scores of the two
// rows a thread holds stay in registers (<= 48), one soft-max pass.
template <bool FP32, int NV, int PPT>
LFT_DEVINL void ang_attention_mma_g(uint32_t trow, int warp, int lane, int q, int kvrow, float rstd, float mr,
                                    const float4* __restrict__ pq4, const float4* tab4, uint8_t* planes) {
  constexpr int N = NV;
  constexpr int NT = (NV + 7) / 8, NTP = (NT + 1) & ~1, KS = NTP / 2, MT = (NV + 15) / 16;
  constexpr int NITEMS = 4 * PPT * MT;   // per head half
  uint8_t* qo_ptr = planes;           // R1: Q / O of head half 0 (16 KB) | of head half 1 (16 KB)
  uint8_t* ks_ptr = planes + 32768;   // R2: K 16 KB | V 16 KB
  uint8_t* vs_ptr = ks_ptr + 16384;
  const float scale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e)/sqrt(8), folded into Q (softmax via exp2)
  const int g = lane >> 2, c = lane & 3;
  const int mat = lane >> 3, mrow = lane & 7;
  const int srow = kvrow >= 0 ? kvrow : 0;     // idle rows stage nothing and convert row 0's values (their output is discarded)
  const float ninf = -INFINITY;
  auto put = [&](uint8_t* plane, int rh, const float* x) {
    uint4 hi, lo;
    split8(x, hi, lo, FP32);
    if (kvrow >= 0) {
      *reinterpret_cast<uint4*>(plane + (rh * 2) * 2048 + kvrow * 16) = hi;
      *reinterpret_cast<uint4*>(plane + (rh * 2 + 1) * 2048 + kvrow * 16) = lo;
    }
  };
  auto put_qk = [&](uint8_t* plane, int col, int rh0, float sc) {
    float kv[16];
    float4 pv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pv[j] = __ldg(pq4 + (col / 4 + j) * N);
    tmem_ld16(trow + col, kv);
    const f32x2 nmr2 = pack2(-mr, -mr), rstd2 = pack2(rstd, rstd), sc2 = pack2(sc, sc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // sc * (rstd * (acc + PE W') - mr * u + c), two columns per packed instruction (same roundings)
      const float4 uv = tab4[col / 4 + j], cv = tab4[32 + col / 4 + j];
      const f32x2 a = mul2(sc2, fma2(rstd2, add2(pack2(kv[4 * j], kv[4 * j + 1]), pack2(pv[j].x, pv[j].y)),
                                     fma2(nmr2, pack2(uv.x, uv.y), pack2(cv.x, cv.y))));
      const f32x2 b = mul2(sc2, fma2(rstd2, add2(pack2(kv[4 * j + 2], kv[4 * j + 3]), pack2(pv[j].z, pv[j].w)),
                                     fma2(nmr2, pack2(uv.z, uv.w), pack2(cv.z, cv.w))));
      unpack2(a, kv[4 * j], kv[4 * j + 1]);
      unpack2(b, kv[4 * j + 2], kv[4 * j + 3]);
    }
    put(plane, rh0, kv);
    put(plane, rh0 + 1, kv + 8);
  };
  auto convert = [&](int hg) {
    const uint8_t* src = qo_ptr + hg * 16384 + srow * 16;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int rh = 2 * q + hh;
      float o[8];
      *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(src + (rh * 2) * 2048);
      *reinterpret_cast<float4*>(o + 4) = *reinterpret_cast<const float4*>(src + (rh * 2 + 1) * 2048);
      uint4 hi, lo;
      split8(o, hi, lo, FP32);
      tmem_st4u(trow + 64 + 16 * hg + 4 * rh, hi);
      if (FP32) tmem_st4u(trow + 96 + 16 * hg + 4 * rh, lo);
    }
  };
#pragma unroll 1
  for (int hg = 0; hg < 2; ++hg) {
    if (q == 0) {
      put_qk(ks_ptr, 64 + 32 * hg, 0, 1.f);
      put_qk(ks_ptr, 64 + 32 * hg + 16, 2, 1.f);
      put_qk(qo_ptr + hg * 16384, 32 * hg, 0, scale);
    } else {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float kv[16];
        tmem_ld16(trow + 128 + 32 * hg + 16 * cc, kv);
        put(vs_ptr, 2 * cc, kv);
        put(vs_ptr, 2 * cc + 1, kv + 8);
      }
      put_qk(qo_ptr + hg * 16384, 32 * hg + 16, 2, scale);
    }
    tc_fence_before();
    rows_bar_sync256();
    tc_fence_after();
    if (hg == 1) convert(0);
    uint8_t* qo_half = qo_ptr + hg * 16384;
    const uint32_t qo_u = smem_u32(qo_half), ks_u = smem_u32(ks_ptr), vs_u = smem_u32(vs_ptr);
#pragma unroll 1
    for (int it = warp; it < NITEMS; it += 8) {   // item = (rel head, pixel, m16 tile)
      const int rh = it / (PPT * MT), rem = it - rh * (PPT * MT);
      const int p = rem / MT, mt = rem - p * MT;
      const uint32_t ioff = (uint32_t)rh * 4096u + (uint32_t)p * 16u;
      uint32_t a[4];
      ang_ldsm4(qo_u + ioff + (uint32_t)(mat >> 1) * 2048u + (uint32_t)min(16 * mt + 8 * (mat & 1) + mrow, N - 1) * (PPT * 16), a);
      float s[NTP][4];
#pragma unroll
      for (int n = 0; n < NTP; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool dead = 8 * n + 6 + e >= N;      // can the column 8n + 2c + e fall beyond the last key for some c?
          const float b = !dead ? 0.f : (8 * n + 2 * c + e < N ? 0.f : ninf);
          s[n][e] = s[n][2 + e] = b;
        }
#pragma unroll
      for (int grp = 0; grp < (NT + 3) / 4; ++grp) {   // K fragments of n8 tiles 4 grp .. 4 grp + 3
        uint32_t kh[4], kl[4];
        const uint32_t ko = (uint32_t)min(32 * grp + 8 * mat + mrow, N - 1) * (PPT * 16);
        ang_ldsm4(ks_u + ioff + ko, kh);
        if (FP32) ang_ldsm4(ks_u + ioff + 2048u + ko, kl);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = 4 * grp + j;
          if (n < NT) {
            ang_hmma(s[n], a, kh[j], kh[j]);
            if (FP32) ang_hmma(s[n], a, kl[j], 0u);
          }
        }
      }
      uint32_t ph[KS][4], pl_[KS][4];
      float linv[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float m0 = s[0][2 * hf], m1 = s[0][2 * hf + 1];
#pragma unroll
        for (int n = 1; n < NT; ++n) {
          m0 = fmaxf(m0, s[n][2 * hf]);
          m1 = fmaxf(m1, s[n][2 * hf + 1]);
        }
        float mx = fmaxf(m0, m1);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int n = 0; n < NTP; ++n) {
          float p0 = 0.f, p1 = 0.f;
          if (n < NT) {
            p0 = fast_exp2(s[n][2 * hf] - mx);
            p1 = fast_exp2(s[n][2 * hf + 1] - mx);
          }
          l0 += p0;
          l1 += p1;
          uint32_t hi, lo;
          if (FP32) {
            const uint32_t u0 = __float_as_uint(p0), u1 = __float_as_uint(p1);
            hi = __byte_perm(u0, u1, 0x7632);
            lo = pack_bf16(p0 - __uint_as_float(u0 & 0xffff0000u), p1 - __uint_as_float(u1 & 0xffff0000u));
          } else {
            hi = pack_bf16(p0, p1);
            lo = 0u;
          }
          ph[n >> 1][2 * (n & 1) + hf] = hi;
          pl_[n >> 1][2 * (n & 1) + hf] = lo;
        }
        float l = l0 + l1;
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        linv[hf] = fast_rcp(l);
      }
      float o[4] = {0.f, 0.f, 0.f, 0.f}, o2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int grp = 0; grp < (KS + 1) / 2; ++grp) {   // V fragments of keys 32 grp .. 32 grp + 31 = k16 steps 2 grp, 2 grp + 1
        uint32_t vh[4], vl[4];
        const uint32_t ko = (uint32_t)min(32 * grp + 8 * mat + mrow, N - 1) * (PPT * 16);
        ang_ldsm4t(vs_u + ioff + ko, vh);
        if (FP32) ang_ldsm4t(vs_u + ioff + 2048u + ko, vl);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int t = 2 * grp + j;
          if (t < KS) {
            float* oo = j ? o2 : o;
            ang_hmma(oo, ph[t], vh[2 * j], vh[2 * j + 1]);
            if (FP32) {
              ang_hmma(oo, pl_[t], vh[2 * j], vh[2 * j + 1]);
              ang_hmma(oo, ph[t], vl[2 * j], vl[2 * j + 1]);
            }
          }
        }
      }
      uint8_t* obase = qo_half + rh * 4096 + (c >> 1) * 2048 + p * 16 + (c & 1) * 8;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int view = 16 * mt + 8 * hf + g;
        if (view < N)
          *reinterpret_cast<float2*>(obase + view * (PPT * 16)) =
              make_float2((o[2 * hf] + o2[2 * hf]) * linv[hf], (o[2 * hf + 1] + o2[2 * hf + 1]) * linv[hf]);
      }
    }
    rows_bar_sync256();
  }
  convert(1);
}

// NV > 0: compile-time number of views (scores kept in registers, single QK pass); NV == 0: runtime N.
template <int NV>
__global__ void __launch_bounds__(kThreads2, 2)
k_ang(const float* __restrict__ in, float* __restrict__ out, const uint8_t* __restrict__ wqk,
      const uint8_t* __restrict__ wv, const uint8_t* __restrict__ wo, const uint8_t* __restrict__ w1,
      const uint8_t* __restrict__ w2, const __grid_constant__ Tab512 tab, const float* __restrict__ peqk,
      const float* __restrict__ pe, int Nrt, int P, long long npix, int passes, int ntiles, Region rg) {
  // rg: the pixels of every patch view this launch computes (full view: {0, P}); npix = B * rg.rn^2 compacted pixels
  const int PP = P * P;
  const int N = NV > 0 ? NV : Nrt;
  // Specialised views-per-pixel counts (A = 3, 5, 7, 9) pair two views of one pixel in lanes l / l+16 and share every K/V read
  // between them (kPair); pixels per tile: 12 / 5 / 2 / 1, softmax chunk = A keys.
  constexpr bool kPair = NV == 9 || NV == 25 || NV == 49 || NV == 81;
  constexpr int kPPT = NV == 9 ? 12 : NV == 25 ? 5 : NV == 49 ? 2 : NV == 81 ? 1 : 0;
  constexpr int kCH = NV == 9 ? 9 : NV == 25 ? 5 : NV == 49 ? 7 : NV == 81 ? 9 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t R1 = s_base + kCtlBytes;  // 32 KB
  const uint32_t R2 = R1 + 32768;          // 32 KB
  const uint32_t ring = R2 + 32768;
  uint8_t* vs_ptr = smem + kCtlBytes;          // V (fp32, all heads) reuses R1 during the attention
  uint8_t* ks_ptr = smem + kCtlBytes + 32768;  // K (fp32, all heads) in R2
  const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t a_ready = smem_u32(&ctl->a_ready), mma_done = smem_u32(&ctl->mma_done);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t LBO = 128 * 16;

  pdl_trigger();
  cta_setup<kAngNST>(ctl, warp, lane, kRowThreads2, 256, kWarpMma2);
  const uint32_t tmem = ctl->tmem;

  const GemmPhase g_qk{wqk, 128, 1}, g_v{wv, 64, 1}, g_o{wo, 64, 1}, g_1{w1, 128, 1}, g_2{w2, 64, 2};

  if (warp == kWarpProducer2) {

    RingState<kAngNST> rs;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // the ring runs ahead into the next tile's first slabs
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_qk, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_v, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_o, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_1, passes);
      ring_produce<kAngNST>(rs, ring, kAngStage, full0, empty0, g_2, passes);
    }
  } else if (warp == kWarpMma2) {

    RingState<kAngNST> rs;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // a_ready / mma_done complete four phases per tile, so the literal parities 0,1,0,1 below hold for every tile; the
    // next tile's first MMA is gated by a_ready arrivals the row owners make only after their phase 4 reads of this tile
    mbar_wait(a_ready, 0);
    tc_fence_after();
    ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_qk, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                              tmem + 0, true);
    ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_v, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                              tmem + 128, true);
    umma_commit_elected(mma_done);
    mbar_wait(a_ready, 1);
    tc_fence_after();
    if constexpr (kPair)  // O operand in TMEM columns [64,96) hi | [96,128) lo (TS form)
      ring_consume_mma_ts<kAngNST>(rs, ring, kAngStage, full0, empty0, g_o, passes, tmem + 64, tmem + 96, tmem + 0, true);
    else
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_o, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                                tmem + 0, true);
    umma_commit_elected(mma_done);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    if constexpr (kPair)  // X1 operand in TMEM columns [128,160) hi | [160,192) lo (the V accumulator is dead)
      ring_consume_mma_ts<kAngNST>(rs, ring, kAngStage, full0, empty0, g_1, passes, tmem + 128, tmem + 160, tmem + 0, true);
    else
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_1, passes, R1, R1 + 16384, LBO, 0, NoShift{},
                                tmem + 0, true);
    umma_commit_elected(mma_done);
    mbar_wait(a_ready, 1);
    tc_fence_after();
    if constexpr (kPair)  // hidden operand written in place over the FFN1 accumulator: [0,64) hi | [64,128) lo
      ring_consume_mma_ts<kAngNST>(rs, ring, kAngStage, full0, empty0, g_2, passes, tmem + 0, tmem + 64, tmem + 128, true);
    else
      ring_consume_mma<kAngNST>(rs, ring, kAngStage, full0, empty0, g_2, passes, R1, R2, LBO, 8 * LBO, NoShift{},
                                tmem + 128, true);
    umma_commit_elected(mma_done);
    }  // tiles
  } else {
    // ------------------------------------------------------------ row owner: row m, channel half q
    const int m = (warp & 3) * 32 + lane, q = warp >> 2;
    // rows are view-major inside the tile (m = a*PPT + pl): consecutive lanes = consecutive pixels of one view,
    // so one 16-byte request of a warp touches ~7 cache lines of the T32 layout instead of 25
    const int PPT = kPair ? kPPT : (NV > 0 ? 128 / NV : 128 / N);
    int a, pl, kvrow;
    if constexpr (kPair && NV != 25) {
      // pair k = 16*(warp&3) + (lane&15): view pair k / PPT, pixel k % PPT; lower lane = even view, upper lane = odd view
      // (an odd view count leaves the last even view paired with an idle row)
      const int k = 16 * (warp & 3) + (lane & 15), upper = lane >> 4;
      const int vp = k / kPPT;
      pl = k - kPPT * vp;
      a = 2 * vp + upper;
      kvrow = a < NV ? a * kPPT + pl : -1;  // idle rows stage nothing
    } else if constexpr (NV == 25) {
      // paired rows: lanes l and l+16 of a warp hold views (2v, 2v+1) of ONE pixel, so that the attention can share every
      // K/V read between two queries (partner = lane ^ 16; a quarter-warp always reads one head: no bank conflicts).
      // Pair k = 16*(warp&3) + (lane&15) < 60: view pair k/5, pixel k%5.  The 8 remaining lanes of the last row warp
      // hold the 5 rows of view 24 (slots 0..4) and 3 idle rows.
      {
        const int k = 16 * (warp & 3) + (lane & 15), upper = lane >> 4;
        if (k < 60) {
          const int vp = k / 5;
          pl = k - 5 * vp;
          a = 2 * vp + upper;
        } else {
          const int slot = (k - 60) + 4 * upper;
          a = slot < 5 ? 24 : 25;
          pl = slot < 5 ? slot : slot - 5;
        }
      }
      kvrow = a < 25 ? a * 5 + pl : m;  // K/V planes stay view-major: one wavefront per key read
    } else {
      a = m / PPT;
      pl = m - a * PPT;
      kvrow = m;
    }
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int bar_id = 1 + (warp & 3);
    // token of this row in tile t (-1: idle row / beyond the last pixel)
    auto row_token = [&](int t) -> long long {
      const long long gp = (long long)t * PPT + pl;
      if (a >= N || gp >= npix) return -1;
      const unsigned RR = (unsigned)(rg.rn * rg.rn);
      const unsigned b = (unsigned)gp / RR;  // npix < 2^31
      const int rem = (int)((unsigned)gp - b * RR);
      const int yy = rem / rg.rn;
      const int p = (rg.r0 + yy) * P + rg.r0 + rem - yy * rg.rn;
      return ((long long)b * N + a) * PP + p;
    };
    // persistent: the CTA walks over the tiles blockIdx.x, blockIdx.x + gridDim.x, ... keeping its TMEM, barriers and the
    // weight ring (which already holds the next tile's first slabs when its phase 0 publishes)
    pdl_wait();  // `in` is the previous kernel's output
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long tok_or = row_token(tile);
    const bool rowok = tok_or >= 0;
    const long long tok = rowok ? tok_or : 0;
    const int aa = rowok ? a : 0;
    float mean, rstd;

    LFT_TL(0);
    // ---- phase 0: load X (own 32 channels), stash in TMEM, LN1 statistics of X+PE, raw X -> R1
    {
      float x[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#ifdef LFT_X_ANG_NOLOAD
        const float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#else
        const float4 f = rowok ? __ldg(reinterpret_cast<const float4*>(in + t32_off(tok, 8 * q + i, 16)))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
        x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
      }
      tmem_st16(trow + 192 + 32 * q, x);
      tmem_st16(trow + 192 + 32 * q + 16, x + 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        split8(x + 8 * c, hi, lo, passes == 3);
        st_shared_v4(R1 + (4 * q + c) * LBO + m * 16, hi);
        st_shared_v4(R1 + 16384 + (4 * q + c) * LBO + m * 16, lo);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(pe) + (8 * q + i) * N + aa);  // [chunk 16][N][4]
        x[4 * i] += f.x; x[4 * i + 1] += f.y; x[4 * i + 2] += f.z; x[4 * i + 3] += f.w;
      }
      pair_ln_stats<32>(x, trow + 4 * q, trow + 4 * (1 - q), bar_id, mean, rstd);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    LFT_TL(1);
    // ---- phase 1: attention. K (affine-corrected) -> R2, V -> R1 (fp32, all 8 heads), Q in registers
    mbar_wait(mma_done, 0);
    tc_fence_after();
    LFT_TL(2);
    {
      const float mr = mean * rstd;
      const float4* pq4 = reinterpret_cast<const float4*>(peqk) + aa;  // [chunk 32][N][4]: Q chunks 0..15, K 16..31
      const float4* tab4 = reinterpret_cast<const float4*>(tab.v);    // [u_qk 128 | c_qk 128 | u_1 128 | c_1 128] (constant bank)
      if constexpr (NV == 25 || NV == 49 || NV == 81) {
#ifndef LFT_ANG_ITEMS_V1
        if constexpr (NV == 25) {
#ifndef LFT_X_ANG_NOATTN2   // (defined: timing experiment - no exports, no items, no convert)
          if (passes == 3) ang_attention_mma25<true>(trow, warp, lane, q, kvrow, rstd, mr, pq4, tab4, smem + kCtlBytes);
          else ang_attention_mma25<false>(trow, warp, lane, q, kvrow, rstd, mr, pq4, tab4, smem + kCtlBytes);
#endif
        } else {
          if (passes == 3) ang_attention_mma_g<true, NV, kPPT>(trow, warp, lane, q, kvrow, rstd, mr, pq4, tab4, smem + kCtlBytes);
          else ang_attention_mma_g<false, NV, kPPT>(trow, warp, lane, q, kvrow, rstd, mr, pq4, tab4, smem + kCtlBytes);
        }
#else
        ang_attention_items<NV, kPPT, kCH>(trow, warp, lane, q, kvrow, rstd, mr, pq4, tab4, smem + kCtlBytes, passes == 3);
#endif
        LFT_TL(4);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(a_ready);
      } else {
      float kv[16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // K columns 64 + 32q + 16c
        const int col = 64 + 32 * q + 16 * c;
        tmem_ld16(trow + col, kv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 pv = __ldg(pq4 + (col / 4 + j) * N);
          const float4 uv = tab4[col / 4 + j], cv = tab4[32 + col / 4 + j];
          kv[4 * j] = fmaf(rstd, kv[4 * j] + pv.x, fmaf(-mr, uv.x, cv.x));
          kv[4 * j + 1] = fmaf(rstd, kv[4 * j + 1] + pv.y, fmaf(-mr, uv.y, cv.y));
          kv[4 * j + 2] = fmaf(rstd, kv[4 * j + 2] + pv.z, fmaf(-mr, uv.z, cv.z));
          kv[4 * j + 3] = fmaf(rstd, kv[4 * j + 3] + pv.w, fmaf(-mr, uv.w, cv.w));
        }
        if (kvrow >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)  // [head][half][row][4 floats]: head = 4q + 2c + j/2, half = j%2 (conflict-free)
            *reinterpret_cast<float4*>(ks_ptr + ((4 * q + 2 * c + (j >> 1)) * 2 + (j & 1)) * 2048 + kvrow * 16) =
                make_float4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // V columns 128 + 32q + 16c (raw)
        tmem_ld16(trow + 128 + 32 * q + 16 * c, kv);
        if (kvrow >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(vs_ptr + ((4 * q + 2 * c + (j >> 1)) * 2 + (j & 1)) * 2048 + kvrow * 16) =
                make_float4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
        }
      }
      float qv[32];
      tmem_ld16_nowait(trow + 32 * q, qv);
      tmem_ld16_nowait(trow + 32 * q + 16, qv + 16);
      tmem_wait_ld();
      const float scale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e)/sqrt(8), folded into Q (softmax via exp2)
      if constexpr (kPair) {
        // corrected, pre-scaled Q back into its own TMEM columns (fp32 stash); the pair kernels re-read one head pair at a time
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 pv = __ldg(pq4 + (8 * q + j) * N);
          const float4 uv = tab4[8 * q + j], cv = tab4[32 + 8 * q + j];
          qv[4 * j] = scale * fmaf(rstd, qv[4 * j] + pv.x, fmaf(-mr, uv.x, cv.x));
          qv[4 * j + 1] = scale * fmaf(rstd, qv[4 * j + 1] + pv.y, fmaf(-mr, uv.y, cv.y));
          qv[4 * j + 2] = scale * fmaf(rstd, qv[4 * j + 2] + pv.z, fmaf(-mr, uv.z, cv.z));
          qv[4 * j + 3] = scale * fmaf(rstd, qv[4 * j + 3] + pv.w, fmaf(-mr, uv.w, cv.w));
        }
        tmem_st16(trow + 32 * q, qv);
        tmem_st16(trow + 32 * q + 16, qv + 16);
        tmem_wait_st();
        rows_bar_sync256();  // K/V planes complete; every K/V accumulator column has been consumed
        LFT_TL(3);
        const int wq = warp & 3;
        const bool up = lane >= 16;
        int partner = lane ^ 16;
        if (NV == 25 && wq == 3 && (lane & 15) >= 12) partner = lane;  // rows of view 24 / idle rows: no partner view
        const uint8_t* ks_q = ks_ptr + (4 * q) * 4096;
        const uint8_t* vs_q = vs_ptr + (4 * q) * 4096;
        const uint32_t tq = trow + 32 * q, to_hi = trow + 64 + 16 * q, to_lo = trow + 96 + 16 * q;
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2)
          ang_pair_heads<NV, kPPT, kCH>(tq, to_hi, to_lo, h2, up, partner, ks_q, vs_q, pl, passes == 3);
        if (NV == 25 && wq == 3) ang_singles25(tq, to_hi, to_lo, lane, ks_q, vs_q, passes == 3);
        LFT_TL(4);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(a_ready);
      } else {
      f32x2 q2[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 pv = __ldg(pq4 + (8 * q + j) * N);
        const float4 uv = tab4[8 * q + j], cv = tab4[32 + 8 * q + j];
        q2[2 * j] = pack2(scale * fmaf(rstd, qv[4 * j] + pv.x, fmaf(-mr, uv.x, cv.x)),
                          scale * fmaf(rstd, qv[4 * j + 1] + pv.y, fmaf(-mr, uv.y, cv.y)));
        q2[2 * j + 1] = pack2(scale * fmaf(rstd, qv[4 * j + 2] + pv.z, fmaf(-mr, uv.z, cv.z)),
                              scale * fmaf(rstd, qv[4 * j + 3] + pv.w, fmaf(-mr, uv.w, cv.w)));
      }
      rows_bar_sync256();
      LFT_TL(3);
      float o[32];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        f32x2 acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = 0ull;
        float l = 1.f;
        if (a < N) {
          // K/V of this pixel: rows m' = t*PPT + pl of the [head][half][row][4 floats] planes; the PPT pixels
          // a warp touches for one key are 16*PPT contiguous bytes -> one wavefront per load
          const ulonglong2* kb = reinterpret_cast<const ulonglong2*>(ks_ptr + (4 * q + hh) * 4096) + pl;
          const ulonglong2* vb = reinterpret_cast<const ulonglong2*>(vs_ptr + (4 * q + hh) * 4096) + pl;
          const int ts = PPT;
          float mx = -INFINITY;
          l = 0.f;
          if constexpr (NV > 0 && NV <= 32) {
            float sc[NV];
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              sc[t] = dot8(q2 + 4 * hh, kb[ts * t], kb[ts * t + 128]);
              mx = fmaxf(mx, sc[t]);
            }
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              const float pw = fast_exp2(sc[t] - mx);
              l += pw;
              const f32x2 pp = pack2(pw, pw);
              const ulonglong2 v0 = vb[ts * t], v1 = vb[ts * t + 128];
              acc[0] = fma2(pp, v0.x, acc[0]); acc[1] = fma2(pp, v0.y, acc[1]);
              acc[2] = fma2(pp, v1.x, acc[2]); acc[3] = fma2(pp, v1.y, acc[3]);
            }
          } else {
#pragma unroll 4
            for (int t = 0; t < N; ++t) mx = fmaxf(mx, dot8(q2 + 4 * hh, kb[ts * t], kb[ts * t + 128]));
#pragma unroll 4
            for (int t = 0; t < N; ++t) {
              const float pw = fast_exp2(dot8(q2 + 4 * hh, kb[ts * t], kb[ts * t + 128]) - mx);
              l += pw;
              const f32x2 pp = pack2(pw, pw);
              const ulonglong2 v0 = vb[ts * t], v1 = vb[ts * t + 128];
              acc[0] = fma2(pp, v0.x, acc[0]); acc[1] = fma2(pp, v0.y, acc[1]);
              acc[2] = fma2(pp, v1.x, acc[2]); acc[3] = fma2(pp, v1.y, acc[3]);
            }
          }
        }
        const float inv = 1.f / l;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float a, b;
          unpack2(acc[e], a, b);
          o[8 * hh + 2 * e] = a * inv;
          o[8 * hh + 2 * e + 1] = b * inv;
        }
      }
      LFT_TL(4);
      rows_bar_sync256();  // everyone is done reading K/V: R1 can take the O operand
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        split8(o + 8 * c, hi, lo, passes == 3);
        st_shared_v4(R1 + (4 * q + c) * LBO + m * 16, hi);
        st_shared_v4(R1 + 16384 + (4 * q + c) * LBO + m * 16, lo);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
      }  // !kPair
      }  // NV != 25
    }

    LFT_TL(5);
    // ---- phase 2: X1 = X + O Wo^T (own 32 channels; stash), LN2 statistics, raw X1 -> R1
    mbar_wait(mma_done, 1);
    tc_fence_after();
    LFT_TL(6);
    {
      float d[32], x[32];
      tmem_ld16_nowait(trow + 32 * q, d);
      tmem_ld16_nowait(trow + 32 * q + 16, d + 16);
      tmem_ld16_nowait(trow + 192 + 32 * q, x);
      tmem_ld16_nowait(trow + 192 + 32 * q + 16, x + 16);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] += d[i];
      tmem_st16(trow + 192 + 32 * q, x);
      tmem_st16(trow + 192 + 32 * q + 16, x + 16);
      if constexpr (kPair) {  // TS form: the operand never touches shared memory
        a_tmem_store16(trow + 128, trow + 160, 32 * q, x, passes == 3);
        a_tmem_store16(trow + 128, trow + 160, 32 * q + 16, x + 16, passes == 3);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 hi, lo;
          split8(x + 8 * c, hi, lo, passes == 3);
          st_shared_v4(R1 + (4 * q + c) * LBO + m * 16, hi);
          st_shared_v4(R1 + 16384 + (4 * q + c) * LBO + m * 16, lo);
        }
      }
      pair_ln_stats<32>(x, trow + 64 + 4 * q, trow + 64 + 4 * (1 - q), bar_id, mean, rstd);  // includes tcgen05.wait::st
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    LFT_TL(7);
    {  // the next tile's X rows: pull them into L2 while FFN1 runs (phase 0 of the next tile then misses only L1)
      const int nt = tile + (int)gridDim.x;
      const long long ntok = nt < ntiles ? row_token(nt) : -1;
      if (ntok >= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) prefetch_l2(in + t32_off(ntok, 8 * q + i, 16));
      }
    }
    // ---- phase 3: hidden = relu(LN2-folded D[0,128)), own 64 columns -> K=128 operand (hi in R1, lo in R2)
    mbar_wait(mma_done, 0);
    tc_fence_after();
    LFT_TL(8);
    {
      const float mr = mean * rstd;
      float dd[64];  // all four accumulator loads in flight, one wait
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(trow + 64 * q + 16 * c, dd + 16 * c);
      tmem_wait_ld();
      if constexpr (kPair) {  // the operand goes back into the same columns: both threads of the row must have read first
        tc_fence_before();
        pair_bar_sync(warp & 3);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* d = dd + 16 * c;
        const int col = 64 * q + 16 * c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 uv = reinterpret_cast<const float4*>(tab.v + 256 + col)[j];
          const float4 cv = reinterpret_cast<const float4*>(tab.v + 384 + col)[j];
          d[4 * j] = fmaxf(fmaf(rstd, d[4 * j], fmaf(-mr, uv.x, cv.x)), 0.f);
          d[4 * j + 1] = fmaxf(fmaf(rstd, d[4 * j + 1], fmaf(-mr, uv.y, cv.y)), 0.f);
          d[4 * j + 2] = fmaxf(fmaf(rstd, d[4 * j + 2], fmaf(-mr, uv.z, cv.z)), 0.f);
          d[4 * j + 3] = fmaxf(fmaf(rstd, d[4 * j + 3], fmaf(-mr, uv.w, cv.w)), 0.f);
        }
        if constexpr (kPair) {  // TS form: hidden hi -> columns [0,64), lo -> [64,128)
          a_tmem_store16(trow + 0, trow + 64, 64 * q + 16 * c, d, passes == 3);
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 hi, lo;
            split8(d + 8 * j, hi, lo, passes == 3);
            st_shared_v4(R1 + (8 * q + 2 * c + j) * LBO + m * 16, hi);
            st_shared_v4(R2 + (8 * q + 2 * c + j) * LBO + m * 16, lo);
          }
        }
      }
      if constexpr (kPair) tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    LFT_TL(9);
    // ---- phase 4: X2 = X1 + D[128,192) -> global (own 32 channels)
    mbar_wait(mma_done, 1);
    tc_fence_after();
    LFT_TL(10);
    {
      float d[32], x[32];
      tmem_ld16_nowait(trow + 128 + 32 * q, d);
      tmem_ld16_nowait(trow + 128 + 32 * q + 16, d + 16);
      tmem_ld16_nowait(trow + 192 + 32 * q, x);
      tmem_ld16_nowait(trow + 192 + 32 * q + 16, x + 16);
      tmem_wait_ld();
      if (rowok) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(out + t32_off(tok, 8 * q + i, 16)) =
              make_float4(x[4 * i] + d[4 * i], x[4 * i + 1] + d[4 * i + 1], x[4 * i + 2] + d[4 * i + 2],
                          x[4 * i + 3] + d[4 * i + 3]);
      }
    }
    tc_fence_before();
    LFT_TL(11);
    }  // tiles
  }
  cta_teardown(ctl, warp, 256, kWarpMma2);
}

int debug_timeline_ang(long long* out) {
#ifdef LFT_TIMELINE
  CUDA_TRY(cudaMemcpyFromSymbol(out, g_tl, sizeof(long long) * 64));
  return 0;
#else
  (void)out;
  return fail(LFT_ERR_STATE, "library built without -DLFT_TIMELINE");
#endif
}

int configure_ang() {
  CUDA_TRY(cudaFuncSetAttribute(k_ang<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  CUDA_TRY(cudaFuncSetAttribute(k_ang<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  CUDA_TRY(cudaFuncSetAttribute(k_ang<49>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  CUDA_TRY(cudaFuncSetAttribute(k_ang<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  CUDA_TRY(cudaFuncSetAttribute(k_ang<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemAng));
  return 0;
}

int run_ang(Handle* h, int layer, const float* in, float* out, int B, int P, Region rg, cudaStream_t st) {
  const int N = h->cfg.ang_res * h->cfg.ang_res;
  const long long npix = (long long)B * rg.rn * rg.rn;
  const int PPT = N == 9 ? 12 : N == 25 ? 5 : N == 49 ? 2 : N == 81 ? 1 : 128 / N;  // keep in step with kPPT in k_ang
  const unsigned ntiles = (unsigned)((npix + PPT - 1) / PPT);
  const unsigned grid = ntiles < 2u * h->num_sms ? ntiles : 2u * h->num_sms;  // persistent: two CTAs per SM
  const Layer& L = h->layer[layer];
  Tab512 ta;
  memcpy(ta.v, L.a_tab[h->mode()].data(), sizeof(ta.v));
  Scope sc(h, K_ANG, st, npix * N);
#define LFT_ANG_LAUNCH(NV)                                                                                          \
  do {                                                                                                              \
    auto kern = k_ang<NV>;                                                                                          \
    LFT_LAUNCH(h, kern, grid, kThreads2, kSmemAng, st, in, out, L.a_wqk, L.a_wv, L.a_wo, L.a_w1, L.a_w2, ta,        \
               L.a_peqk[h->mode()], h->pe_ang, N, P, npix, h->passes(), (int)ntiles, rg);                           \
  } while (0)
  if (N == 25) LFT_ANG_LAUNCH(25);
  else if (N == 9) LFT_ANG_LAUNCH(9);
  else if (N == 49) LFT_ANG_LAUNCH(49);
  else if (N == 81) LFT_ANG_LAUNCH(81);
  else LFT_ANG_LAUNCH(0);
#undef LFT_ANG_LAUNCH
  return sc.finish();
}

}  // namespace lft
