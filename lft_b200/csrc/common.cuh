// Blackwell (sm_100a) device primitives shared by the LFT kernels: mbarrier, bulk async copy
// (TMA engine, 1-D), tcgen05 MMA / TMEM alloc / ld / st, UMMA descriptors, bf16 hi/lo splitting.
//
// Operand layout used everywhere ("chunk-major, no swizzle"):
//   element (row r, k) of a K-major bf16 operand lives at  base + (k/8)*LBO + r*16 + (k%8)*2
//   i.e. 8-row x 16-byte core matrices are contiguous (128 B), consecutive rows are 16 B apart,
//   SBO (8-row group stride) = 128 B and LBO (k-chunk stride) = ROWS*16 B.  Because the row stride
//   is uniform, a descriptor whose start address is moved by d*16 B addresses the same tile shifted
//   by d rows - which is how the 3x3 convolutions read their nine taps without an im2col copy.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace lft {

#define LFT_DEVINL __device__ __forceinline__

LFT_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
LFT_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
LFT_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
LFT_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
LFT_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
LFT_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
LFT_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// One elected lane of a fully active warp.  MMA / bulk-copy issue loops run warp-uniformly and predicate only the
// issuing instruction on this, so descriptors stay in uniform registers (issuing from inside `if (lane == 0)`
// makes the compiler wrap every tcgen05.mma in a per-lane R2UR "waterfall" loop, ~100 cycles per MMA).
LFT_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- proxies / fences
LFT_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LFT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LFT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
LFT_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// pair barrier for the two warps sharing TMEM lane quarter `quarter` (ids 1..4, compile-time so that the
// kernel only reserves the barriers it uses)
LFT_DEVINL void pair_bar_sync(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
LFT_DEVINL void rows_bar_sync256() { asm volatile("bar.sync 5, 256;" ::: "memory"); }

// ---------------------------------------------------------------- bulk async copy global -> smem
LFT_DEVINL void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Programmatic dependent launch (launch attribute programmaticStreamSerialization, set by LFT_LAUNCH in host.h): a kernel
// lets the next kernel of the stream start its CTAs - barrier set-up, TMEM allocation, first weight slabs - while it is still
// running (pdl_trigger, called by every thread at the top), and every thread of the next kernel calls pdl_wait before its
// first access to global memory another kernel may have written: the wait returns once the preceding kernel has completed
// and its writes are visible.  Without the attribute both instructions are no-ops - which is the default: the mechanism is
// correct (all tests pass with LFT_PDL=1) but measured slower on this workload (host.h, LFT_LAUNCH).
LFT_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
LFT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

LFT_DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------- TMEM
LFT_DEVINL void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LFT_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
LFT_DEVINL void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
LFT_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (lane_base+i), columns c..c+15
LFT_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
LFT_DEVINL void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}

LFT_DEVINL void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
LFT_DEVINL void tmem_st4(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
LFT_DEVINL void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
LFT_DEVINL uint4 tmem_ld4u(uint32_t taddr) {
  uint4 r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return r;
}
LFT_DEVINL void tmem_st4u(uint32_t taddr, uint4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
LFT_DEVINL void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 32 consecutive columns without the trailing wait (caller batches several loads, then tmem_wait_ld()).
LFT_DEVINL void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors + MMA
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE, sm_100 version field = 1.
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor; LBO = k-chunk stride, SBO = 8-row stride)
LFT_DEVINL uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor: kind::f16, A=B=bf16, D=f32, both K-major, M=128, N (multiple of 16).
LFT_DEVINL uint32_t umma_idesc_bf16(uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}
LFT_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TS form: A operand from tensor memory (lane = row, 32-bit column c = bf16 elements 2c (low half), 2c+1), B from smem.
LFT_DEVINL void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
LFT_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// commit issued by the elected lane of a fully active warp (same lane that issued the MMAs)
LFT_DEVINL void umma_commit_elected(uint32_t bar) {
  if (elect_one()) umma_commit(bar);
  __syncwarp();
}

// ---------------------------------------------------------------- bf16 split
// x = hi + lo (+ O(2^-17 |x|)); hi = bf16 truncation of x, lo = bf16_rn(x - hi).  3 MMAs (hi*hi, lo*hi, hi*lo)
// reproduce an fp32 product to ~2^-16 relative.
struct bf16x8 { uint4 v; };
LFT_DEVINL uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
LFT_DEVINL void split8(const float* x, uint4& hi, uint4& lo, bool fp32_mode) {
  // fp32 mode: hi = x truncated to its top 16 bits (one PRMT per pair), lo = bf16_rn(x - hi) (exact subtraction):
  // |x - hi - lo| <= 2^-17 |x|.  bf16 mode (no lo pass): hi = bf16_rn(x), unbiased.
  uint32_t h[4], l[4];
  if (!fp32_mode) {
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t a = __float_as_uint(x[2 * i]), b = __float_as_uint(x[2 * i + 1]);
    h[i] = __byte_perm(a, b, 0x7632);
    // x - trunc(x) for the pair in one packed FFMA (exact: the same bits as two subtractions)
    unsigned long long t2, x2, r2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "r"(a & 0xffff0000u), "r"(b & 0xffff0000u));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "r"(a), "r"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(t2), "l"(0xbf800000bf800000ull), "l"(x2));
    float r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r2));
    l[i] = pack_bf16(r0, r1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
LFT_DEVINL void st_shared_v4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------- weight-stream ring (B operand)
// Weights are pre-packed on the host into "slabs": one slab = [N rows x 64 k] bf16 in the chunk-major
// layout above (N*128 bytes), hi slab followed by lo slab for every 64-wide k block.  A producer thread
// streams slabs through an NST-deep ring with bulk copies; the MMA thread consumes them in order and
// frees each stage with tcgen05.commit.
template <int NST>
struct RingState {
  int stage = 0;
  uint32_t phase = 0;
  LFT_DEVINL void advance() {
    if (++stage == NST) { stage = 0; phase ^= 1; }
  }
};

struct GemmPhase {
  const uint8_t* w;   // packed weights for this GEMM: k-slab ks at w + ks*2*N*128 (hi then lo)
  uint32_t N;         // output columns (multiple of 16, <= 256)
  uint32_t kslabs;    // K / 64
};

// Producer side of one GEMM phase (called by ALL lanes of the producer warp). `passes` = 3 (fp32 via hi/lo) or 1
// (bf16: hi slabs only).
#ifdef LFT_TIMELINE
// producer timeline of the middle CTA (debug): [kind][slab][0: before the empty wait, 1: after it, 2: after the copy issue]
static __device__ long long g_ring_tl[4][40][3];
static __device__ int g_ring_tl_n[4];
#endif
template <int NST>
LFT_DEVINL void ring_produce(RingState<NST>& rs, uint32_t ring_base, uint32_t stage_bytes, uint32_t full0,
                             uint32_t empty0, const GemmPhase& g, int passes, int tlk = -1, int* tln = nullptr) {
  const uint32_t slab = g.N * 128u;
  for (uint32_t ks = 0; ks < g.kslabs; ++ks) {
    const int nparts = passes == 3 ? 2 : 1;
    for (int part = 0; part < nparts; ++part) {
#ifdef LFT_TIMELINE
      const bool rec = tlk >= 0 && tln && blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 0 && *tln < 40;
      if (rec) g_ring_tl[tlk][*tln][0] = clock64();
#endif
      mbar_wait(empty0 + 8u * rs.stage, rs.phase ^ 1u);
#ifdef LFT_TIMELINE
      if (rec) g_ring_tl[tlk][*tln][1] = clock64();
#endif
      if (elect_one()) {
#ifdef LFT_EXPERIMENT_NOSTREAM  // timing experiment only (wrong results): how much does weight streaming cost?
        mbar_arrive(full0 + 8u * rs.stage);
#else
        mbar_arrive_expect_tx(full0 + 8u * rs.stage, slab);
        bulk_g2s(ring_base + rs.stage * stage_bytes, g.w + (size_t)(ks * 2 + part) * slab, slab, full0 + 8u * rs.stage);
#endif
      }
      __syncwarp();
#ifdef LFT_TIMELINE
      if (rec) { g_ring_tl[tlk][*tln][2] = clock64(); g_ring_tl_n[tlk] = *tln + 1; }
      if (tln) ++*tln;
#endif
      rs.advance();
    }
  }
}

// MMA side of one GEMM phase: D[128 x N] (+)= A[128 x K] * W[N x K]^T.
//   a_hi/a_lo : smem byte addresses of the A operand (row 0, k 0), a_lbo its k-chunk stride in bytes.
//   a_kslab_stride: byte distance between consecutive 64-wide k slabs of A (= 8*a_lbo for a plain
//   operand; 0 together with per-slab `row_shift` for the conv taps).
//   row_shift(ks): rows to shift A by for slab ks (conv taps), else 0.
// Descriptors are built once per phase/stage and advanced by adding to their low word (the start-address
// field never carries: shared addresses are < 256 KB).
LFT_DEVINL uint64_t umma_desc_from(uint32_t lo32) { return ((uint64_t)0x4008u << 32) | lo32; }  // SBO=128 B, version 1
LFT_DEVINL uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return (saddr >> 4) | ((lbo_bytes >> 4) << 16); }

// Called by ALL lanes of the MMA warp; one elected lane issues the MMAs and the commits.
template <int NST, typename ShiftFn>
LFT_DEVINL void ring_consume_mma(RingState<NST>& rs, uint32_t ring_base, uint32_t stage_bytes, uint32_t full0,
                                 uint32_t empty0, const GemmPhase& g, int passes, uint32_t a_hi, uint32_t a_lo,
                                 uint32_t a_lbo, uint32_t a_kslab_stride, ShiftFn row_shift, uint32_t d_tmem,
                                 bool fresh) {
  const uint32_t idesc = umma_idesc_bf16(g.N);
  const uint32_t b_lbo = g.N * 16u;
  const uint32_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;  // one K=16 step, in 16-byte units
  const uint32_t ahi0 = umma_desc_lo(a_hi, a_lbo), alo0 = umma_desc_lo(a_lo, a_lbo);
  uint32_t acc = fresh ? 0u : 1u;
  for (uint32_t ks = 0; ks < g.kslabs; ++ks) {
    const int sh = row_shift(ks);
    // offset in 16-byte rows; a negative conv-tap shift is a plain 32-bit subtraction from the start-address
    // field (callers keep >= kConvOff rows of headroom, so it never borrows from the LBO field)
    const uint32_t a_off = (uint32_t)((int)((ks * a_kslab_stride) >> 4) + sh);
    const uint32_t ah = ahi0 + a_off, al = alo0 + a_off;
    // The issuing thread pays ~250 cycles of bookkeeping per wait / elect / commit round (measured), as much as 4 MMAs:
    // in fp32 mode the hi and the lo slab of a k block are therefore waited for together and their 12 MMAs
    // (A_hi*W_hi, A_lo*W_hi | A_hi*W_lo) go out in ONE elected region, each stage still released by its own commit.
    const uint32_t st_hi = rs.stage, ph_hi = rs.phase;
    rs.advance();
    mbar_wait(full0 + 8u * st_hi, ph_hi);
    const uint32_t bh = umma_desc_lo(ring_base + st_hi * stage_bytes, b_lbo);
    if (passes == 3) {
      const uint32_t st_lo = rs.stage, ph_lo = rs.phase;
      rs.advance();
      mbar_wait(full0 + 8u * st_lo, ph_lo);
      tc_fence_after();
      const uint32_t bl = umma_desc_lo(ring_base + st_lo * stage_bytes, b_lbo);
      if (elect_one()) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16(d_tmem, umma_desc_from(ah + j * a_step), umma_desc_from(bh + j * b_step), idesc, j ? 1u : acc);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16(d_tmem, umma_desc_from(al + j * a_step), umma_desc_from(bh + j * b_step), idesc, 1u);
        umma_commit(empty0 + 8u * st_hi);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16(d_tmem, umma_desc_from(ah + j * a_step), umma_desc_from(bl + j * b_step), idesc, 1u);
        umma_commit(empty0 + 8u * st_lo);
      }
    } else {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16(d_tmem, umma_desc_from(ah + j * a_step), umma_desc_from(bh + j * b_step), idesc, j ? 1u : acc);
        umma_commit(empty0 + 8u * st_hi);
      }
    }
    __syncwarp();
    acc = 1u;
  }
}

// One K=64 GEMM step against a B operand that is resident in shared memory (hi at b_hi, lo at b_lo):
// D[128 x N] = A[128 x 64] * B[N x 64]^T with the usual 1 or 3 passes.
LFT_DEVINL void mma_resident64(uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t b_hi, uint32_t b_lo, uint32_t N,
                               uint32_t d_tmem, int passes) {
  const uint32_t idesc = umma_idesc_bf16(N);
  const uint32_t b_lbo = N * 16u;
  const uint32_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;
  const uint32_t ah = umma_desc_lo(a_hi, a_lbo), al = umma_desc_lo(a_lo, a_lbo);
  const uint32_t bh = umma_desc_lo(b_hi, b_lbo), bl = umma_desc_lo(b_lo, b_lbo);
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j)
    umma_bf16(d_tmem, umma_desc_from(ah + j * a_step), umma_desc_from(bh + j * b_step), idesc, j ? 1u : 0u);
  if (passes == 3) {
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      umma_bf16(d_tmem, umma_desc_from(al + j * a_step), umma_desc_from(bh + j * b_step), idesc, 1u);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      umma_bf16(d_tmem, umma_desc_from(ah + j * a_step), umma_desc_from(bl + j * b_step), idesc, 1u);
  }
}

// 16 fp32 values of this thread's row -> bf16 hi/lo pairs in the TMEM A operand (TS form): elements k0..k0+15 occupy
// the 8 columns k0/2.. of the hi and of the lo operand (lane = row).
LFT_DEVINL void a_tmem_store16(uint32_t t_hi, uint32_t t_lo, int k0, const float* x, bool fp32_mode) {
  uint4 h0, l0, h1, l1;
  split8(x, h0, l0, fp32_mode);
  split8(x + 8, h1, l1, fp32_mode);
  const uint32_t hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  tmem_st8(t_hi + (k0 >> 1), hv);
  if (fp32_mode) {
    const uint32_t lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
    tmem_st8(t_lo + (k0 >> 1), lv);
  }
}

// TS-form GEMM phase: D[128 x N] (+)= A(tmem)[128 x K] * W[N x K]^T, weights through the ring (all lanes call it).
template <int NST>
LFT_DEVINL void ring_consume_mma_ts(RingState<NST>& rs, uint32_t ring_base, uint32_t stage_bytes, uint32_t full0,
                                    uint32_t empty0, const GemmPhase& g, int passes, uint32_t ta_hi, uint32_t ta_lo,
                                    uint32_t d_tmem, bool fresh) {
  const uint32_t idesc = umma_idesc_bf16(g.N);
  const uint32_t b_lbo = g.N * 16u, b_step = (2u * b_lbo) >> 4;
  uint32_t acc = fresh ? 0u : 1u;
  for (uint32_t ks = 0; ks < g.kslabs; ++ks) {
    const uint32_t st_hi = rs.stage, ph_hi = rs.phase;  // hi and lo slab in one round, see ring_consume_mma
    rs.advance();
    mbar_wait(full0 + 8u * st_hi, ph_hi);
    const uint32_t bh = umma_desc_lo(ring_base + st_hi * stage_bytes, b_lbo);
    if (passes == 3) {
      const uint32_t st_lo = rs.stage, ph_lo = rs.phase;
      rs.advance();
      mbar_wait(full0 + 8u * st_lo, ph_lo);
      tc_fence_after();
      const uint32_t bl = umma_desc_lo(ring_base + st_lo * stage_bytes, b_lbo);
      if (elect_one()) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16_ts(d_tmem, ta_hi + ks * 32 + j * 8, umma_desc_from(bh + j * b_step), idesc, j ? 1u : acc);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16_ts(d_tmem, ta_lo + ks * 32 + j * 8, umma_desc_from(bh + j * b_step), idesc, 1u);
        umma_commit(empty0 + 8u * st_hi);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16_ts(d_tmem, ta_hi + ks * 32 + j * 8, umma_desc_from(bl + j * b_step), idesc, 1u);
        umma_commit(empty0 + 8u * st_lo);
      }
    } else {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
          umma_bf16_ts(d_tmem, ta_hi + ks * 32 + j * 8, umma_desc_from(bh + j * b_step), idesc, j ? 1u : acc);
        umma_commit(empty0 + 8u * st_hi);
      }
    }
    __syncwarp();
    acc = 1u;
  }
}

struct NoShift {
  LFT_DEVINL int operator()(uint32_t) const { return 0; }
};

// LayerNorm statistics helpers operate on register chunks; see kernels.
// ---------------------------------------------------------------- packed fp32x2 math (Blackwell FFMA2)
typedef unsigned long long f32x2;
LFT_DEVINL f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
LFT_DEVINL void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
LFT_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
LFT_DEVINL f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
LFT_DEVINL f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
LFT_DEVINL float hsum2(f32x2 v) {
  float a, b;
  unpack2(v, a, b);
  return a + b;
}

// 16-byte global store of data this kernel never reads back (Q/K/V/tok/O planes).  Cache policy, measured A/B with
// tools/gpu_ab.py (profiles/r01_ab_cache_hints.md): default .L1::no_allocate (k_spa_embed_qkv -3.7 ... -6 %);
// -DLFT_EXPERIMENT_ST=0 plain st.global; =1 st.global.cs (evict-first, -3.1 %); =3 .wt (no change).  The same hint on the
// other kernels' outputs, and ld.global.nc.L1::no_allocate on their read-once inputs, changed nothing (+-0.5 %).
#ifndef LFT_EXPERIMENT_ST
#define LFT_EXPERIMENT_ST 2
#endif
LFT_DEVINL void st_stream_v4(float* p, const float4& v) {
#if LFT_EXPERIMENT_ST == 0
  *reinterpret_cast<float4*>(p) = v;
#elif LFT_EXPERIMENT_ST == 1
  __stcs(reinterpret_cast<float4*>(p), v);
#elif LFT_EXPERIMENT_ST == 2
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
  __stwt(reinterpret_cast<float4*>(p), v);
#endif
}

LFT_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
LFT_DEVINL float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
LFT_DEVINL float lrelu02(float x) { return fmaxf(x, 0.2f * x); }  // == x >= 0 ? x : 0.2x

}  // namespace lft
