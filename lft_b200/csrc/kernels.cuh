// Kernel-side shared declarations: CTA roles, control block, setup/teardown, kernel prototypes.
#pragma once
#include "common.cuh"

namespace lft {

// CTA roles for the tcgen05 kernels: warps 0-3 own one accumulator row each (TMEM lane = thread id),
// warp 4 streams weights (one elected lane), warp 5 allocates TMEM and issues the MMAs (one lane).
constexpr int kThreads = 192;
constexpr int kWarpProducer = 4;
constexpr int kWarpMma = 5;
constexpr int kCtlBytes = 256;
// conv staging: rows <-> positions g0-kConvOff .. g0-kConvOff+kConvRows-1 (P <= 32 -> |shift| <= 34)
constexpr int kConvRows = 201;  // odd: k-chunk planes start in different banks
constexpr int kConvOff = 36;

struct Ctl {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t a_ready;
  uint64_t mma_done;
  uint64_t aux[4];
  uint32_t tmem;
};
static_assert(sizeof(Ctl) <= kCtlBytes, "control block too large");

template <int NST>
LFT_DEVINL void cta_setup(Ctl* ctl, int warp, int lane, uint32_t a_ready_count, uint32_t tmem_cols) {
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
  if (warp == kWarpMma) tmem_alloc(smem_u32(&ctl->tmem), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

LFT_DEVINL void cta_teardown(Ctl* ctl, int warp, uint32_t tmem_cols) {
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(ctl->tmem, tmem_cols);
  }
}

}  // namespace lft
