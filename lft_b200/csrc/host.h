// Host-side internals shared by api.cu and launch.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/lft_b200.h"

namespace lft {

constexpr int kLayers = 4;

int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return ::lft::fail(LFT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// A square pixel region [r0, r0 + rn)^2 of every P x P view.  The light-field path (lft_forward_lf*) keeps only the
// central crop of every SR patch view, so the last layers are evaluated on the shrinking regions that crop depends on
// (launch.cu: crop_regions); every other caller passes the full view {0, P}.
struct Region {
  int r0, rn;
};

enum Kind {
  K_CONV64 = 0,
  K_ANG,
  K_SPA_QKV,
  K_SPA_ATTN,
  K_SPA_FFN,
  K_UP_GEMM,
  K_UP_GATHER,
  K_DIVIDE,
  K_INTEGRATE,
  K_COUNT
};
extern const char* const kKindNames[K_COUNT];

struct Layer {
  // AngTrans (LFT.py:194-238)
  const uint8_t *a_wqk = nullptr, *a_wv = nullptr, *a_wo = nullptr, *a_w1 = nullptr, *a_w2 = nullptr;
  const float* a_ln = nullptr;    // [norm.w | norm.b | ff0.w | ff0.b] x 64
  // Constants that must agree with the weights the MMAs really multiply by exist once per precision mode
  // ([0]: fp32 mode, bf16 hi + lo weights; [1]: bf16 mode, hi weights only): the LN fold subtracts mean * u with
  // u = sum_k W'[n,k], which cancels the row mean only if it sums exactly the values the tensor core used.
  std::vector<float> a_tab[2];    // LN-folded epilogue constants [u_qk 128 | c_qk 128 | u_1 128 | c_1 128] (host; kernel param)
  const float* a_peqk[2] = {nullptr, nullptr};  // [A*A][128]  PE_a W'qk^T
  // SpaTrans (LFT.py:118-191)
  const uint8_t *s_wmlp = nullptr, *s_wq = nullptr, *s_wk = nullptr, *s_wv = nullptr, *s_wo = nullptr;
  const uint8_t *s_w1a = nullptr, *s_w1b = nullptr, *s_w2a = nullptr, *s_w2b = nullptr, *s_wlin = nullptr;
  const float* s_ln = nullptr;  // [norm.w | norm.b | ff0.w | ff0.b] x 128
  const float* s_pe = nullptr;    // [P*P][128] SAI2Token(spa_position) for the current patch size (points into Handle::pe_cache)
  const float* s_pev[2] = {nullptr, nullptr};   // chunk-planar [32][P*P][4]: PE_s Wv^T, likewise (per precision mode)
  std::vector<float> s_tab[2];    // [u_q 128 | u_k 128 | c_q 128 | c_k 128 | u_1 256 | c_1 256] (host; kernel params)

};

struct ProfEvent {
  cudaEvent_t start, stop;
  int kind;
  long long units;  // what the launch processed: LR tokens (pixels x views) for the network kernels, output pixels for the tilers
};

// spatial position tables of one patch size (per layer), kept for the life of the weights they were built from
struct SpaPe {
  const float* pe[kLayers] = {nullptr, nullptr, nullptr, nullptr};
  const float* pev[kLayers][2] = {};
};

// RAII: make `dev` current for the duration of an API call and restore the caller's device afterwards (the library must not
// change the current device of the torch process that calls it)
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    else if (prev == dev) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Handle {
  lft_config cfg{};
  std::map<std::string, std::vector<int64_t>> spec;
  std::map<std::string, std::vector<float>> host_w;
  bool finalized = false;
  std::vector<void*> allocs;       // device allocations of the current weight generation (freed on re-finalize / destroy)
  std::map<int, SpaPe> pe_cache;   // patch size -> spatial position tables (same life time as `allocs`)
  std::vector<float> w_conv0_host;  // conv_init0 weight [64][9] (kernel parameter of the fused first/last conv)
  const uint8_t* w_conv[3] = {nullptr, nullptr, nullptr};
  const uint8_t* w_conv_st[3] = {nullptr, nullptr, nullptr};  // fp32 mode: hi and lo rows stacked to one [128 x 64] slab per tap
  Layer layer[kLayers];
  const uint8_t* w_up = nullptr;
  const uint8_t* w_up3 = nullptr;  // upsampling.3.weight packed as a [16 x 64] bf16 hi/lo B operand (rows 9..15 zero)
  const float* pe_ang = nullptr;
  int pe_P = -1;
  // the light-field path runs the two halves of its patch range on two streams (the caller's and `side`): the kernels are
  // persistent grids of all 296 CTA slots, so at small batches (8 patches per GPU when a light field is sharded over 8 GPUs = 5.5
  // waves per kernel) the last, partly filled wave of one half's kernel is filled by the other half's: -6.7 % at 8 patches, -2.7 %
  // at 16, -1.3 % at 64 (tools/gpu_two_stream_lf.py).  LFT_STREAMS=1 disables.
  bool two_streams = true;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pdl = false;                // programmatic dependent launch between the kernels of a forward (LFT_PDL=1 enables)
  bool profiling = false;
  std::vector<ProfEvent> events;   // event pool: created once, reused by every profiling session
  size_t n_events = 0;             // events of the current session
  int64_t launches = 0;
  int num_sms = 148;
  int passes() const { return cfg.precision == LFT_PREC_FP32 ? 3 : 1; }
  int mode() const { return cfg.precision == LFT_PREC_FP32 ? 0 : 1; }  // index of the per-precision constant tables
};

int upload(Handle* h, const void* src, size_t bytes, void** dst);
int ensure_spa_pe(Handle* h, int P);
std::vector<uint16_t> pack_weight(int N, int Npad, int K, const std::function<float(int, int)>& w);

// Kernel launch, optionally with the programmatic-stream-serialization attribute (see pdl_trigger / pdl_wait in common.cuh).
// Measured on the benchmark (profiles/r02_pdl_ab.md): all parity / repeatability tests pass with it, but the step gets SLOWER -
// 2.310 vs 2.273 ms at N = 8 (two runs each, identical to 1e-3), no difference beyond noise at N = 1: the kernels are persistent
// grids that fill every CTA slot, so a dependent's CTAs only get on an SM when one of the running kernel's CTAs retires early, and
// then hold TMEM / shared memory and poll while that SM's other CTA finishes; launch gaps were not the cost they looked like
// (the per-launch profiling events were: 2.40 vs 2.27 ms per step at N = 8 with / without them).  Hence off by default.
// The attribute is dropped while per-launch profiling events are recorded: an event between two kernels serialises them anyway.
#define LFT_LAUNCH(h, kernel, grid, block, smem, st, ...)                                             \
  do {                                                                                                \
    cudaLaunchConfig_t cfg_ = {};                                                                     \
    cfg_.gridDim = dim3(grid);                                                                        \
    cfg_.blockDim = dim3(block);                                                                      \
    cfg_.dynamicSmemBytes = (smem);                                                                   \
    cfg_.stream = (st);                                                                               \
    cudaLaunchAttribute at_[1];                                                                       \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                   \
    at_[0].val.programmaticStreamSerializationAllowed = 1;                                            \
    cfg_.attrs = at_;                                                                                 \
    cfg_.numAttrs = ((h)->pdl && !(h)->profiling) ? 1 : 0;                                            \
    cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                                   \
  } while (0)

struct Scope {  // profiling + launch accounting around one kernel launch
  Handle* h;
  cudaStream_t st;
  ProfEvent ev{};
  bool on;
  Scope(Handle* h_, int kind, cudaStream_t st_, long long units = 0) : h(h_), st(st_), on(h_->profiling) {
    if (on) {
      if (h->n_events == h->events.size()) {  // grow the pool (only the first session of a given length pays for this)
        ProfEvent e{};
        if (cudaEventCreate(&e.start) != cudaSuccess || cudaEventCreate(&e.stop) != cudaSuccess) { on = false; return; }
        h->events.push_back(e);
      }
      ev = h->events[h->n_events];
      ev.kind = kind;
      ev.units = units;
      cudaEventRecord(ev.start, st);
    }
  }
  int finish() {
    h->launches++;
    if (on) {
      cudaEventRecord(ev.stop, st);
      h->events[h->n_events++] = ev;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LFT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
};

// launch.cu
int configure_kernels(int device);
// per-file kernel configuration (max dynamic smem opt-in) and launchers
int configure_conv();
int configure_ang();
int configure_spa();
int configure_up();
int debug_timeline_spa(long long* out);
int debug_timeline_ang(long long* out);
int debug_timeline_embed(long long* out);
int debug_timeline_ring_embed(long long* out);
int launch_conv3x3(Handle* h, int N, const float* in, const uint8_t* wp, const uint8_t* wst, float* out, const float* res,
                   int V, int P, int epi, const float* lr, cudaStream_t st);
int launch_layout(Handle* h, const float* in, float* out, long long T, int C, int to_t32, cudaStream_t st);
int launch_mma_bench(int N, int K, int reps, int mode, int grid, int smem_bytes, long long* host_out);
int launch_selftest(const float* dA, int K, const uint8_t* dW, int N, float* dD, float* dX, int M, int passes,
                    int variant);

// Workspace carve-up for a chunk of Bc patches (floats): see launch.cu
struct Workspace {
  float *f0, *f1, *f2, *fres, *tok, *q, *k, *v, *o, *pp, *lrp;
};
size_t ws_floats_per_token(int scale);
Workspace carve(void* ws, long long T, int scale);

int run_conv_init(Handle* h, const float* lr, float* out, float* tmp0, float* tmp1, float* tmp2, int B, int P,
                  cudaStream_t st);
int run_ang(Handle* h, int layer, const float* in, float* out, int B, int P, Region rg, cudaStream_t st);
int run_spa(Handle* h, int layer, const float* in, float* out, const float* final_res, Workspace& w, int B, int P,
            Region need, cudaStream_t st);
int launch_divide(Handle* h, const float* lf, float* patches, int h0, int w0, int numV, int p0, int n, int P, int S,
                  cudaStream_t st);
int launch_integrate(Handle* h, const float* crops, float* sr, int h0, int w0, int numV, int p0, int n, int S,
                     cudaStream_t st);
// where the up-sampling tail writes: whole SR patches, the crops LFintegrate keeps, or those crops at their final position in
// the assembled SR light field (which may be a peer-mapped buffer of another GPU)
struct UpTarget {
  int mode = 0;           // 0: sr [B,1,H,H]   1: crops [B][A][A][cs][cs]   2: sr_lf [A*h0*s, A*w0*s] (LFintegrate fused)
  int crop_stride = 0;    // LR stride S of the tiling (modes 1, 2)
  int h0 = 0, w0 = 0, numV = 0, p0 = 0;  // mode 2: light-field geometry and the index of the first patch of this chunk
};
Region up_region(int P, int s, const UpTarget& t);  // LR pixels of every view the kept SR pixels depend on
int run_upsample(Handle* h, const float* feat, const float* lr, float* out, float* pp, int B, int P, const UpTarget& t,
                 cudaStream_t st);

}  // namespace lft
