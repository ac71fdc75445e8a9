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

enum Kind {
  K_CONV0 = 0,
  K_CONV64,
  K_CONV128,
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
  std::vector<float> a_tab;       // LN-folded epilogue constants [u_qk 128 | c_qk 128 | u_1 128 | c_1 128] (host; kernel param)
  const float* a_peqk = nullptr;  // [A*A][128]  PE_a W'qk^T
  // SpaTrans (LFT.py:118-191)
  const uint8_t *s_wmlp = nullptr, *s_wq = nullptr, *s_wk = nullptr, *s_wv = nullptr, *s_wo = nullptr;
  const uint8_t *s_w1a = nullptr, *s_w1b = nullptr, *s_w2a = nullptr, *s_w2b = nullptr, *s_wlin = nullptr;
  const float* s_ln = nullptr;  // [norm.w | norm.b | ff0.w | ff0.b] x 128
  const float* s_pe = nullptr;    // [P*P][128] SAI2Token(spa_position), rebuilt when P changes
  const float* s_pev = nullptr;   // chunk-planar [32][P*P][4]: PE_s Wv^T, rebuilt when P changes
  std::vector<float> s_tab;       // [u_q 128 | u_k 128 | c_q 128 | c_k 128 | u_1 256 | c_1 256] (host; kernel params)

};

struct ProfEvent {
  cudaEvent_t start, stop;
  int kind;
};

struct Handle {
  lft_config cfg{};
  std::map<std::string, std::vector<int64_t>> spec;
  std::map<std::string, std::vector<float>> host_w;
  bool finalized = false;
  std::vector<void*> allocs;
  std::vector<float> w_conv0_host;  // conv_init0 weight [64][9] (kernel parameter of the fused first/last conv)
  const uint8_t* w_conv[3] = {nullptr, nullptr, nullptr};
  const uint8_t* w_conv_st[3] = {nullptr, nullptr, nullptr};  // fp32 mode: hi and lo rows stacked to one [128 x 64] slab per tap
  Layer layer[kLayers];
  const uint8_t* w_up = nullptr;
  const uint8_t* w_up3 = nullptr;  // upsampling.3.weight packed as a [16 x 64] bf16 hi/lo B operand (rows 9..15 zero)
  const float* pe_ang = nullptr;
  int pe_P = -1;
  bool profiling = false;
  std::vector<ProfEvent> events;
  int64_t launches = 0;
  int num_sms = 148;
  int passes() const { return cfg.precision == LFT_PREC_FP32 ? 3 : 1; }
};

int upload(Handle* h, const void* src, size_t bytes, void** dst);
int ensure_spa_pe(Handle* h, int P);
std::vector<uint16_t> pack_weight(int N, int Npad, int K, const std::function<float(int, int)>& w);

struct Scope {  // profiling + launch accounting around one kernel launch
  Handle* h;
  cudaStream_t st;
  ProfEvent ev{};
  bool on;
  Scope(Handle* h_, int kind, cudaStream_t st_) : h(h_), st(st_), on(h_->profiling) {
    if (on) {
      cudaEventCreate(&ev.start);
      cudaEventCreate(&ev.stop);
      ev.kind = kind;
      cudaEventRecord(ev.start, st);
    }
  }
  int finish() {
    h->launches++;
    if (on) {
      cudaEventRecord(ev.stop, st);
      h->events.push_back(ev);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LFT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
};

// launch.cu
int configure_kernels();
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
int run_ang(Handle* h, int layer, const float* in, float* out, int B, int P, cudaStream_t st);
int run_spa(Handle* h, int layer, const float* in, float* out, const float* final_res, Workspace& w, int B, int P,
            cudaStream_t st);
int launch_divide(Handle* h, const float* lf, float* patches, int h0, int w0, int numV, int p0, int n, int P, int S,
                  cudaStream_t st);
int launch_integrate(Handle* h, const float* crops, float* sr, int h0, int w0, int numV, int p0, int n, int S,
                     cudaStream_t st);
int run_upsample(Handle* h, const float* feat, const float* lr, float* sr, float* pp, int B, int P, int crop_stride,
                 cudaStream_t st);

}  // namespace lft
