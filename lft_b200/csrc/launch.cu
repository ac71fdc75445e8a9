// Launch sequences: which kernels run, in what order, on which buffers (all on the caller's stream).
#include "host.h"

namespace lft {

const char* const kKindNames[K_COUNT] = {"conv0",    "conv3x3_64", "conv3x3_128", "ang_fused", "spa_qkv",  "spa_attn",
                                         "spa_ffn",  "up_gemm",    "up_gather",   "lf_divide", "lf_integrate"};

int configure_kernels() {
  static bool done = false;
  if (done) return 0;
  int rc;
  if ((rc = configure_conv())) return rc;
  if ((rc = configure_ang())) return rc;
  if ((rc = configure_spa())) return rc;
  if ((rc = configure_up())) return rc;
  done = true;
  return 0;
}

size_t ws_floats_per_token(int scale) { return 4 * 64 + 5 * 128 + 9 * (size_t)scale * scale; }

Workspace carve(void* ws, long long T, int scale) {
  Workspace w;
  float* p = reinterpret_cast<float*>(ws);
  w.f0 = p; p += T * 64;
  w.f1 = p; p += T * 64;
  w.f2 = p; p += T * 64;
  w.fres = p; p += T * 64;
  w.tok = p; p += T * 128;
  w.q = p; p += T * 128;
  w.k = p; p += T * 128;
  w.v = p; p += T * 128;
  w.o = p; p += T * 128;
  w.pp = p;
  return w;
}

// conv_init0 -> 3 x (conv + LeakyReLU) -> + conv_init0 output   (LFT.py:65-66)
int run_conv_init(Handle* h, const float* lr, float* out, float* tmp0, float* tmp1, float* tmp2, int B, int P,
                  cudaStream_t st) {
  const int A = h->cfg.ang_res, V = B * A * A;
  const long long T = (long long)V * P * P;
  (void)T;
  int rc;
  if ((rc = launch_conv0(h, lr, tmp0, B, P, st))) return rc;
  if ((rc = launch_conv3x3(h, 64, tmp0, h->w_conv[0], tmp1, nullptr, V, P, 1, st))) return rc;
  if ((rc = launch_conv3x3(h, 64, tmp1, h->w_conv[1], tmp2, nullptr, V, P, 1, st))) return rc;
  if ((rc = launch_conv3x3(h, 64, tmp2, h->w_conv[2], out, tmp0, V, P, 3, st))) return rc;
  return 0;
}

}  // namespace lft

using namespace lft;

static int check_ready(Handle* h, int B, int P) {
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  if (!h->finalized) return fail(LFT_ERR_STATE, "weights not finalized (call lft_finalize_weights)");
  if (B < 1) return fail(LFT_ERR_ARG, "B must be >= 1");
  if (P < 4 || P > 32) return fail(LFT_ERR_ARG, "patch size P=%d unsupported (4..32, square patches only)", P);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  return ensure_spa_pe(h, P);
}

extern "C" {

int lft_workspace_bytes(lft_handle* hh, int32_t B, int32_t P, size_t* bytes) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !bytes || B < 1 || P < 1) return fail(LFT_ERR_ARG, "bad argument");
  const size_t T = (size_t)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  *bytes = T * ws_floats_per_token(h->cfg.scale) * sizeof(float) + 1024;
  return 0;
}

int lft_stage_conv_init(lft_handle* hh, const float* lr, float* feat, int32_t B, int32_t P, void* ws, size_t ws_bytes,
                        void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  return run_conv_init(h, lr, feat, w.f0, w.f1, w.f2, B, P, (cudaStream_t)stream);
}

int lft_stage_ang(lft_handle* hh, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if (layer < 0 || layer >= kLayers) return fail(LFT_ERR_ARG, "layer out of range");
  (void)ws; (void)ws_bytes;
  return run_ang(h, layer, in, out, B, P, (cudaStream_t)stream);
}

int lft_stage_spa(lft_handle* hh, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if (layer < 0 || layer >= kLayers) return fail(LFT_ERR_ARG, "layer out of range");
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  return run_spa(h, layer, in, out, nullptr, w, B, P, (cudaStream_t)stream);
}

int lft_gemm_selftest(const float* A, const float* W, float* D, float* aux, int32_t M, int32_t N, int32_t K,
                      int32_t precision, int32_t variant) {
  if (!A || !W || !D || !aux || M % 128 || K % 64 || N % 16 || N > 256 || N < 16 || K > 256)
    return fail(LFT_ERR_ARG, "bad selftest shape");
  int rc = configure_kernels();
  if (rc) return rc;
  std::vector<uint16_t> p = pack_weight(N, N, K, [=](int n, int k) { return W[(size_t)n * K + k]; });
  float *dA = nullptr, *dD = nullptr, *dX = nullptr;
  uint8_t* dW = nullptr;
  CUDA_TRY(cudaMalloc(&dA, (size_t)M * K * 4));
  CUDA_TRY(cudaMalloc(&dD, (size_t)M * N * 4));
  CUDA_TRY(cudaMalloc(&dX, (size_t)M * 16 * 4));
  CUDA_TRY(cudaMalloc(&dW, p.size() * 2));
  CUDA_TRY(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dW, p.data(), p.size() * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(dD, 0, (size_t)M * N * 4));
  if ((rc = launch_selftest(dA, K, dW, N, dD, dX, M, precision == LFT_PREC_FP32 ? 3 : 1, variant))) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(D, dD, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(aux, dX, (size_t)M * 16 * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dD); cudaFree(dX); cudaFree(dW);
  return 0;
}

}  // extern "C"
