// Launch sequences: which kernels run, in what order, on which buffers (all on the caller's stream).
#include "host.h"

#include <cstring>
#include <mutex>
#include <set>

namespace lft {

const char* const kKindNames[K_COUNT] = {"conv3x3_64", "ang_fused", "spa_embed_qkv", "spa_attn", "spa_ffn",
                                         "up_gemm",    "up_gather", "lf_divide",     "lf_integrate"};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: configure every device a handle is created on once
// (the caller has made `device` current).
int configure_kernels(int device) {
  static std::mutex mu;
  static std::set<int> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(device)) return 0;
  int rc;
  if ((rc = configure_conv())) return rc;
  if ((rc = configure_ang())) return rc;
  if ((rc = configure_spa())) return rc;
  if ((rc = configure_up())) return rc;
  done.insert(device);
  return 0;
}

size_t ws_floats_per_token(int scale) { return 4 * 64 + 5 * 128 + 9 * (size_t)scale * scale + 1; }

Workspace carve(void* ws, long long T, int scale) {
  T = (T + 127) / 128 * 128;  // T32 layout: whole 32-token blocks; the O operand tiles of k_spa_ffn2: whole 128-token tiles
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
  w.pp = p; p += T * 9 * scale * scale;
  w.lrp = p;
  return w;
}

// conv_init0 -> 3 x (conv + LeakyReLU) -> + conv_init0 output   (LFT.py:65-66)
int run_conv_init(Handle* h, const float* lr, float* out, float* tmp0, float* tmp1, float* tmp2, int B, int P,
                  cudaStream_t st) {
  const int A = h->cfg.ang_res, V = B * A * A;
  const long long T = (long long)V * P * P;
  (void)T;
  int rc;
  (void)tmp0;  // conv_init0 is fused into the first conv's loader and the last conv's residual (never materialised)
  if ((rc = launch_conv3x3(h, 64, nullptr, h->w_conv[0], h->w_conv_st[0], tmp1, nullptr, V, P, 1, lr, st))) return rc;
  if ((rc = launch_conv3x3(h, 64, tmp1, h->w_conv[1], h->w_conv_st[1], tmp2, nullptr, V, P, 1, lr, st))) return rc;
  if ((rc = launch_conv3x3(h, 64, tmp2, h->w_conv[2], h->w_conv_st[2], out, nullptr, V, P, 3, lr, st))) return rc;
  return 0;
}

static Region grow3(Region r, int P) {  // receptive field of one AltFilter: 5x5 window (2) + 3x3 token embedding (1)
  const int lo = r.r0 - 3 < 0 ? 0 : r.r0 - 3;
  const int hi = r.r0 + r.rn + 3 > P ? P : r.r0 + r.rn + 3;
  return Region{lo, hi - lo};
}

// get_model.forward for one chunk of B patches (LFT.py:52-83); `up` says where the result goes (whole SR patches, the crops
// LFintegrate keeps, or those crops at their place in the assembled light field).
// Dead-work elimination on the light-field path: LFintegrate keeps the central stride*s block of every SR patch view, which
// depends on the LR pixels up_region() of the last feature map; each AltFilter widens that by 3 pixels (AngTrans is
// pixel-wise, SpaTrans = 5x5 window + 3x3 embedding).  Layer i therefore runs on need[i] only - for the default 32 / 16
// tiling 18^2, 24^2, 30^2, 32^2 of the 32^2 pixels from the last layer backwards.  Every kernel computes a token exactly as
// it does on the full view, so the kept pixels are bit-identical to the crop of the full forward.
int run_forward_chunk(Handle* h, const float* lr, float* out, Workspace& w, int B, int P, const UpTarget& up,
                      cudaStream_t st) {
  int rc;
  if ((rc = run_conv_init(h, lr, w.fres, w.f0, w.f1, w.f2, B, P, st))) return rc;
  Region need[kLayers];
  need[kLayers - 1] = up_region(P, h->cfg.scale, up);
  for (int i = kLayers - 2; i >= 0; --i) need[i] = grow3(need[i + 1], P);
  const float* x = w.fres;
  for (int i = 0; i < kLayers; ++i) {  // AltFilter: ang_trans then spa_trans (LFT.py:248-252)
    if ((rc = run_ang(h, i, x, w.f1, B, P, grow3(need[i], P), st))) return rc;
    if ((rc = run_spa(h, i, w.f1, w.f2, i == kLayers - 1 ? w.fres : nullptr, w, B, P, need[i], st))) return rc;
    x = w.f2;
  }
  return run_upsample(h, w.f2, lr, out, w.pp, B, P, up, st);
}

// numU / numV of LFdivide (utils.py:93-104) for patch size P and stride S (test.py:83 passes args.patch_size_for_test /
// args.stride_for_test, defaults 32 / 16).  h - P may be negative (view smaller than a patch: Python's floor division then
// gives one zero-padded patch as long as h - P > -S; below that the reference returns an empty tiling -> rejected here).
static int num_patches_1d(int n0, int P, int S, int bdr) {
  const int n = n0 + 2 * bdr;
  if (n >= P) return (n - P) / S + (((n - P) % S) ? 2 : 1);
  return (n - P > -S) ? 1 : 0;
}

static int tiling(int h0, int w0, int P, int S, int* numU, int* numV) {
  if (P < 4 || P > 64) return fail(LFT_ERR_ARG, "patch size %d unsupported (4..64)", P);
  if (S < 1 || S > P) return fail(LFT_ERR_ARG, "stride %d outside [1, patch size %d]", S, P);
  const int bdr = (P - S) / 2;
  if (h0 < 1 || w0 < 1 || h0 < bdr || w0 < bdr)
    return fail(LFT_ERR_ARG, "light field %dx%d per view is smaller than the mirror border %d", h0, w0, bdr);
  const int nu = num_patches_1d(h0, P, S, bdr), nv = num_patches_1d(w0, P, S, bdr);
  if (nu < 1 || nv < 1) return fail(LFT_ERR_ARG, "light field %dx%d per view yields no patch of size %d at stride %d", h0, w0, P, S);
  // an odd (patch - stride) can leave numU*stride == h0 - 1: the reference's LFintegrate then fails with a shape mismatch
  // (utils.py:155 assigns a [numU*stride, ...] block to [h0, ...]); rejected instead of returning an unwritten last row
  if (nu * S < h0 || nv * S < w0)
    return fail(LFT_ERR_ARG, "patch %d / stride %d does not cover a %dx%d view (%d x %d kept pixels): the reference's LFintegrate fails here too", P, S, h0, w0, nu * S, nv * S);
  *numU = nu;
  *numV = nv;
  return 0;
}

}  // namespace lft

using namespace lft;

// Run fn(begin, end, workspace part, its capacity in patches, stream) over the units [0, n): as one range on the caller's
// stream, or - Handle::two_streams - as two halves on the caller's stream and the handle's side stream, forked and joined by
// events, each with its share of the workspace (`per` bytes per patch).  Not while per-launch profiling events are recorded: a
// kernel queued behind the other stream's grid would be timed with its wait.
template <typename F>
static int run_split(Handle* h, long long n, void* ws, size_t ws_bytes, size_t per, cudaStream_t st, F&& fn) {
  const long long cap = (long long)((ws_bytes - 1024) / per);
  if (h->two_streams && !h->profiling && n >= 2 && cap >= 2) {
    const long long nA = (n + 1) / 2;
    long long capA = cap / 2 + (cap & 1), capB = cap - capA;
    if (capA > nA) { capA = nA; capB = cap - capA; }
    char* wsB = reinterpret_cast<char*>(ws) + (size_t)capA * per;
    wsB = reinterpret_cast<char*>(((uintptr_t)wsB + 1023) & ~(uintptr_t)1023);
    if ((size_t)(wsB - reinterpret_cast<char*>(ws)) + (size_t)capB * per > ws_bytes) capB -= 1;
    if (capB >= 1) {
      int rc;
      CUDA_TRY(cudaEventRecord(h->ev_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
      if ((rc = fn(0LL, nA, ws, capA, st))) return rc;
      if ((rc = fn(nA, n, (void*)wsB, capB, h->side))) return rc;
      CUDA_TRY(cudaEventRecord(h->ev_join, h->side));
      CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join, 0));
      return 0;
    }
  }
  return fn(0LL, n, ws, cap, st);
}

static int check_ready(Handle* h, int B, int P) {
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  if (!h->finalized) return fail(LFT_ERR_STATE, "weights not finalized (call lft_finalize_weights)");
  if (B < 1) return fail(LFT_ERR_ARG, "B must be >= 1");
  if (P < 4 || P > 64) return fail(LFT_ERR_ARG, "patch size P=%d unsupported (4..64, square patches only)", P);
  return ensure_spa_pe(h, P);
}

// the stage entry points do not chunk: the whole batch must fit the kernels' 32-bit index spaces
static int check_stage_size(Handle* h, int B, int P) {
  const long long A = h->cfg.ang_res, s = h->cfg.scale;
  if ((long long)B * A * P * s * A * P * s >= (1LL << 30))
    return fail(LFT_ERR_ARG, "stage entry points support B*(A*P*s)^2 < 2^30; use lft_forward for larger batches");
  return 0;
}

extern "C" {

int lft_workspace_bytes(lft_handle* hh, int32_t B, int32_t P, size_t* bytes) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !bytes || B < 1 || P < 1) return fail(LFT_ERR_ARG, "bad argument");
  size_t T = (size_t)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  T = (T + 127) / 128 * 128;
  *bytes = T * ws_floats_per_token(h->cfg.scale) * sizeof(float) + 1024;
  return 0;
}

int lft_stage_conv_init(lft_handle* hh, const float* lr, float* feat, int32_t B, int32_t P, void* ws, size_t ws_bytes,
                        void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if ((rc = check_stage_size(h, B, P))) return rc;
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  if ((rc = run_conv_init(h, lr, w.fres, w.f0, w.f1, w.f2, B, P, (cudaStream_t)stream))) return rc;
  return launch_layout(h, w.fres, feat, T, 64, 0, (cudaStream_t)stream);
}

int lft_stage_ang(lft_handle* hh, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if ((rc = check_stage_size(h, B, P))) return rc;
  if (layer < 0 || layer >= kLayers) return fail(LFT_ERR_ARG, "layer out of range");
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  if ((rc = launch_layout(h, in, w.f0, T, 64, 1, (cudaStream_t)stream))) return rc;
  if ((rc = run_ang(h, layer, w.f0, w.f1, B, P, Region{0, P}, (cudaStream_t)stream))) return rc;
  return launch_layout(h, w.f1, out, T, 64, 0, (cudaStream_t)stream);
}

int lft_stage_spa(lft_handle* hh, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if ((rc = check_stage_size(h, B, P))) return rc;
  if (layer < 0 || layer >= kLayers) return fail(LFT_ERR_ARG, "layer out of range");
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  if ((rc = launch_layout(h, in, w.f0, T, 64, 1, (cudaStream_t)stream))) return rc;
  if ((rc = run_spa(h, layer, w.f0, w.f1, nullptr, w, B, P, Region{0, P}, (cudaStream_t)stream))) return rc;
  return launch_layout(h, w.f1, out, T, 64, 0, (cudaStream_t)stream);
}

int lft_stage_upsample(lft_handle* hh, const float* feat, const float* lr, float* sr, int32_t B, int32_t P, void* ws,
                       size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if ((rc = check_stage_size(h, B, P))) return rc;
  size_t need;
  lft_workspace_bytes(hh, B, P, &need);
  if (ws_bytes < need) return fail(LFT_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  const long long T = (long long)B * h->cfg.ang_res * h->cfg.ang_res * P * P;
  Workspace w = carve(ws, T, h->cfg.scale);
  if ((rc = launch_layout(h, feat, w.f0, T, 64, 1, (cudaStream_t)stream))) return rc;
  return run_upsample(h, w.f0, lr, sr, w.pp, B, P, UpTarget{}, (cudaStream_t)stream);
}

int lft_forward(lft_handle* hh, const float* lr, float* sr, int32_t B, int32_t P, void* ws, size_t ws_bytes,
                void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  int rc = check_ready(h, B, P);
  if (rc) return rc;
  if (!lr || !sr || !ws) return fail(LFT_ERR_ARG, "null pointer");
  size_t per;
  lft_workspace_bytes(hh, 1, P, &per);
  per -= 1024;
  if (ws_bytes < per + 1024) return fail(LFT_ERR_WORKSPACE, "workspace too small for one patch: %zu < %zu", ws_bytes, per + 1024);
  const int A = h->cfg.ang_res, s = h->cfg.scale;
  // kernels use 32-bit index math: keep B*(A*P*s)^2 (largest per-chunk index space) below 2^30
  const long long cap32 = (1LL << 30) / ((long long)A * P * s * A * P * s);
  const size_t lr_stride = (size_t)A * P * A * P, sr_stride = lr_stride * s * s;
  return run_split(h, B, ws, ws_bytes, per, (cudaStream_t)stream,
                   [&](long long b_begin, long long b_end, void* wsp, long long cap, cudaStream_t st) -> int {
                     const long long chunk = cap > cap32 ? (cap32 < 1 ? 1 : cap32) : cap;
                     for (long long b0 = b_begin; b0 < b_end; b0 += chunk) {
                       const int Bc = (int)((b_end - b0) < chunk ? (b_end - b0) : chunk);
                       const long long T = (long long)Bc * A * A * P * P;
                       Workspace w = carve(wsp, T, s);
                       int r = run_forward_chunk(h, lr + b0 * lr_stride, sr + b0 * sr_stride, w, Bc, P, UpTarget{}, st);
                       if (r) return r;
                     }
                     return 0;
                   });
}

int lft_lf_num_patches_ex(int32_t h0, int32_t w0, int32_t patch, int32_t stride, int32_t* numU, int32_t* numV) {
  if (!numU || !numV) return fail(LFT_ERR_ARG, "null argument");
  int a, b;
  int rc = tiling(h0, w0, patch, stride, &a, &b);
  if (rc) return rc;
  *numU = a;
  *numV = b;
  return 0;
}

int lft_lf_num_patches(int32_t h0, int32_t w0, int32_t* numU, int32_t* numV) {
  return lft_lf_num_patches_ex(h0, w0, 32, 16, numU, numV);
}

int lft_divide_ex(lft_handle* hh, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride, int32_t p0,
                  int32_t p1, float* patches, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !lr_lf || !patches) return fail(LFT_ERR_ARG, "bad argument");
  int nu, nv;
  int rc = tiling(h0, w0, patch, stride, &nu, &nv);
  if (rc) return rc;
  if (p0 < 0 || p1 > nu * nv || p0 > p1) return fail(LFT_ERR_ARG, "patch range [%d,%d) outside [0,%d)", p0, p1, nu * nv);
  if (p0 == p1) return 0;
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  return launch_divide(h, lr_lf, patches, h0, w0, nv, p0, p1 - p0, patch, stride, (cudaStream_t)stream);
}

int lft_divide(lft_handle* hh, const float* lr_lf, int32_t h0, int32_t w0, int32_t p0, int32_t p1, float* patches,
               void* stream) {
  return lft_divide_ex(hh, lr_lf, h0, w0, 32, 16, p0, p1, patches, stream);
}

int lft_integrate_ex(lft_handle* hh, const float* sr_crops, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                     int32_t p0, int32_t p1, float* sr_lf, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !sr_crops || !sr_lf) return fail(LFT_ERR_ARG, "bad argument");
  int nu, nv;
  int rc = tiling(h0, w0, patch, stride, &nu, &nv);
  if (rc) return rc;
  if (p0 < 0 || p1 > nu * nv || p0 > p1) return fail(LFT_ERR_ARG, "patch range [%d,%d) outside [0,%d)", p0, p1, nu * nv);
  if (p0 == p1) return 0;
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  return launch_integrate(h, sr_crops, sr_lf, h0, w0, nv, p0, p1 - p0, stride, (cudaStream_t)stream);
}

int lft_integrate(lft_handle* hh, const float* sr_crops, int32_t h0, int32_t w0, int32_t p0, int32_t p1, float* sr_lf,
                  void* stream) {
  return lft_integrate_ex(hh, sr_crops, h0, w0, 32, 16, p0, p1, sr_lf, stream);
}

// LFdivide -> forward -> (crops | crops at their place in sr_lf) for the patch range [p0, p1), chunked to the workspace
static int forward_lf_impl(lft_handle* hh, const float* lr_lf, int h0, int w0, int patch, int stride, int p0, int p1,
                           float* dst, bool direct, void* ws, size_t ws_bytes, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  if (!lr_lf || !dst || !ws) return fail(LFT_ERR_ARG, "bad argument");
  int nu, nv;
  int rc = tiling(h0, w0, patch, stride, &nu, &nv);
  if (rc) return rc;
  const int P = patch;
  if ((rc = check_ready(h, 1, P))) return rc;
  if (p0 < 0 || p1 > nu * nv || p0 > p1) return fail(LFT_ERR_ARG, "patch range [%d,%d) outside [0,%d)", p0, p1, nu * nv);
  size_t per;
  lft_workspace_bytes(hh, 1, P, &per);
  per -= 1024;
  if (ws_bytes < per + 1024) return fail(LFT_ERR_WORKSPACE, "workspace too small for one patch: %zu < %zu", ws_bytes, per + 1024);
  const int A = h->cfg.ang_res, s = h->cfg.scale;
  if (direct && (long long)A * h0 * s * A * w0 * s >= (1LL << 31))
    return fail(LFT_ERR_ARG, "assembled light field too large for 32-bit pixel indices");
  const long long cap32 = (1LL << 30) / ((long long)A * P * s * A * P * s);  // 32-bit index spaces per chunk
  const size_t crop_stride = (size_t)A * A * stride * s * stride * s;
  // patches [q_begin, q_end) on stream `st`, chunked to the workspace part [wsp, wsp + cap patches)
  auto run_range = [&](long long q_begin, long long q_end, void* wsp, long long cap, cudaStream_t st) -> int {
    long long chunk = cap > cap32 ? (cap32 < 1 ? 1 : cap32) : cap;
    for (long long q0 = q_begin; q0 < q_end; q0 += chunk) {
      const int Bc = (int)((q_end - q0) < chunk ? (q_end - q0) : chunk);
      const long long T = (long long)Bc * A * A * P * P;
      Workspace w = carve(wsp, T, s);
      int r = launch_divide(h, lr_lf, w.lrp, h0, w0, nv, (int)q0, Bc, P, stride, st);
      if (r) return r;
      UpTarget up;
      up.mode = direct ? 2 : 1;
      up.crop_stride = stride;
      up.h0 = h0; up.w0 = w0; up.numV = nv; up.p0 = (int)q0;
      float* out = direct ? dst : dst + (q0 - p0) * crop_stride;
      if ((r = run_forward_chunk(h, w.lrp, out, w, Bc, P, up, st))) return r;
    }
    return 0;
  };
  return run_split(h, p1 - p0, ws, ws_bytes, per, (cudaStream_t)stream,
                   [&](long long a, long long b, void* wsp, long long cap, cudaStream_t st) { return run_range(p0 + a, p0 + b, wsp, cap, st); });
}

int lft_forward_lf_ex(lft_handle* hh, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                      int32_t p0, int32_t p1, float* sr_crops, void* ws, size_t ws_bytes, void* stream) {
  return forward_lf_impl(hh, lr_lf, h0, w0, patch, stride, p0, p1, sr_crops, false, ws, ws_bytes, stream);
}

int lft_forward_lf_sr(lft_handle* hh, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                      int32_t p0, int32_t p1, float* sr_lf, void* ws, size_t ws_bytes, void* stream) {
  return forward_lf_impl(hh, lr_lf, h0, w0, patch, stride, p0, p1, sr_lf, true, ws, ws_bytes, stream);
}

// ---- peer-visible device buffers (CUDA IPC): rank 0 allocates the assembled SR light field with lft_peer_alloc, sends the
// 64-byte handle to the other ranks of the node (any byte transport: torch.distributed broadcast), they map it with
// lft_peer_open and pass the mapped pointer to lft_forward_lf_sr, whose last kernel then stores the kept crops straight into
// rank 0's memory over NVLink.
int lft_peer_alloc(int32_t device, size_t bytes, void** dev_ptr, lft_peer_handle* handle) {
  if (!dev_ptr || !handle || bytes == 0) return fail(LFT_ERR_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(lft_peer_handle), "lft_peer_handle too small");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", device);
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(LFT_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memset(handle, 0, sizeof(*handle));
  memcpy(handle->bytes, &hd, sizeof(hd));
  *dev_ptr = p;
  return 0;
}

int lft_peer_free(int32_t device, void* dev_ptr) {
  if (!dev_ptr) return 0;
  DeviceGuard dg(device);
  CUDA_TRY(cudaFree(dev_ptr));
  return 0;
}

int lft_peer_open(int32_t device, const lft_peer_handle* handle, void** mapped_ptr) {
  if (!handle || !mapped_ptr) return fail(LFT_ERR_ARG, "bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", device);
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle->bytes, sizeof(hd));
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *mapped_ptr = p;
  return 0;
}

int lft_peer_close(int32_t device, void* mapped_ptr) {
  if (!mapped_ptr) return 0;
  DeviceGuard dg(device);
  CUDA_TRY(cudaIpcCloseMemHandle(mapped_ptr));
  return 0;
}

int lft_forward_lf(lft_handle* hh, const float* lr_lf, int32_t h0, int32_t w0, int32_t p0, int32_t p1, float* sr_crops,
                   void* ws, size_t ws_bytes, void* stream) {
  return lft_forward_lf_ex(hh, lr_lf, h0, w0, 32, 16, p0, p1, sr_crops, ws, ws_bytes, stream);
}

int lft_debug_timeline(int32_t which, int64_t* out64) {
  long long* o = reinterpret_cast<long long*>(out64);
  if (which == 4) return debug_timeline_ring_embed(o);  // needs out64[120]
  return which == 0 ? debug_timeline_spa(o) : (which == 1 ? debug_timeline_ang(o) : debug_timeline_embed(o));
}

int lft_mma_bench(int32_t N, int32_t K, int32_t reps, int32_t mode, int32_t grid, int32_t smem_bytes, int64_t* cycles) {
  if (!cycles || N % 16 || N < 16 || N > 256 || K % 16 || K > 256 || grid < 1) return fail(LFT_ERR_ARG, "bad mma bench shape");
  return launch_mma_bench(N, K, reps, mode, grid, smem_bytes, reinterpret_cast<long long*>(cycles));
}

int lft_gemm_selftest(const float* A, const float* W, float* D, float* aux, int32_t M, int32_t N, int32_t K,
                      int32_t precision, int32_t variant) {
  if (!A || !W || !D || !aux || M % 128 || K % 64 || N % 16 || N > 256 || N < 16 || K > 256)
    return fail(LFT_ERR_ARG, "bad selftest shape");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  int rc = configure_kernels(dev);
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
