/*
 * lft_b200 - C ABI of the B200-native LFT inference forward path.
 *
 * The reference (HydrogenSulfate/LFT) is pure Python and has no FFI of its own; its only plug-in
 * point is `importlib.import_module('model.' + args.model_name).get_model(args)` (test.py:29-31,
 * train.py:31-33) and the nn.Module contract of `model/LFT.py`.  The Python drop-in
 * (`lft_b200/model.py::get_model`) honours that contract and calls THIS library through ctypes;
 * each entry point names the reference code it replaces.  See INTEGRATION.md for the binding.
 *
 * Conventions: every function returns 0 on success or a negative lft_status; the message of the
 * last failure on the calling thread is available from lft_last_error().  Nothing throws across
 * the boundary.  The library never takes ownership of caller pointers.  All device work is
 * enqueued on the caller-supplied CUDA stream (passed as void* = cudaStream_t).  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with LFT_ERR_CUDA.
 *
 * Tensor layouts at the boundary are the reference's own:
 *   lr  : float32 [B, 1, A*h, A*w]      SAI mosaic (view (u,v) at rows u*h.., cols v*w..), LFT.py:52
 *   sr  : float32 [B, 1, A*h*s, A*w*s]                                                      LFT.py:83
 * Internal stage tensors are channels-last tokens: feat [B, A*A, h, w, C] float32.
 */
#ifndef LFT_B200_H
#define LFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lft_handle lft_handle;

typedef enum {
  LFT_OK = 0,
  LFT_ERR_ARG = -1,        /* bad argument / unsupported shape */
  LFT_ERR_CUDA = -2,       /* CUDA runtime error (no device, launch failure, ...) */
  LFT_ERR_STATE = -3,      /* weights missing / not finalized */
  LFT_ERR_WORKSPACE = -4,  /* workspace too small */
  LFT_ERR_KEY = -5         /* unknown state_dict key or shape mismatch */
} lft_status;

/* precision of the contractions (all other arithmetic is fp32):
 *   LFT_PREC_FP32  : every product as 3 bf16 tcgen05 MMAs (hi*hi + lo*hi + hi*lo), fp32 accumulate
 *                    -> matches the fp32 reference to ~1e-5 (gate: max-abs 1e-4)
 *   LFT_PREC_BF16  : single bf16 MMA (gate: PSNR delta <= 0.01 dB) */
typedef enum { LFT_PREC_FP32 = 0, LFT_PREC_BF16 = 1 } lft_precision;

/* Mirrors what get_model.__init__ reads from `args` (LFT.py:11-20): channels, angRes, scale_factor;
 * layer_num=4 and num_heads=8 are hard-coded in the reference and fixed here as well. */
typedef struct {
  int32_t ang_res;      /* A: 3, 5, 7 or 9 */
  int32_t scale;        /* s: 2 or 4 */
  int32_t channels;     /* must be 64 */
  int32_t precision;    /* lft_precision */
  int32_t device;       /* CUDA device ordinal */
} lft_config;

const char* lft_last_error(void);
int lft_version(void);

/* get_model.__init__ (LFT.py:9-50) */
int lft_create(const lft_config* cfg, lft_handle** out);
int lft_destroy(lft_handle* h);

/* net.load_state_dict (test.py:37-51): one call per state_dict entry, `key` exactly as in the
 * reference state_dict (no 'module.' prefix), `host_data` fp32 in HOST memory, row-major `shape`.
 * lft_finalize_weights checks that all 78 tensors arrived (strict), packs them into the bf16 hi/lo
 * operand slabs and uploads them; it must be called before any compute entry point. */
int lft_set_weight(lft_handle* h, const char* key, const float* host_data, const int64_t* shape, int32_t ndim);
int lft_finalize_weights(lft_handle* h);
int lft_set_precision(lft_handle* h, int32_t precision);

/* bytes of device scratch lft_forward needs for a batch of B patches of P x P pixels per view.
 * Threading: a handle owns no scratch of its own, but every compute call on it uses the workspace it is given for the whole
 * call - two calls that overlap in time (two streams, two threads) need two workspaces; weight reloads
 * (lft_set_weight / lft_finalize_weights) must not overlap compute calls on the same handle. */
int lft_workspace_bytes(lft_handle* h, int32_t B, int32_t P, size_t* bytes);

/* get_model.forward (LFT.py:52-83). lr/sr are DEVICE pointers in the layouts above. */
int lft_forward(lft_handle* h, const float* lr, float* sr, int32_t B, int32_t P, void* workspace, size_t ws_bytes,
                void* stream);

/* Full light-field path = test.py:83-101 (LFdivide -> net per patch -> LFintegrate) for the patch
 * range [patch_begin, patch_end) of the numU*numV patches (row-major kh*numV+kw, utils.py:115-116).
 *   lr_lf    : device float32 [A*h0, A*w0]  (Lr_SAI_y)
 *   sr_crops : device float32 [patch_end-patch_begin, A, A, 16*s, 16*s]  kept central crops
 * lft_integrate scatters crops of the given patch range into sr_lf [A*h0*s, A*w0*s]
 * (utils.py:141-157 + test.py:100-101), clipping the ragged last row/column. */
int lft_lf_num_patches(int32_t h0, int32_t w0, int32_t* numU, int32_t* numV);
int lft_forward_lf(lft_handle* h, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch_begin, int32_t patch_end,
                   float* sr_crops, void* workspace, size_t ws_bytes, void* stream);
int lft_integrate(lft_handle* h, const float* sr_crops, int32_t h0, int32_t w0, int32_t patch_begin,
                  int32_t patch_end, float* sr_lf, void* stream);
/* LFdivide alone (utils.py:91-138): patches [patch_end-patch_begin, 1, A*32, A*32] */
int lft_divide(lft_handle* h, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch_begin, int32_t patch_end,
               float* patches, void* stream);

/* The same four entry points for test.py's `--patch_size_for_test` / `--stride_for_test` (option.py:16-17; test.py:83-96
 * passes them to LFdivide / LFintegrate; the entry points above are these with patch = 32, stride = 16):
 *   patch  : 4..32 (the kernels' patch-size range), stride : 1..patch, bdr = (patch - stride) / 2 (utils.py:95)
 *   patches  [n, 1, A*patch, A*patch];  sr_crops [n, A, A, stride*s, stride*s] = the block
 *   [c0, c0 + stride*s)^2 of every SR patch view with c0 = ((patch - stride)*s) / 2 (utils.py:145,152).
 * Views smaller than the mirror border, or tilings the reference would return empty, are LFT_ERR_ARG. */
int lft_lf_num_patches_ex(int32_t h0, int32_t w0, int32_t patch, int32_t stride, int32_t* numU, int32_t* numV);
int lft_divide_ex(lft_handle* h, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                  int32_t patch_begin, int32_t patch_end, float* patches, void* stream);
int lft_forward_lf_ex(lft_handle* h, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                      int32_t patch_begin, int32_t patch_end, float* sr_crops, void* workspace, size_t ws_bytes,
                      void* stream);
int lft_integrate_ex(lft_handle* h, const float* sr_crops, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                     int32_t patch_begin, int32_t patch_end, float* sr_lf, void* stream);

/* The same path with LFintegrate (utils.py:141-157) and the SAI re-mosaic of test.py:100-101 fused into the last kernel:
 * the kept crop of every patch in [patch_begin, patch_end) is stored at its final position in the assembled SR light field
 *   sr_lf : device float32 [A*h0*s, A*w0*s]   (`Sr_SAI_y`)
 * (ragged last row / column clipped).  Patches are independent, so ranks that each own a patch range may all write into ONE
 * sr_lf: on a multi-GPU node sr_lf may be the peer mapping (lft_peer_open) of a buffer that lives on another GPU - the
 * stores then travel over NVLink and no gather collective is needed, only a barrier before the owner reads the result.
 * Nothing else is written; pixels of patches outside the range keep their previous contents. */
int lft_forward_lf_sr(lft_handle* h, const float* lr_lf, int32_t h0, int32_t w0, int32_t patch, int32_t stride,
                      int32_t patch_begin, int32_t patch_end, float* sr_lf, void* workspace, size_t ws_bytes,
                      void* stream);

/* Peer-visible device buffers for lft_forward_lf_sr (CUDA IPC, one process per GPU on one node; the reference has no
 * multi-GPU inference at all - test.py:18 pins one device).  The owner allocates with lft_peer_alloc and ships the 64-byte
 * handle to the other processes by any transport; they map it with lft_peer_open (peer access is enabled lazily by the
 * runtime) and unmap it with lft_peer_close before the owner frees it with lft_peer_free. */
typedef struct { unsigned char bytes[64]; } lft_peer_handle;
int lft_peer_alloc(int32_t device, size_t bytes, void** dev_ptr, lft_peer_handle* handle);
int lft_peer_free(int32_t device, void* dev_ptr);
int lft_peer_open(int32_t device, const lft_peer_handle* handle, void** mapped_ptr);
int lft_peer_close(int32_t device, void* mapped_ptr);

/* Stage-level entry points (device pointers, channels-last tokens [B, A*A, P, P, C]) for parity tests
 * against the reference's own sub-modules:
 *   conv_init : conv_init0 + conv_init + residual           LFT.py:65-66   lr [B,1,A*P,A*P] -> [T,64]
 *   ang       : altblock[layer].ang_trans                   LFT.py:225-238 [T,64] -> [T,64]
 *   spa       : altblock[layer].spa_trans                   LFT.py:176-191 [T,64] -> [T,64]
 *   upsample  : upsampling(mosaic) + bicubic(lr)            LFT.py:79-81   [T,64], lr -> sr */
int lft_stage_conv_init(lft_handle* h, const float* lr, float* feat, int32_t B, int32_t P, void* ws, size_t ws_bytes,
                        void* stream);
int lft_stage_ang(lft_handle* h, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream);
int lft_stage_spa(lft_handle* h, int32_t layer, const float* in, float* out, int32_t B, int32_t P, void* ws,
                  size_t ws_bytes, void* stream);
int lft_stage_upsample(lft_handle* h, const float* feat, const float* lr, float* sr, int32_t B, int32_t P, void* ws,
                       size_t ws_bytes, void* stream);

/* Per-kernel device timing (CUDA events on the launching stream around every launch while enabled).
 * lft_profile_read syncs the recorded events and returns, per kernel kind, launches and total ms.
 * names: static strings, valid for the life of the library. */
#define LFT_PROFILE_MAX_KINDS 16
int lft_profile_enable(lft_handle* h, int32_t on);
int lft_profile_read(lft_handle* h, int32_t* n_kinds, const char** names, int64_t* launches, double* total_ms);
/* the same plus, per kind, the units the recorded launches processed (LR tokens = pixels x views for the network kernels - on
 * the light-field path the last layers run on the pixels the kept crop depends on only -, output pixels for the tilers) */
int lft_profile_read2(lft_handle* h, int32_t* n_kinds, const char** names, int64_t* launches, double* total_ms,
                      int64_t* units);
int64_t lft_launch_count(lft_handle* h); /* kernels launched by this handle since creation */

/* Debug: per-phase clock64() marks of the middle CTA of the last k_spa_ffn (which=0) / k_ang (1) / k_spa_embed_qkv (2) launch
 * (row warp: out[0..31], MMA thread: out[32..63]); only in builds with -DLFT_TIMELINE, else LFT_ERR_STATE. */
int lft_debug_timeline(int32_t which, int64_t* out64);

/* Bring-up self test of the tcgen05 GEMM machinery: D[M x N] = A[M x K] * W[N x K]^T (M multiple of 128,
 * K multiple of 64, N multiple of 16 <= 256), host pointers in/out; aux[M x 16] = 2*A[:, :16]+1 via TMEM.
 * variant 0: both operands from shared memory (SS, what the kernels use); variant 2: A operand from tensor memory (TS). */
int lft_gemm_selftest(const float* A, const float* W, float* D, float* aux, int32_t M, int32_t N, int32_t K,
                      int32_t precision, int32_t variant);

/* Micro-benchmark of the tcgen05 issue/operand path: every CTA issues reps*(K/16) MMAs (128 x N x 16, bf16) on resident
 * operands; cycles[grid] = clock64 cycles from first issue to completion.  mode % 16: 0 A,B from smem (SS); 1 A from TMEM (TS);
 * 2 SS with the A start shifted by one 16-byte row; 3 / 4 A in the SWIZZLE_128B layout (4: shifted by one row); 5 / 6 / 7 the
 * 3x3-conv issue pattern (108 MMAs per rep; k-chunk planes of 201 rows / 208 rows / 201 rows without tap shifts).
 * mode / 16 = n > 0: an extra tcgen05.commit after every n-th group of K/16 MMAs (commit cost).
 * smem_bytes sets the dynamic shared memory per CTA (and thereby how many CTAs share an SM). */
int lft_mma_bench(int32_t N, int32_t K, int32_t reps, int32_t mode, int32_t grid, int32_t smem_bytes, int64_t* cycles);

#ifdef __cplusplus
}
#endif
#endif /* LFT_B200_H */
