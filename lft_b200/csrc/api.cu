// Host side of the C ABI (include/lft_b200.h): handle, strict state_dict intake, weight packing into
// bf16 hi/lo operand slabs, constant position-encoding tables, launch sequences, per-kernel profiling.
#include "../../include/lft_b200.h"
#include "host.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lft {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ------------------------------------------------------------------ bf16 helpers (host, RNE like the device)
static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Pack W[N][K] (accessor) into k-slabs: for ks in K/64: hi slab then lo slab, each [kc=8][n=Npad][8] bf16.
std::vector<uint16_t> pack_weight(int N, int Npad, int K, const std::function<float(int, int)>& w) {
  const int ks_n = K / 64;
  std::vector<uint16_t> out((size_t)ks_n * 2 * Npad * 64, 0);
  for (int ks = 0; ks < ks_n; ++ks)
    for (int kc = 0; kc < 8; ++kc)
      for (int n = 0; n < N; ++n)
        for (int e = 0; e < 8; ++e) {
          const float x = w(n, ks * 64 + kc * 8 + e);
          const uint16_t hi = f2bf(x);
          const uint16_t lo = f2bf(x - bf2f(hi));
          const size_t base = (size_t)ks * 2 * Npad * 64;
          out[base + ((size_t)kc * Npad + n) * 8 + e] = hi;
          out[base + (size_t)Npad * 64 + ((size_t)kc * Npad + n) * 8 + e] = lo;
        }
  return out;
}

// ------------------------------------------------------------------ spec of the reference state_dict
static void build_spec(Handle* h) {
  const int C = 64, S = 128, s2 = h->cfg.scale * h->cfg.scale;
  auto add = [&](const std::string& k, std::vector<int64_t> shp) { h->spec[k] = shp; };
  add("conv_init0.0.weight", {C, 1, 1, 3, 3});
  for (int i : {0, 2, 4}) add("conv_init." + std::to_string(i) + ".weight", {C, C, 1, 3, 3});
  for (int i = 0; i < kLayers; ++i) {
    std::string p = "altblock." + std::to_string(i) + ".spa_trans.";
    add(p + "MLP.weight", {S, 9 * C});
    add(p + "norm.weight", {S});
    add(p + "norm.bias", {S});
    add(p + "attention.in_proj_weight", {3 * S, S});
    add(p + "attention.out_proj.weight", {S, S});
    add(p + "feed_forward.0.weight", {S});
    add(p + "feed_forward.0.bias", {S});
    add(p + "feed_forward.1.weight", {2 * S, S});
    add(p + "feed_forward.4.weight", {S, 2 * S});
    add(p + "linear.0.weight", {C, S, 1, 1, 1});
    p = "altblock." + std::to_string(i) + ".ang_trans.";
    add(p + "norm.weight", {C});
    add(p + "norm.bias", {C});
    add(p + "attention.in_proj_weight", {3 * C, C});
    add(p + "attention.out_proj.weight", {C, C});
    add(p + "feed_forward.0.weight", {C});
    add(p + "feed_forward.0.bias", {C});
    add(p + "feed_forward.1.weight", {2 * C, C});
    add(p + "feed_forward.4.weight", {C, 2 * C});
  }
  add("upsampling.0.weight", {C * s2, C, 1, 1});
  add("upsampling.3.weight", {1, C, 3, 3});
}

int upload(Handle* h, const void* src, size_t bytes, void** dst) {
  void* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, bytes));
  h->allocs.push_back(d);
  CUDA_TRY(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  *dst = d;
  return 0;
}

static int upload_packed(Handle* h, int N, int Npad, int K, const std::function<float(int, int)>& w,
                         const uint8_t** dst) {
  std::vector<uint16_t> p = pack_weight(N, Npad, K, w);
  void* d;
  int rc = upload(h, p.data(), p.size() * 2, &d);
  *dst = (const uint8_t*)d;
  return rc;
}
static int upload_f32(Handle* h, const std::vector<float>& v, const float** dst) {
  void* d;
  int rc = upload(h, v.data(), v.size() * 4, &d);
  *dst = (const float*)d;
  return rc;
}

// One axis of PositionEncoding.forward (LFT.py:94-104) in fp32, as the reference computes it.
static std::vector<float> pos_axis(int length, int C) {
  std::vector<float> t((size_t)length * C);
  std::vector<float> grid(C);
  for (int i = 0; i < C; ++i) grid[i] = powf(10000.f, 2.f * (float)(i / 2) / (float)C);
  for (int p = 0; p < length; ++p)
    for (int j = 0; j < C / 2; ++j) {
      t[(size_t)p * C + j] = sinf((float)p / grid[2 * j]);
      t[(size_t)p * C + C / 2 + j] = cosf((float)p / grid[2 * j + 1]);
    }
  return t;
}

// Release every device allocation of the current weight generation (packed slabs, constant tables, the per-patch-size
// position tables).  cudaFree waits for the device, so kernels still reading the old weights finish first.
static void free_weights(Handle* h) {
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->pe_cache.clear();
  h->pe_P = -1;
  for (int i = 0; i < kLayers; ++i) h->layer[i] = Layer();
  for (int i = 0; i < 3; ++i) h->w_conv[i] = h->w_conv_st[i] = nullptr;
  h->w_up = h->w_up3 = nullptr;
  h->pe_ang = nullptr;
  h->finalized = false;
}

static int finalize(Handle* h) {
  const int C = 64, S = 128, A = h->cfg.ang_res, s = h->cfg.scale, s2 = s * s;
  for (auto& kv : h->spec)
    if (!h->host_w.count(kv.first)) return fail(LFT_ERR_STATE, "missing state_dict key '%s'", kv.first.c_str());
  free_weights(h);  // a reload replaces the previous generation instead of leaking it
  auto W = [&](const std::string& k) -> const float* { return h->host_w[k].data(); };
  int rc;
  // conv_init0: fp32 [64][9]
  {
    const float* w = W("conv_init0.0.weight");
    std::vector<float> v(w, w + 576);
    h->w_conv0_host = v;  // kernel parameter (constant bank) of the first / last conv of the stack
  }
  // conv weights [N][C][ky][kx] -> tap-major K: k = tap*64 + c
  auto conv_pack = [&](const float* w, int N, const uint8_t** dst) {
    return upload_packed(h, N, N, 576, [=](int n, int k) { return w[((size_t)n * 64 + (k % 64)) * 9 + (k / 64)]; }, dst);
  };
  for (int i = 0; i < 3; ++i) {
    if ((rc = conv_pack(W("conv_init." + std::to_string(2 * i) + ".weight"), 64, &h->w_conv[i]))) return rc;
    {  // stacked slabs for the fp32 path: per tap [k-chunk 8][rows 0..63 = bf16 hi | rows 64..127 = bf16 lo][8]
      const float* w = W("conv_init." + std::to_string(2 * i) + ".weight");
      std::vector<uint16_t> st((size_t)9 * 128 * 64);
      for (int t = 0; t < 9; ++t)
        for (int kc = 0; kc < 8; ++kc)
          for (int n = 0; n < 64; ++n)
            for (int e = 0; e < 8; ++e) {
              const float x = w[((size_t)n * 64 + kc * 8 + e) * 9 + t];
              const uint16_t hi = f2bf(x), lo = f2bf(x - bf2f(hi));
              st[(size_t)t * 128 * 64 + ((size_t)kc * 128 + n) * 8 + e] = hi;
              st[(size_t)t * 128 * 64 + ((size_t)kc * 128 + 64 + n) * 8 + e] = lo;
            }
      void* d;
      if ((rc = upload(h, st.data(), st.size() * 2, &d))) return rc;
      h->w_conv_st[i] = (const uint8_t*)d;
    }
  }
  auto lin_pack = [&](const float* w, int row0, int N, int ldk, int col0, int K, const uint8_t** dst) {
    return upload_packed(h, N, N, K, [=](int n, int k) { return w[(size_t)(row0 + n) * ldk + col0 + k]; }, dst);
  };
  // LayerNorm folded into the following linear layer: LN(z) W^T = rstd (z W'^T - mean u) + c with
  // W' = W diag(gamma), u = W' 1 (summed over the bf16 hi+lo values the MMA really uses), c = W beta.
  // u[0] sums hi + lo (what three passes multiply by), u[1] hi only (the single bf16 pass)
  auto fold_pack = [&](const float* w, int row0, int N, int K, const float* g, const float* b, const uint8_t** dst,
                       std::vector<float>* u, std::vector<float>& c, std::vector<float>* keep) -> int {
    std::vector<float> wf((size_t)N * K);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) wf[(size_t)n * K + k] = w[(size_t)(row0 + n) * K + k] * g[k];
    for (int n = 0; n < N; ++n) {
      double su = 0.0, sh = 0.0, sc = 0.0;
      for (int k = 0; k < K; ++k) {
        const float x = wf[(size_t)n * K + k];
        const uint16_t hi = f2bf(x);
        const uint16_t lo = f2bf(x - bf2f(hi));
        su += (double)bf2f(hi) + (double)bf2f(lo);
        sh += (double)bf2f(hi);
        sc += (double)w[(size_t)(row0 + n) * K + k] * (double)b[k];
      }
      u[0].push_back((float)su);
      u[1].push_back((float)sh);
      c.push_back((float)sc);
    }
    const float* wp = wf.data();
    int r = upload_packed(h, N, N, K, [=](int n, int k) { return wp[(size_t)n * K + k]; }, dst);
    if (keep) *keep = wf;
    return r;
  };
  std::vector<float> pe_a = pos_axis(A * A, C);
  for (int i = 0; i < kLayers; ++i) {
    Layer& L = h->layer[i];
    std::string p = "altblock." + std::to_string(i) + ".ang_trans.";
    const float* in_w = W(p + "attention.in_proj_weight");
    {
      std::vector<float> uqk[2], cqk, u1[2], c1, wqk_fold;
      if ((rc = fold_pack(in_w, 0, 128, C, W(p + "norm.weight"), W(p + "norm.bias"), &L.a_wqk, uqk, cqk, &wqk_fold)))
        return rc;
      if ((rc = fold_pack(W(p + "feed_forward.1.weight"), 0, 128, C, W(p + "feed_forward.0.weight"),
                          W(p + "feed_forward.0.bias"), &L.a_w1, u1, c1, nullptr)))
        return rc;
      const int NA = A * A;
      for (int m = 0; m < 2; ++m) {
        std::vector<float> tab;
        tab.insert(tab.end(), uqk[m].begin(), uqk[m].end());
        tab.insert(tab.end(), cqk.begin(), cqk.end());
        tab.insert(tab.end(), u1[m].begin(), u1[m].end());
        tab.insert(tab.end(), c1.begin(), c1.end());
        L.a_tab[m] = tab;
        std::vector<float> peqk((size_t)NA * 128);  // chunk-planar [n/4][a][4]; bf16 mode: with the weights that mode multiplies by
        for (int a = 0; a < NA; ++a)
          for (int n = 0; n < 128; ++n) {
            double acc = 0.0;
            for (int k = 0; k < C; ++k) {
              const float wv = wqk_fold[(size_t)n * C + k];
              acc += (double)pe_a[(size_t)a * C + k] * (double)(m ? bf2f(f2bf(wv)) : wv);
            }
            peqk[((size_t)(n / 4) * NA + a) * 4 + (n % 4)] = (float)acc;
          }
        if ((rc = upload_f32(h, peqk, &L.a_peqk[m]))) return rc;
      }
    }
    if ((rc = lin_pack(in_w, 128, 64, C, 0, C, &L.a_wv))) return rc;
    if ((rc = lin_pack(W(p + "attention.out_proj.weight"), 0, 64, C, 0, C, &L.a_wo))) return rc;
    if ((rc = lin_pack(W(p + "feed_forward.4.weight"), 0, 64, 2 * C, 0, 2 * C, &L.a_w2))) return rc;
    {
      std::vector<float> ln(4 * C);
      memcpy(&ln[0], W(p + "norm.weight"), C * 4);
      memcpy(&ln[C], W(p + "norm.bias"), C * 4);
      memcpy(&ln[2 * C], W(p + "feed_forward.0.weight"), C * 4);
      memcpy(&ln[3 * C], W(p + "feed_forward.0.bias"), C * 4);
      if ((rc = upload_f32(h, ln, &L.a_ln))) return rc;
    }
    p = "altblock." + std::to_string(i) + ".spa_trans.";
    if ((rc = conv_pack(W(p + "MLP.weight"), 128, &L.s_wmlp))) return rc;  // [128][64*9]: c*9+tap (LFT.py:167)
    in_w = W(p + "attention.in_proj_weight");
    {
      std::vector<float> uq[2], cq, uk[2], ck, u1[2], c1;
      const float *g1 = W(p + "norm.weight"), *b1 = W(p + "norm.bias");
      const float *g2 = W(p + "feed_forward.0.weight"), *b2 = W(p + "feed_forward.0.bias");
      if ((rc = fold_pack(in_w, 0, 128, S, g1, b1, &L.s_wq, uq, cq, nullptr))) return rc;
      if ((rc = fold_pack(in_w, 128, 128, S, g1, b1, &L.s_wk, uk, ck, nullptr))) return rc;
      if ((rc = fold_pack(W(p + "feed_forward.1.weight"), 0, 128, S, g2, b2, &L.s_w1a, u1, c1, nullptr))) return rc;
      if ((rc = fold_pack(W(p + "feed_forward.1.weight"), 128, 128, S, g2, b2, &L.s_w1b, u1, c1, nullptr))) return rc;
      for (int m = 0; m < 2; ++m) {
        std::vector<float> tab;
        tab.insert(tab.end(), uq[m].begin(), uq[m].end());
        tab.insert(tab.end(), uk[m].begin(), uk[m].end());
        tab.insert(tab.end(), cq.begin(), cq.end());
        tab.insert(tab.end(), ck.begin(), ck.end());
        tab.insert(tab.end(), u1[m].begin(), u1[m].end());
        tab.insert(tab.end(), c1.begin(), c1.end());
        L.s_tab[m] = tab;
      }
    }
    if ((rc = lin_pack(in_w, 256, 128, S, 0, S, &L.s_wv))) return rc;
    if ((rc = lin_pack(W(p + "attention.out_proj.weight"), 0, 128, S, 0, S, &L.s_wo))) return rc;
    if ((rc = lin_pack(W(p + "feed_forward.4.weight"), 0, 128, 2 * S, 0, S, &L.s_w2a))) return rc;
    if ((rc = lin_pack(W(p + "feed_forward.4.weight"), 0, 128, 2 * S, S, S, &L.s_w2b))) return rc;
    if ((rc = lin_pack(W(p + "linear.0.weight"), 0, 64, S, 0, S, &L.s_wlin))) return rc;
    {
      std::vector<float> ln(4 * S);
      memcpy(&ln[0], W(p + "norm.weight"), S * 4);
      memcpy(&ln[S], W(p + "norm.bias"), S * 4);
      memcpy(&ln[2 * S], W(p + "feed_forward.0.weight"), S * 4);
      memcpy(&ln[3 * S], W(p + "feed_forward.0.bias"), S * 4);
      if ((rc = upload_f32(h, ln, &L.s_ln))) return rc;
    }
  }
  // upsampling 1x1: rows permuted to n' = ij*64 + c  (PixelShuffle: in channel c*s^2 + i*s + j, LFT.py:41)
  {
    const float* w = W("upsampling.0.weight");
    std::vector<uint16_t> all;
    for (int ij = 0; ij < s2; ++ij) {  // one [64 x 64] GEMM per sub-pixel ij: rows c -> original row c*s^2 + ij
      std::vector<uint16_t> p1 = pack_weight(64, 64, 64, [=](int n, int k) { return w[(size_t)(n * s2 + ij) * 64 + k]; });
      all.insert(all.end(), p1.begin(), p1.end());
    }
    void* d;
    if ((rc = upload(h, all.data(), all.size() * 2, &d))) return rc;
    h->w_up = (const uint8_t*)d;
    const float* w3 = W("upsampling.3.weight");  // [1][64][3][3] -> B operand of the tap GEMM: row = tap (9 of 16), k = channel
    if ((rc = upload_packed(h, 9, 16, 64, [=](int n, int k) { return w3[(size_t)k * 9 + n]; }, &h->w_up3))) return rc;
  }
  // angular position table, chunk-planar [c/4][A*A][4]
  {
    const int NA = A * A;
    std::vector<float> pl((size_t)NA * C);
    for (int a = 0; a < NA; ++a)
      for (int c = 0; c < C; ++c) pl[((size_t)(c / 4) * NA + a) * 4 + (c % 4)] = pe_a[(size_t)a * C + c];
    if ((rc = upload_f32(h, pl, &h->pe_ang))) return rc;
  }
  h->finalized = true;
  return 0;
}

// spatial PE token table per layer: SAI2Token(spa_position) (LFT.py:180) = conv3x3(PE_hw, MLP.weight), [P*P][128]
// Built on the host once per (weights, patch size) and cached: alternating patch sizes costs nothing after the first use.
int ensure_spa_pe(Handle* h, int P) {
  if (h->pe_P == P) return 0;
  {
    auto it = h->pe_cache.find(P);
    if (it != h->pe_cache.end()) {
      for (int i = 0; i < kLayers; ++i) {
        h->layer[i].s_pe = it->second.pe[i];
        for (int m = 0; m < 2; ++m) h->layer[i].s_pev[m] = it->second.pev[i][m];
      }
      h->pe_P = P;
      return 0;
    }
  }
  const int C = 64, S = 128;
  std::vector<float> ax = pos_axis(P, C);
  std::vector<float> pe((size_t)P * P * C);
  for (int y = 0; y < P; ++y)
    for (int x = 0; x < P; ++x)
      for (int c = 0; c < C; ++c) pe[((size_t)y * P + x) * C + c] = (ax[(size_t)y * C + c] + ax[(size_t)x * C + c]) / 2.f;
  for (int i = 0; i < kLayers; ++i) {
    const float* w = h->host_w["altblock." + std::to_string(i) + ".spa_trans.MLP.weight"].data();
    std::vector<float> tab((size_t)P * P * S);
    for (int y = 0; y < P; ++y)
      for (int x = 0; x < P; ++x)
        for (int n = 0; n < S; ++n) {
          double acc = 0.0;
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= P) continue;
            for (int kx = 0; kx < 3; ++kx) {
              const int xx = x + kx - 1;
              if (xx < 0 || xx >= P) continue;
              const float* pv = &pe[((size_t)yy * P + xx) * C];
              const float* wv = w + (size_t)n * 576 + ky * 3 + kx;
              for (int c = 0; c < C; ++c) acc += (double)pv[c] * (double)wv[(size_t)c * 9];
            }
          }
          tab[((size_t)y * P + x) * S + n] = (float)acc;
        }
    const int PPn = P * P;
    std::vector<float> tab_pl((size_t)PPn * S);  // chunk-planar [n/4][p][4]
    for (int t = 0; t < PPn; ++t)
      for (int n = 0; n < S; ++n) tab_pl[((size_t)(n / 4) * PPn + t) * 4 + (n % 4)] = tab[(size_t)t * S + n];
    int rc = upload_f32(h, tab_pl, &h->layer[i].s_pe);
    if (rc) return rc;
    // PE_s Wv^T: V = tok Wv^T is computed from the operand z = tok + PE_s, so this constant is subtracted
    const float* wv = h->host_w["altblock." + std::to_string(i) + ".spa_trans.attention.in_proj_weight"].data() +
                      (size_t)256 * S;
    for (int m = 0; m < 2; ++m) {  // bf16 mode: with the bf16-rounded Wv that mode multiplies z = tok + PE_s by
      std::vector<float> wvm((size_t)S * S);
      for (size_t j = 0; j < wvm.size(); ++j) wvm[j] = m ? bf2f(f2bf(wv[j])) : wv[j];
      std::vector<float> pvt((size_t)PPn * S);
      for (int t = 0; t < PPn; ++t)
        for (int n = 0; n < S; ++n) {
          double a = 0.0;
          for (int k = 0; k < S; ++k) a += (double)tab[(size_t)t * S + k] * (double)wvm[(size_t)n * S + k];
          pvt[((size_t)(n / 4) * PPn + t) * 4 + (n % 4)] = (float)a;
        }
      if ((rc = upload_f32(h, pvt, &h->layer[i].s_pev[m]))) return rc;
    }
  }
  SpaPe& e = h->pe_cache[P];
  for (int i = 0; i < kLayers; ++i) {
    e.pe[i] = h->layer[i].s_pe;
    for (int m = 0; m < 2; ++m) e.pev[i][m] = h->layer[i].s_pev[m];
  }
  h->pe_P = P;
  return 0;
}

}  // namespace lft

using namespace lft;

extern "C" {

const char* lft_last_error(void) { return g_err.c_str(); }
int lft_version(void) { return 100; }

int lft_create(const lft_config* cfg, lft_handle** out) {
  if (!cfg || !out) return fail(LFT_ERR_ARG, "null argument");
  if (cfg->channels != 64) return fail(LFT_ERR_ARG, "channels must be 64 (got %d)", cfg->channels);
  if (cfg->ang_res < 2 || cfg->ang_res > 9) return fail(LFT_ERR_ARG, "angRes %d unsupported (2..9)", cfg->ang_res);
  if (cfg->scale != 2 && cfg->scale != 4) return fail(LFT_ERR_ARG, "scale_factor must be 2 or 4");
  if (cfg->precision != LFT_PREC_FP32 && cfg->precision != LFT_PREC_BF16) return fail(LFT_ERR_ARG, "bad precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(LFT_ERR_CUDA, "no CUDA device: lft_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(LFT_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
  DeviceGuard dg(cfg->device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", cfg->device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(LFT_ERR_CUDA, "device is sm_%d%d; this library is sm_100a only", prop.major, prop.minor);
  Handle* h = new Handle();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  {
    const char* e = getenv("LFT_STREAMS");
    h->two_streams = !(e && e[0] == '1' && e[1] == 0);
    if (h->two_streams) {
      if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        h->two_streams = false;
      }
    }
  }
  {  // programmatic dependent launch: built, measured, off by default (see host.h) - LFT_PDL=1 turns it on
    const char* e = getenv("LFT_PDL");
    h->pdl = e && e[0] && e[0] != '0';
  }
  build_spec(h);
  int rc = configure_kernels(cfg->device);
  if (rc) { delete h; return rc; }
  *out = reinterpret_cast<lft_handle*>(h);
  return 0;
}

int lft_destroy(lft_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return 0;
  DeviceGuard dg(h->cfg.device);
  free_weights(h);
  for (auto& e : h->events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return 0;
}

int lft_set_weight(lft_handle* hh, const char* key, const float* host_data, const int64_t* shape, int32_t ndim) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !key || !host_data || !shape) return fail(LFT_ERR_ARG, "null argument");
  auto it = h->spec.find(key);
  if (it == h->spec.end()) return fail(LFT_ERR_KEY, "unexpected state_dict key '%s'", key);
  const std::vector<int64_t>& want = it->second;
  bool ok = (int)want.size() == ndim;
  size_t n = 1;
  for (int i = 0; ok && i < ndim; ++i) { ok = want[i] == shape[i]; n *= (size_t)shape[i]; }
  if (!ok) return fail(LFT_ERR_KEY, "shape mismatch for '%s'", key);
  h->host_w[key].assign(host_data, host_data + n);
  h->finalized = false;
  return 0;
}

int lft_finalize_weights(lft_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  DeviceGuard dg(h->cfg.device);
  if (!dg.ok) return fail(LFT_ERR_CUDA, "cannot select device %d", h->cfg.device);
  return finalize(h);
}

int lft_set_precision(lft_handle* hh, int32_t precision) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || (precision != LFT_PREC_FP32 && precision != LFT_PREC_BF16)) return fail(LFT_ERR_ARG, "bad precision");
  h->cfg.precision = precision;
  return 0;
}

int lft_profile_enable(lft_handle* hh, int32_t on) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return fail(LFT_ERR_ARG, "null handle");
  h->profiling = on != 0;
  if (on) h->n_events = 0;  // a new session reuses the pooled events
  return 0;
}

int lft_profile_read(lft_handle* hh, int32_t* n_kinds, const char** names, int64_t* launches, double* total_ms) {
  return lft_profile_read2(hh, n_kinds, names, launches, total_ms, nullptr);
}

int lft_profile_read2(lft_handle* hh, int32_t* n_kinds, const char** names, int64_t* launches, double* total_ms,
                      int64_t* units) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !n_kinds || !names || !launches || !total_ms) return fail(LFT_ERR_ARG, "null argument");
  for (int k = 0; k < K_COUNT; ++k) { names[k] = kKindNames[k]; launches[k] = 0; total_ms[k] = 0.0; if (units) units[k] = 0; }
  DeviceGuard dg(h->cfg.device);
  for (size_t i = 0; i < h->n_events; ++i) {
    const ProfEvent& e = h->events[i];
    CUDA_TRY(cudaEventSynchronize(e.stop));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e.start, e.stop));
    launches[e.kind] += 1;
    total_ms[e.kind] += ms;
    if (units) units[e.kind] += e.units;
  }
  *n_kinds = K_COUNT;
  return 0;
}

int64_t lft_launch_count(lft_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  return h ? h->launches : -1;
}

}  // extern "C"
