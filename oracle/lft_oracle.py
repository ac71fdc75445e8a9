"""ORACLE — CPU restatement of the reference LFT inference forward path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it. The product path
(`lft_b200`) never imports anything under `oracle/` and has no CPU fallback.

What it restates (all citations into the reference tree, `model/LFT.py` unless noted):
  * get_model.forward                LFT.py:52-83
  * PositionEncoding.forward         LFT.py:91-115
  * conv_init0 / conv_init           LFT.py:23-33, 65-66
  * AngTrans.forward (+MHA)          LFT.py:194-238
  * SpaTrans.gen_mask/SAI2Token/forward/Token2SAI   LFT.py:147-191
  * upsampling (1x1, PixelShuffle, LeakyReLU, 3x3 on the view mosaic)   LFT.py:39-44, 79-80
  * interpolate (per-view bicubic, A=-0.75, align_corners=False)        LFT.py:255-266
  * LFdivide / ImageExtend / LFintegrate    utils/utils.py:91-157
  * the per-patch test loop          test.py:83-101

The arithmetic the reference relies on lives in PyTorch (nn.MultiheadAttention ->
F.multi_head_attention_forward -> scaled_dot_product_attention; F.unfold; F.interpolate), which is
not vendored in the reference. The semantics restated here: q/k/v use `in_proj_weight.chunk(3)`
separately because `key is not value`; heads are contiguous channel blocks; scale 1/sqrt(hd);
additive float mask; LayerNorm eps 1e-5 with biased variance; bicubic taps with A=-0.75 and
index clamping. Everything is written out with plain tensor ops (matmul, softmax, conv2d) in a
functional, state_dict-driven style - no nn.Module, no nn.MultiheadAttention, no F.interpolate.

Pinning: the reference ships no tests or golden vectors. This oracle is pinned
against outputs of the reference itself, executed in the authoring container by
`tests/golden/make_golden.py` (imports /root/reference/model/LFT.py and utils/utils.py unmodified)
and committed under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks it on every run.

Two attention execution modes for SpaTrans:
  mode="dense"  - as executed by the reference: the [hw,hw] 0/-inf mask is rebuilt per call
                  with the same per-pixel loop (LFT.py:147-162) and attention is dense.
                  Used for the CPU baseline (it is what test.py costs).
  mode="window" - the mathematically identical 5x5 clamped-window gather (<=25 keys/query).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LRELU = 0.2
LN_EPS = 1e-5
HEADS = 8
TEMPERATURE = 10000.0


# ----------------------------------------------------------------------------- helpers
def _ln(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """LayerNorm over the last dim, biased variance, eps 1e-5 (nn.LayerNorm; LFT.py:127,137,201,208)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * g + b


def _lrelu(x: torch.Tensor) -> torch.Tensor:
    return torch.where(x >= 0, x, x * LRELU)


def pos_table(length: int, C: int, dtype=torch.float32) -> torch.Tensor:
    """One axis of PositionEncoding (LFT.py:94-104): [length, C]; first C/2 channels sin of the
    even-indexed columns, last C/2 cos of the odd-indexed columns (concat, not interleaved).
    The reference builds this in fp32 on the CPU (LFT.py:94,103); the fp32 ops are kept so the
    table is bit-identical, then cast."""
    grid = torch.linspace(0, C - 1, C, dtype=torch.float32)
    grid = 2 * (grid // 2) / C
    grid = TEMPERATURE ** grid
    pos = torch.linspace(0, length - 1, length, dtype=torch.float32).view(-1, 1) / grid
    tab = torch.cat([pos[:, 0::2].sin(), pos[:, 1::2].cos()], dim=1)
    return tab.to(dtype)


def ang_position(A: int, C: int, dtype=torch.float32) -> torch.Tensor:
    """pos_encoding(dim=[2]) (LFT.py:70): [A*A, C] over the linear view index."""
    return pos_table(A * A, C, dtype)


def spa_position(h: int, w: int, C: int, dtype=torch.float32) -> torch.Tensor:
    """pos_encoding(dim=[3,4]) (LFT.py:69,106-115): (PE(y)+PE(x))/2 -> [h, w, C]."""
    ty = pos_table(h, C, torch.float32)
    tx = pos_table(w, C, torch.float32)
    return ((ty[:, None, :] + tx[None, :, :]) / 2).to(dtype)


def gen_mask_loop(h: int, w: int, k: int = 5) -> torch.Tensor:
    """SpaTrans.gen_mask as executed (LFT.py:147-162), including the `min(h, j+k_right)` column
    clamp (LFT.py:155; harmless for h == w). Returns [hw, hw] with 0 / -inf."""
    m = torch.zeros(h, w, h, w)
    kl = k // 2
    kr = k - kl
    for i in range(h):
        for j in range(w):
            t = torch.zeros(h, w)
            t[max(0, i - kl):min(h, i + kr), max(0, j - kl):min(h, j + kr)] = 1
            m[i, j] = t
    m = m.reshape(h * w, h * w)
    return torch.zeros_like(m).masked_fill(m == 0, float("-inf"))


def window_index(h: int, w: int, k: int = 5):
    """For every query pixel the <=k*k in-image keys of its clamped window: (idx [hw,k*k] long,
    valid [hw,k*k] bool). Equivalent to the finite entries of gen_mask for h == w."""
    r = k // 2
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    dy, dx = torch.meshgrid(torch.arange(-r, r + 1), torch.arange(-r, r + 1), indexing="ij")
    ky = yy.reshape(-1, 1) + dy.reshape(1, -1)
    kx = xx.reshape(-1, 1) + dx.reshape(1, -1)
    valid = (ky >= 0) & (ky < h) & (kx >= 0) & (kx < w)
    idx = ky.clamp(0, h - 1) * w + kx.clamp(0, w - 1)
    return idx, valid


# ----------------------------------------------------------------------------- stages
def conv_stack(lr_views: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """conv_init0 then conv_init + residual (LFT.py:23-33,65-66). Conv3d with kernel (1,3,3) and
    padding (0,1,1) is a per-view zero-padded 2-D conv. lr_views [V,1,h,w] -> [V,C,h,w]."""
    w0 = sd["conv_init0.0.weight"][:, :, 0]
    buf0 = F.conv2d(lr_views, w0, padding=1)
    x = buf0
    for i in (0, 2, 4):
        x = _lrelu(F.conv2d(x, sd[f"conv_init.{i}.weight"][:, :, 0], padding=1))
    return x + buf0


def _mha(q_in: torch.Tensor, v_in: torch.Tensor, w_in: torch.Tensor, w_out: torch.Tensor,
         mask: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.MultiheadAttention(query=key=q_in, value=v_in) with bias=False, dropout 0, as used at
    LFT.py:183-187,229-232. q_in, v_in: [S(sequences), L, E]. Dense softmax(QK^T/sqrt(hd)+mask)V."""
    S, L, E = q_in.shape
    hd = E // HEADS
    wq, wk, wv = w_in[:E], w_in[E:2 * E], w_in[2 * E:]
    q = (q_in @ wq.t()).view(S, L, HEADS, hd).transpose(1, 2)
    k = (q_in @ wk.t()).view(S, L, HEADS, hd).transpose(1, 2)
    v = (v_in @ wv.t()).view(S, L, HEADS, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(S, L, E)
    return o @ w_out.t()


def ang_trans(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, pe_a: torch.Tensor) -> torch.Tensor:
    """AngTrans.forward (LFT.py:225-238). x: [B, N, h, w, C] (channels-last view of `b c a h w`).
    Sequences = pixels, length N = A*A. V is projected from the raw tokens (no PE, no norm)."""
    B, N, h, w, C = x.shape
    tok = x.permute(0, 2, 3, 1, 4).reshape(B * h * w, N, C)
    tn = _ln(tok + pe_a[None], sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    tok = _mha(tn, tok, sd[pre + "attention.in_proj_weight"], sd[pre + "attention.out_proj.weight"], None) + tok
    f = _ln(tok, sd[pre + "feed_forward.0.weight"], sd[pre + "feed_forward.0.bias"])
    f = torch.relu(f @ sd[pre + "feed_forward.1.weight"].t()) @ sd[pre + "feed_forward.4.weight"].t()
    tok = f + tok
    return tok.view(B, h, w, N, C).permute(0, 3, 1, 2, 4).contiguous()


def spa_embed(x_vchw: torch.Tensor, mlp_w: torch.Tensor) -> torch.Tensor:
    """SpaTrans.SAI2Token (LFT.py:164-169): F.unfold(3x3, pad 1) (channel-major: c*9+ky*3+kx) then
    Linear(9C -> 2C) == a zero-padded 3x3 conv with weight MLP.weight.view(2C, C, 3, 3).
    [V,C,h,w] -> [V, h*w, 2C]."""
    S, K = mlp_w.shape
    C = K // 9
    y = F.conv2d(x_vchw, mlp_w.view(S, C, 3, 3), padding=1)
    return y.flatten(2).transpose(1, 2)


def spa_trans(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, pe_hw: torch.Tensor,
              mode: str = "window") -> torch.Tensor:
    """SpaTrans.forward (LFT.py:176-191) + Token2SAI/linear (171-174). x: [B,N,h,w,C] -> same shape.
    Output REPLACES the feature map (no residual at this level, LFT.py:248-252)."""
    B, N, h, w, C = x.shape
    assert h == w, "reference gen_mask is only correct for square patches (LFT.py:155)"
    mlp_w = sd[pre + "MLP.weight"]
    S = mlp_w.shape[0]
    xv = x.reshape(B * N, h, w, C).permute(0, 3, 1, 2)
    tok = spa_embed(xv, mlp_w)                                             # [V, hw, S]
    pe = spa_embed(pe_hw.permute(2, 0, 1)[None], mlp_w)                    # [1, hw, S]  (LFT.py:180)
    tn = _ln(tok + pe, sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    w_in, w_out = sd[pre + "attention.in_proj_weight"], sd[pre + "attention.out_proj.weight"]
    if mode == "dense":
        mask = gen_mask_loop(h, w, 5).to(tok)                             # built on the host like LFT.py:147-162, then moved
        att = _mha(tn, tok, w_in, w_out, mask)
    else:
        hd = S // HEADS
        V_ = tok.shape[0]
        q = (tn @ w_in[:S].t()).view(V_, h * w, HEADS, hd)
        k = (tn @ w_in[S:2 * S].t()).view(V_, h * w, HEADS, hd)
        v = (tok @ w_in[2 * S:].t()).view(V_, h * w, HEADS, hd)
        idx, valid = (t.to(tok.device) for t in window_index(h, w, 5))
        kg = k[:, idx]                                                     # [V, hw, 25, H, hd]
        vg = v[:, idx]
        s = torch.einsum("vqhd,vqkhd->vqhk", q, kg) / math.sqrt(hd)
        s = s.masked_fill(~valid[None, :, None, :], float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = torch.einsum("vqhk,vqkhd->vqhd", p, vg).reshape(V_, h * w, S)
        att = o @ w_out.t()
    tok = att + tok
    f = _ln(tok, sd[pre + "feed_forward.0.weight"], sd[pre + "feed_forward.0.bias"])
    f = torch.relu(f @ sd[pre + "feed_forward.1.weight"].t()) @ sd[pre + "feed_forward.4.weight"].t()
    tok = f + tok
    out = tok @ sd[pre + "linear.0.weight"].view(C, S).t()                 # 1x1x1 conv 2C -> C
    return out.view(B, N, h, w, C)


def _cubic_w(t: torch.Tensor):
    """PyTorch cubic convolution coefficients, A = -0.75 (upsample_bicubic2d)."""
    a = -0.75
    def c1(x):  # |x| <= 1
        return ((a + 2) * x - (a + 3)) * x * x + 1
    def c2(x):  # 1 < |x| < 2
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return c2(t + 1), c1(t), c1(1 - t), c2(2 - t)


def bicubic_views(v: torch.Tensor, s: int) -> torch.Tensor:
    """interpolate() core (LFT.py:261): per-view bicubic, align_corners=False: src=(dst+0.5)/s-0.5,
    4 taps at floor(src)-1..+2 with indices clamped to the view, no output clamp. v:[V,h,w]->[V,hs,ws]."""
    V, h, w = v.shape
    def axis(n):
        dst = torch.arange(n * s, dtype=v.dtype, device=v.device)
        src = (dst + 0.5) / s - 0.5
        i0 = torch.floor(src)
        t = src - i0
        ws_ = torch.stack(_cubic_w(t), dim=1)                               # [n*s, 4]
        ii = (i0.long()[:, None] + torch.arange(-1, 3, device=v.device)[None]).clamp(0, n - 1)
        return ii, ws_
    iy, wy = axis(h)
    ix, wx = axis(w)
    rows = (v[:, iy, :] * wy[None, :, :, None]).sum(2)                      # [V, hs, w]
    out = (rows[:, :, ix] * wx[None, None]).sum(3)                          # [V, hs, ws]
    return out


def upsample_mosaic(feat: torch.Tensor, sd: Dict[str, torch.Tensor], A: int, s: int) -> torch.Tensor:
    """upsampling (LFT.py:39-44,79-80) on the view mosaic. feat: [B,N,h,w,C] -> [B,1,A*h*s,A*w*s].
    PixelShuffle: out channel c, sub-pixel (i,j) <- in channel c*s*s + i*s + j. The final 3x3 conv is
    zero-padded only at the mosaic border and reads the neighbouring view across interior borders."""
    B, N, h, w, C = feat.shape
    m = feat.view(B, A, A, h, w, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, A * h, A * w)
    y = F.conv2d(m, sd["upsampling.0.weight"])
    y = y.view(B, C, s, s, A * h, A * w).permute(0, 1, 4, 2, 5, 3).reshape(B, C, A * h * s, A * w * s)
    y = _lrelu(y)
    return F.conv2d(y, sd["upsampling.3.weight"], padding=1)


def forward(sd: Dict[str, torch.Tensor], lr: torch.Tensor, angRes: int, scale: int,
            mode: str = "window", dtype=torch.float32, stages: Optional[dict] = None) -> torch.Tensor:
    """get_model.forward (LFT.py:52-83). lr: [B,1,A*h,A*w] SAI mosaic -> [B,1,A*h*s,A*w*s].
    `stages`, if given, is filled with channels-last intermediates for stage-level tests."""
    A, s = angRes, scale
    sd = {k: v.to(device=lr.device, dtype=dtype) for k, v in sd.items()}
    lr = lr.to(dtype)
    B, _, H, W = lr.shape
    h, w = H // A, W // A
    C = sd["conv_init0.0.weight"].shape[0]
    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("altblock."))
    views = lr.view(B, A, h, A, w).permute(0, 1, 3, 2, 4).reshape(B * A * A, h, w)
    up = bicubic_views(views, s)                                            # LFT.py:54
    up = up.view(B, A, A, h * s, w * s).permute(0, 1, 3, 2, 4).reshape(B, 1, H * s, W * s)
    buf = conv_stack(views[:, None], sd)                                    # [V,C,h,w]
    buf = buf.view(B, A * A, C, h, w).permute(0, 1, 3, 4, 2).contiguous()   # [B,N,h,w,C]
    if stages is not None:
        stages["conv_init"] = buf.clone()
    pe_a = ang_position(A, C, dtype).to(lr.device)                          # LFT.py:103 (.to(device) per forward)
    pe_s = spa_position(h, w, C, dtype).to(lr.device)
    x = buf
    for i in range(layers):
        x = ang_trans(x, sd, f"altblock.{i}.ang_trans.", pe_a)
        if stages is not None:
            stages[f"ang{i}"] = x.clone()
        x = spa_trans(x, sd, f"altblock.{i}.spa_trans.", pe_s, mode)
        if stages is not None:
            stages[f"spa{i}"] = x.clone()
    x = x + buf                                                             # LFT.py:76
    out = upsample_mosaic(x, sd, A, s) + up                                 # LFT.py:79-81
    return out


# ----------------------------------------------------------------------------- patch tiler
def lf_divide(data: torch.Tensor, A: int, patch: int, stride: int) -> torch.Tensor:
    """LFdivide + ImageExtend (utils/utils.py:91-138), vectorised. data [A*h0, A*w0] ->
    [numU, numV, A*patch, A*patch]. Mirror-extend each view by bdr=(patch-stride)//2 (edge pixel
    repeated: flip-concat), zero-fill up to hE/wE, cut patches at `stride`."""
    uh, vw = data.shape
    h0, w0 = uh // A, vw // A
    bdr = (patch - stride) // 2
    h, w = h0 + 2 * bdr, w0 + 2 * bdr
    numU = (h - patch) // stride + (2 if (h - patch) % stride else 1)
    numV = (w - patch) // stride + (2 if (w - patch) % stride else 1)
    hE, wE = stride * (numU - 1) + patch, stride * (numV - 1) + patch
    v = data.view(A, h0, A, w0).permute(0, 2, 1, 3)                         # [A,A,h0,w0]
    def mirror(n, ext):  # index map of Im_Ext[n-bdr : 2n+bdr] (flip | id | flip)
        j = torch.arange(-bdr, n + bdr)
        return torch.where(j < 0, -j - 1, torch.where(j >= n, 2 * n - 1 - j, j))
    iy, ix = mirror(h0, bdr), mirror(w0, bdr)
    ext = torch.zeros(A, A, hE, wE, dtype=data.dtype)
    ext[:, :, :h, :w] = v[:, :, iy][:, :, :, ix]
    out = torch.zeros(numU, numV, A * patch, A * patch, dtype=data.dtype)
    for kh in range(numU):
        for kw in range(numV):
            p = ext[:, :, kh * stride:kh * stride + patch, kw * stride:kw * stride + patch]
            out[kh, kw] = p.permute(0, 2, 1, 3).reshape(A * patch, A * patch)
    return out


def lf_integrate(sub: torch.Tensor, A: int, pz: int, stride: int, h0: int, w0: int) -> torch.Tensor:
    """LFintegrate (utils/utils.py:141-157) for square patches: keep the central `stride` square of
    every SR patch view, tile, crop to [h0, w0]. sub [numU,numV,A*pz,A*pz] -> [A,A,h0,w0]."""
    numU, numV, pH, pW = sub.shape
    ph = pH // A
    bdr = (pz - stride) // 2
    s6 = sub.view(numU, numV, A, ph, A, ph)[:, :, :, bdr:bdr + stride, :, bdr:bdr + stride]
    full = s6.permute(2, 4, 0, 3, 1, 5).reshape(A, A, numU * stride, numV * stride)
    return full[:, :, :h0, :w0].contiguous()


def infer_light_field(sd, lr_sai: torch.Tensor, A: int, s: int, patch: int = 32, stride: int = 16,
                      mode: str = "dense", batch: int = 1, max_patches: Optional[int] = None):
    """test.py:83-101: LFdivide -> one net() call per patch (batch=1 as the reference does) ->
    LFintegrate -> SAI mosaic [A*h0*s, A*w0*s]. `max_patches` bounds the work for baseline timing
    (remaining patches stay zero); returns (sr_sai, patches_done)."""
    uh, vw = lr_sai.shape
    h0, w0 = uh // A, vw // A
    sub = lf_divide(lr_sai, A, patch, stride)
    numU, numV = sub.shape[:2]
    flat = sub.view(numU * numV, 1, A * patch, A * patch)
    out = torch.zeros(numU * numV, A * patch * s, A * patch * s, dtype=lr_sai.dtype)
    n = numU * numV if max_patches is None else min(max_patches, numU * numV)
    for i in range(0, n, batch):
        j = min(n, i + batch)
        out[i:j] = forward(sd, flat[i:j], A, s, mode=mode)[:, 0]
    sr4 = lf_integrate(out.view(numU, numV, A * patch * s, A * patch * s), A, patch * s, stride * s, h0 * s, w0 * s)
    sr = sr4.permute(0, 2, 1, 3).reshape(A * h0 * s, A * w0 * s)            # test.py:100-101
    return sr, n
