"""Deterministic synthetic checkpoints and light fields.

The reference's shipped weights (`pth/LFT_5x5_{2x,4x}_epoch_50_model.pth`) and datasets are not
available offline, so parity is tested with checkpoints in the *same format* (test.py:34-51,
train.py:95-103: ``{'epoch': int, 'state_dict': OrderedDict[str, fp32 tensor]}``) filled from a
counter-based generator that does not depend on torch's RNG stream (so the fixture generator, the
CPU tests and the GPU box all see bit-identical weights).

Key names / shapes follow ``model/LFT.py:9-50,118-214`` (78 tensors, 1,163,392 params at 4x).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np


def state_dict_spec(angRes: int = 5, scale: int = 4, channels: int = 64, layers: int = 4):
    """Ordered (key, shape, fan_in, kind) list of the reference state_dict (LFT.py:23-44,125-145,199-214)."""
    C = channels
    S = 2 * C
    spec = [("conv_init0.0.weight", (C, 1, 1, 3, 3), 9, "w")]
    for i in (0, 2, 4):
        spec.append((f"conv_init.{i}.weight", (C, C, 1, 3, 3), 9 * C, "w"))
    for i in range(layers):
        p = f"altblock.{i}.spa_trans."
        spec += [
            (p + "MLP.weight", (S, 9 * C), 9 * C, "w"),
            (p + "norm.weight", (S,), 0, "g"),
            (p + "norm.bias", (S,), 0, "b"),
            (p + "attention.in_proj_weight", (3 * S, S), S, "w"),
            (p + "attention.out_proj.weight", (S, S), S, "w"),
            (p + "feed_forward.0.weight", (S,), 0, "g"),
            (p + "feed_forward.0.bias", (S,), 0, "b"),
            (p + "feed_forward.1.weight", (2 * S, S), S, "w"),
            (p + "feed_forward.4.weight", (S, 2 * S), 2 * S, "w"),
            (p + "linear.0.weight", (C, S, 1, 1, 1), S, "w"),
        ]
        p = f"altblock.{i}.ang_trans."
        spec += [
            (p + "norm.weight", (C,), 0, "g"),
            (p + "norm.bias", (C,), 0, "b"),
            (p + "attention.in_proj_weight", (3 * C, C), C, "w"),
            (p + "attention.out_proj.weight", (C, C), C, "w"),
            (p + "feed_forward.0.weight", (C,), 0, "g"),
            (p + "feed_forward.0.bias", (C,), 0, "b"),
            (p + "feed_forward.1.weight", (2 * C, C), C, "w"),
            (p + "feed_forward.4.weight", (C, 2 * C), 2 * C, "w"),
        ]
    spec += [
        ("upsampling.0.weight", (C * scale * scale, C, 1, 1), C, "w"),
        ("upsampling.3.weight", (1, C, 3, 3), 9 * C, "w"),
    ]
    return spec


def _uniform(tag: str, seed: int, n: int) -> np.ndarray:
    """n doubles in [0,1) from Philox keyed by (crc32(tag), seed); uses only random_raw (stream-stable)."""
    key = (zlib.crc32(tag.encode()) << 32) | (seed & 0xFFFFFFFF)
    raw = np.random.Philox(key=key).random_raw(n)
    return (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_state_dict_np(angRes: int = 5, scale: int = 4, seed: int = 0, channels: int = 64,
                        gain: float = 1.0, qk_gain: float = 1.0, ln_wide: bool = False) -> "OrderedDict[str, np.ndarray]":
    """Weights ~ U(-b, b), b = gain/sqrt(fan_in) (torch's default conv/linear bound, and the
    kaiming_uniform_(a=sqrt(5)) the reference applies to in_proj_weight, LFT.py:132,204);
    LayerNorm gamma = 1 + 0.2(u-0.5), beta = 0.2(u-0.5) so the affine path is exercised.

    `qk_gain` multiplies the Wq and Wk rows (the first 2E rows of every `attention.in_proj_weight`, LFT.py:131,203) only:
    logits grow with qk_gain^2 while every other activation keeps its range, which is how a trained network's peaky
    soft-max is imitated (a global `gain` explodes through the four blocks instead).  `ln_wide` draws the LayerNorm gammas
    from [0.2, 3] and the betas from [-0.5, 0.5]."""
    sd = OrderedDict()
    for key, shape, fan_in, kind in state_dict_spec(angRes, scale, channels):
        n = int(np.prod(shape))
        u = _uniform(key, seed, n)
        if kind == "w":
            b = gain / np.sqrt(fan_in)
            v = (2.0 * u - 1.0) * b
            if qk_gain != 1.0 and key.endswith("attention.in_proj_weight"):
                v = v.reshape(shape)
                v[: 2 * shape[1]] *= qk_gain
        elif kind == "g":
            v = (0.2 + 2.8 * u) if ln_wide else 1.0 + 0.2 * (u - 0.5)
        else:
            v = (u - 0.5) if ln_wide else 0.2 * (u - 0.5)
        sd[key] = v.astype(np.float32).reshape(shape)
    return sd


def synth_state_dict(angRes: int = 5, scale: int = 4, seed: int = 0, channels: int = 64, gain: float = 1.0,
                     qk_gain: float = 1.0, ln_wide: bool = False):
    import torch
    return OrderedDict((k, torch.from_numpy(v.copy())) for k, v in
                       synth_state_dict_np(angRes, scale, seed, channels, gain, qk_gain, ln_wide).items())


def save_checkpoint(path: str, state_dict, epoch: int = 50, module_prefix: bool = False) -> None:
    """Write the reference checkpoint format (train.py:95-103); optional 'module.' prefix (test.py:39-43)."""
    import torch
    sd = OrderedDict((("module." + k) if module_prefix else k, v) for k, v in state_dict.items())
    torch.save({"epoch": epoch, "state_dict": sd}, path)


def synth_lr_mosaic(B: int, angRes: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """[B,1,A*h,A*w] fp32 in [0,1): smooth per-view pattern + noise."""
    A = angRes
    u = _uniform(f"lr{B}x{A}x{h}x{w}", seed, B * A * A * h * w).reshape(B, A, A, h, w)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    out = np.empty((B, A, A, h, w), np.float64)
    for b in range(B):
        for ua in range(A):
            for va in range(A):
                ph = 0.37 * b + 0.11 * ua
                smooth = 0.5 + 0.25 * np.sin(0.21 * (xx + 0.6 * va) + ph) * np.cos(0.17 * (yy + 0.6 * ua) - ph)
                out[b, ua, va] = 0.8 * smooth + 0.2 * u[b, ua, va]
    out = out.transpose(0, 1, 3, 2, 4).reshape(B, 1, A * h, A * w)
    return out.astype(np.float32)


def synth_light_field(angRes: int, h0: int, w0: int, seed: int = 0) -> np.ndarray:
    """One LR light field as the SAI mosaic [A*h0, A*w0] fp32 (test.py:77 `Lr_SAI_y`)."""
    return synth_lr_mosaic(1, angRes, h0, w0, seed)[0, 0]
