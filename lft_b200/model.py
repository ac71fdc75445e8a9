"""Drop-in for the reference's `model/LFT.py` plug-in surface (test.py:29-31, train.py:31-33):

    MODEL = importlib.import_module('model.' + args.model_name)   # e.g. a one-line model/LFT_b200.py:
    net = MODEL.get_model(args)                                    #   from lft_b200.model import *
    net.load_state_dict(checkpoint['state_dict'])                  # strict, unchanged key set
    out = net(lr)                                                  # [B,1,A*h,A*w] -> [B,1,A*h*s,A*w*s]

`get_model` is an nn.Module whose parameter tree reproduces the reference state_dict exactly (78 fp32
tensors, LFT.py:9-50,118-214) so the shipped `pth/LFT_5x5_{2x,4x}_epoch_50_model.pth` load with
`strict=True`, with or without the 'module.' prefix that test.py:39-43 tries first.  `forward` does no
arithmetic in PyTorch: it hands device pointers to liblft_b200.so (include/lft_b200.h).  Inference only
(north_star); there is no CPU fallback - a non-CUDA input raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn

from .engine import Engine
from . import capi

__all__ = ["get_model", "get_loss", "weights_init"]


class _P(nn.Module):
    """A parameter holder: `weight` (and optionally `bias`) of a fixed shape, zero-initialised;
    the values always come from a checkpoint."""

    def __init__(self, shape, bias_shape=None):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(*shape), requires_grad=False)
        if bias_shape is not None:
            self.bias = nn.Parameter(torch.zeros(*bias_shape), requires_grad=False)


class _Slot(nn.Module):
    """Parameter-free placeholder that keeps nn.Sequential indices aligned with the reference
    (LeakyReLU / ReLU / Dropout / PixelShuffle positions)."""


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.zeros(3 * dim, dim), requires_grad=False)
        self.out_proj = _P((dim, dim))


def _ffn(dim):
    # LayerNorm, Linear(dim,2dim), ReLU, Dropout, Linear(2dim,dim), Dropout  (LFT.py:136-143,207-214)
    return nn.Sequential(_P((dim,), (dim,)), _P((2 * dim, dim)), _Slot(), _Slot(), _P((dim, 2 * dim)), _Slot())


class _SpaTrans(nn.Module):
    def __init__(self, C):
        super().__init__()
        S = 2 * C
        self.MLP = _P((S, 9 * C))
        self.norm = _P((S,), (S,))
        self.attention = _Attention(S)
        self.feed_forward = _ffn(S)
        self.linear = nn.Sequential(_P((C, S, 1, 1, 1)))


class _AngTrans(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.norm = _P((C,), (C,))
        self.attention = _Attention(C)
        self.feed_forward = _ffn(C)


class _AltFilter(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.spa_trans = _SpaTrans(C)
        self.ang_trans = _AngTrans(C)


class get_model(nn.Module):
    """Same constructor contract as LFT.py:9-14: reads args.channels / args.angRes / args.scale_factor.
    Optional `args.precision` in {'fp32','bf16'} (default 'fp32': max-abs 1e-4 parity gate)."""

    def __init__(self, args):
        super().__init__()
        C = int(args.channels)
        self.channels = C
        self.angRes = int(args.angRes)
        self.factor = int(args.scale_factor)
        self.precision = str(getattr(args, "precision", "fp32"))
        if C != 64:
            raise ValueError("lft_b200 kernels are specialised for channels=64 (the shipped checkpoints)")
        self.conv_init0 = nn.Sequential(_P((C, 1, 1, 3, 3)))
        self.conv_init = nn.Sequential(_P((C, C, 1, 3, 3)), _Slot(), _P((C, C, 1, 3, 3)), _Slot(),
                                       _P((C, C, 1, 3, 3)), _Slot())
        self.altblock = nn.Sequential(*[_AltFilter(C) for _ in range(4)])
        self.upsampling = nn.Sequential(_P((C * self.factor ** 2, C, 1, 1)), _Slot(), _Slot(), _P((1, C, 3, 3)))
        self._engine: Optional[Engine] = None
        self._engine_key = None

    # -- weights ---------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state_dict.items())
        res = super().load_state_dict(sd, strict=strict, **kw)
        self._engine_key = None
        return res

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device: torch.device) -> Engine:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, self.precision, self._weights_key())
        if self._engine is None or self._engine.device != idx:
            self._engine = Engine(self.angRes, self.factor, self.channels, self.precision, idx)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.set_precision(self.precision)
            self._engine.load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, lr: torch.Tensor) -> torch.Tensor:
        if not lr.is_cuda:
            raise capi.LftError("lft_b200.get_model.forward needs a CUDA tensor: there is no CPU fallback")
        if lr.dim() != 4 or lr.shape[1] != 1:
            raise capi.LftError(f"expected lr of shape [B,1,A*h,A*w], got {tuple(lr.shape)}")
        x = lr.detach()
        if x.dtype != torch.float32:
            raise capi.LftError(f"expected float32 input (the reference forward is fp32-only), got {x.dtype}")
        return self.engine(x.device).forward(x.contiguous())


class get_loss(nn.Module):
    """LFT.py:269-278 (L1). Training is out of scope; kept so `MODEL.get_loss(args)` resolves."""

    def __init__(self, args=None):
        super().__init__()
        self.criterion_Loss = nn.L1Loss()

    def forward(self, SR, HR):
        return self.criterion_Loss(SR, HR)


def weights_init(m):  # LFT.py:281-283: a no-op in the reference
    pass
