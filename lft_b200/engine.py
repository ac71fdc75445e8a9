"""Thin Python owner of one `lft_handle` (include/lft_b200.h). PyTorch supplies device memory and the
current CUDA stream; all arithmetic happens in liblft_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import torch

from . import capi


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class PeerBuffer:
    """A float32 device buffer other processes of the node can map (CUDA IPC through the C ABI: lft_peer_alloc / lft_peer_open).
    The owner creates it with `PeerBuffer.alloc`, sends `handle` (64 bytes) to its peers, they call `PeerBuffer.open`.
    `tensor` is a zero-copy torch view for the owner (and for peers, of the mapped memory)."""

    def __init__(self, device: int, ptr: int, shape, handle: bytes, owner: bool):
        self.device, self.ptr, self.shape, self.handle, self.owner = int(device), int(ptr), tuple(shape), handle, owner
        n = 1
        for d in self.shape:
            n *= int(d)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 2}
        self.tensor = torch.as_tensor(self, device=f"cuda:{self.device}") if (n and owner) else None  # peers only store

    @classmethod
    def alloc(cls, device: int, shape) -> "PeerBuffer":
        lib = capi.load()
        n = 4
        for d in shape:
            n *= int(d)
        ptr, h = C.c_void_p(), C.create_string_buffer(64)
        capi.check(lib.lft_peer_alloc(int(device), n, C.byref(ptr), h))
        return cls(device, ptr.value, shape, h.raw, True)

    @classmethod
    def open(cls, device: int, handle: bytes, shape) -> "PeerBuffer":
        lib = capi.load()
        ptr, h = C.c_void_p(), C.create_string_buffer(bytes(handle), 64)
        capi.check(lib.lft_peer_open(int(device), h, C.byref(ptr)))
        return cls(device, ptr.value, shape, bytes(handle), False)

    def close(self):
        if self.ptr:
            lib = capi.load()
            self.tensor = None
            (lib.lft_peer_free if self.owner else lib.lft_peer_close)(self.device, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, angRes: int, scale: int, channels: int = 64, precision: str = "fp32",
                 device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise capi.LftError("lft_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.A, self.s, self.C = int(angRes), int(scale), int(channels)
        self.device = torch.cuda.current_device() if device is None else int(device)
        cfg = capi.LftConfig(self.A, self.s, self.C, self._prec(precision), self.device)
        self._h = C.c_void_p()
        capi.check(self.lib.lft_create(C.byref(cfg), C.byref(self._h)))
        self._ws: Optional[torch.Tensor] = None
        self._graphs: dict = {}
        self.ready = False

    def _stream(self) -> C.c_void_p:
        """The current torch stream of THIS engine's device (not of whatever device happens to be current)."""
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _prec(p: str) -> int:
        if p in ("fp32", "float32"):
            return capi.PREC_FP32
        if p in ("bf16", "bfloat16"):
            return capi.PREC_BF16
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {p!r}")

    def close(self):
        self._graphs = {}
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.lft_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- weights
    def load_state_dict(self, sd: Mapping[str, torch.Tensor]) -> None:
        """Strict intake of the reference state_dict (78 fp32 tensors; a 'module.' prefix is stripped)."""
        for k, v in sd.items():
            key = k[7:] if k.startswith("module.") else k
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            capi.check(self.lib.lft_set_weight(self._h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
        self._graphs = {}   # captured graphs hold pointers to the weight slabs finalize is about to free
        capi.check(self.lib.lft_finalize_weights(self._h))
        self.ready = True

    def set_precision(self, precision: str) -> None:
        self._graphs = {}   # captured graphs replay the pass count they were captured with
        capi.check(self.lib.lft_set_precision(self._h, self._prec(precision)))

    # ---------------------------------------------------------------- workspace
    def workspace_bytes(self, B: int, P: int) -> int:
        n = C.c_size_t()
        capi.check(self.lib.lft_workspace_bytes(self._h, B, P, C.byref(n)))
        return int(n.value)

    def _workspace(self, B: int, P: int, max_bytes: Optional[int] = None) -> torch.Tensor:
        """The engine's single scratch buffer (grown on demand).  One workspace per engine: calls on one engine must be
        stream-ordered with each other (same stream, or explicit events) - the light-field pipeline does that."""
        need = self.workspace_bytes(B, P)
        if max_bytes is not None:
            need = min(need, max(max_bytes, self.workspace_bytes(1, P)))
        if self._ws is None or self._ws.numel() < need or self._ws.device.index != self.device:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def _check_in(self, t: torch.Tensor, name: str) -> None:
        if not t.is_cuda:
            raise capi.LftError(f"{name} must be a CUDA tensor (no CPU fallback)")
        if t.dtype != torch.float32:
            raise capi.LftError(f"{name} must be float32, got {t.dtype}")
        if not t.is_contiguous():
            raise capi.LftError(f"{name} must be contiguous")

    # ---------------------------------------------------------------- compute
    def forward(self, lr: torch.Tensor, max_ws_bytes: Optional[int] = None, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
        """get_model.forward: lr [B,1,A*P,A*P] -> [B,1,A*P*s,A*P*s] (LFT.py:52-83)."""
        self._check_in(lr, "lr")
        B, c, H, W = lr.shape
        if c != 1 or H != W or H % self.A:
            raise capi.LftError(f"lr must be [B,1,A*P,A*P] with square patches, got {tuple(lr.shape)}")
        P = H // self.A
        ws = self._workspace(B, P, max_ws_bytes) if ws is None else ws
        out = torch.empty(B, 1, H * self.s, W * self.s, dtype=torch.float32, device=lr.device)
        capi.check(self.lib.lft_forward(self._h, _ptr(lr), _ptr(out), B, P, _ptr(ws), ws.numel(), self._stream()))
        return out

    def forward_graphed(self, lr: torch.Tensor) -> torch.Tensor:
        """forward() replayed from a CUDA graph (one per input shape): the 21 launches of a forward cost one graph launch,
        which matters for the reference's own usage pattern - one 32x32 patch per call (test.py:88-95), where the kernels
        take ~15 us each.  The result lives in a buffer owned by the graph: it is overwritten by the next call with the
        same shape (clone it to keep it)."""
        self._check_in(lr, "lr")
        key = tuple(lr.shape)
        g = self._graphs
        if key not in g:
            B, _, H, _ = lr.shape
            # the graph owns everything its kernels point at: input, output and its OWN workspace (the engine's shared
            # workspace may be replaced when a later call needs a larger one); graphs are dropped by load_state_dict /
            # set_precision / close, which change or free what a captured launch refers to
            ws = torch.empty(self.workspace_bytes(B, H // self.A), dtype=torch.uint8, device=lr.device)
            static_in = torch.empty_like(lr)
            static_in.copy_(lr)
            self.forward(static_in, ws=ws)               # warm-up outside capture: builds the per-patch-size tables
            torch.cuda.synchronize(lr.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward(static_in, ws=ws)
            g[key] = (graph, static_in, static_out, ws)
        graph, static_in, static_out, _ = g[key]
        static_in.copy_(lr)
        graph.replay()
        return static_out

    def stage_conv_init(self, lr: torch.Tensor) -> torch.Tensor:
        self._check_in(lr, "lr")
        B, _, H, W = lr.shape
        P = H // self.A
        ws = self._workspace(B, P)
        out = torch.empty(B, self.A * self.A, P, P, 64, dtype=torch.float32, device=lr.device)
        capi.check(self.lib.lft_stage_conv_init(self._h, _ptr(lr), _ptr(out), B, P, _ptr(ws), ws.numel(), self._stream()))
        return out

    def _stage_tok(self, fn, layer: int, x: torch.Tensor) -> torch.Tensor:
        self._check_in(x, "feat")
        B, N, P, P2, Cc = x.shape
        assert N == self.A * self.A and P == P2 and Cc == 64
        ws = self._workspace(B, P)
        out = torch.empty_like(x)
        capi.check(fn(self._h, layer, _ptr(x), _ptr(out), B, P, _ptr(ws), ws.numel(), self._stream()))
        return out

    def stage_ang(self, layer: int, x: torch.Tensor) -> torch.Tensor:
        return self._stage_tok(self.lib.lft_stage_ang, layer, x)

    def stage_spa(self, layer: int, x: torch.Tensor) -> torch.Tensor:
        return self._stage_tok(self.lib.lft_stage_spa, layer, x)

    def stage_upsample(self, feat: torch.Tensor, lr: torch.Tensor) -> torch.Tensor:
        self._check_in(feat, "feat")
        self._check_in(lr, "lr")
        B, N, P, _, _ = feat.shape
        ws = self._workspace(B, P)
        H = self.A * P * self.s
        out = torch.empty(B, 1, H, H, dtype=torch.float32, device=lr.device)
        capi.check(self.lib.lft_stage_upsample(self._h, _ptr(feat), _ptr(lr), _ptr(out), B, P, _ptr(ws), ws.numel(),
                                               self._stream()))
        return out

    # ---------------------------------------------------------------- light-field path
    # patch / stride = test.py's --patch_size_for_test / --stride_for_test (option.py:16-17, defaults 32 / 16)
    def num_patches(self, h0: int, w0: int, patch: int = 32, stride: int = 16):
        nu, nv = C.c_int32(), C.c_int32()
        capi.check(self.lib.lft_lf_num_patches_ex(h0, w0, patch, stride, C.byref(nu), C.byref(nv)))
        return int(nu.value), int(nv.value)

    def divide(self, lr_lf: torch.Tensor, p0: int, p1: int, patch: int = 32, stride: int = 16) -> torch.Tensor:
        self._check_in(lr_lf, "lr_lf")
        h0, w0 = lr_lf.shape[0] // self.A, lr_lf.shape[1] // self.A
        out = torch.empty(p1 - p0, 1, self.A * patch, self.A * patch, dtype=torch.float32, device=lr_lf.device)
        capi.check(self.lib.lft_divide_ex(self._h, _ptr(lr_lf), h0, w0, patch, stride, p0, p1, _ptr(out), self._stream()))
        return out

    def forward_lf_crops(self, lr_lf: torch.Tensor, p0: int, p1: int, out: Optional[torch.Tensor] = None,
                         max_ws_bytes: Optional[int] = None, patch: int = 32, stride: int = 16) -> torch.Tensor:
        """LFdivide + forward for patches [p0,p1) -> kept crops [p1-p0, A, A, stride*s, stride*s]."""
        self._check_in(lr_lf, "lr_lf")
        h0, w0 = lr_lf.shape[0] // self.A, lr_lf.shape[1] // self.A
        n = p1 - p0
        c = stride * self.s
        if out is None:
            out = torch.empty(n, self.A, self.A, c, c, dtype=torch.float32, device=lr_lf.device)
        if n > 0:
            ws = self._workspace(n, patch, max_ws_bytes)
            capi.check(self.lib.lft_forward_lf_ex(self._h, _ptr(lr_lf), h0, w0, patch, stride, p0, p1, _ptr(out),
                                                  _ptr(ws), ws.numel(), self._stream()))
        return out

    def forward_lf_sr(self, lr_lf: torch.Tensor, p0: int, p1: int, sr_lf, max_ws_bytes: Optional[int] = None,
                      patch: int = 32, stride: int = 16) -> None:
        """LFdivide + forward + LFintegrate for patches [p0,p1): their kept crops are stored at their final place in the
        assembled SR light field `sr_lf` [A*h0*s, A*w0*s] - a CUDA tensor of this device, or a `PeerBuffer` mapping of a
        buffer that lives on another GPU of the node (stores over NVLink)."""
        self._check_in(lr_lf, "lr_lf")
        h0, w0 = lr_lf.shape[0] // self.A, lr_lf.shape[1] // self.A
        want = (self.A * h0 * self.s, self.A * w0 * self.s)
        if isinstance(sr_lf, PeerBuffer):
            if tuple(sr_lf.shape) != want:
                raise capi.LftError(f"sr_lf must be {want}, got {tuple(sr_lf.shape)}")
            dst = C.c_void_p(sr_lf.ptr)
        else:
            self._check_in(sr_lf, "sr_lf")
            if tuple(sr_lf.shape) != want:
                raise capi.LftError(f"sr_lf must be {want}, got {tuple(sr_lf.shape)}")
            dst = _ptr(sr_lf)
        if p1 > p0:
            ws = self._workspace(p1 - p0, patch, max_ws_bytes)
            capi.check(self.lib.lft_forward_lf_sr(self._h, _ptr(lr_lf), h0, w0, patch, stride, p0, p1, dst, _ptr(ws),
                                                  ws.numel(), self._stream()))

    def integrate(self, crops: torch.Tensor, h0: int, w0: int, p0: int, p1: int, sr_lf: torch.Tensor,
                  patch: int = 32, stride: int = 16) -> torch.Tensor:
        self._check_in(crops, "crops")
        self._check_in(sr_lf, "sr_lf")
        if p1 > p0:
            capi.check(self.lib.lft_integrate_ex(self._h, _ptr(crops), h0, w0, patch, stride, p0, p1, _ptr(sr_lf),
                                                 self._stream()))
        return sr_lf

    # ---------------------------------------------------------------- profiling
    def profile_enable(self, on: bool = True) -> None:
        capi.check(self.lib.lft_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        n = C.c_int32()
        names = (C.c_char_p * capi.PROFILE_MAX_KINDS)()
        launches = (C.c_int64 * capi.PROFILE_MAX_KINDS)()
        ms = (C.c_double * capi.PROFILE_MAX_KINDS)()
        units = (C.c_int64 * capi.PROFILE_MAX_KINDS)()
        capi.check(self.lib.lft_profile_read2(self._h, C.byref(n), names, launches, ms, units))
        return {names[i].decode(): {"launches": int(launches[i]), "ms": float(ms[i]), "units": int(units[i])}
                for i in range(n.value)}

    def launch_count(self) -> int:
        return int(self.lib.lft_launch_count(self._h))
