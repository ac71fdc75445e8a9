"""Full light-field inference = the reference's test loop (test.py:83-101): LFdivide -> net() per
patch -> LFintegrate, re-designed for B200:

  * all patches of a light field run as ONE batch (the forward is batch-independent),
  * LFdivide / LFintegrate are index arithmetic inside CUDA kernels (lft_divide / lft_integrate),
    and only the central 16s x 16s crop of every SR patch view is ever written,
  * LFintegrate is fused into the last kernel (lft_forward_lf_sr): every kept crop is stored at its final place in the
    assembled SR light field, no crop slab and no separate integrate pass,
  * multi-GPU: patches are independent units -> contiguous blocks of ceil(P/G) patches per rank, no
    data-path collective.  Re-assembly on rank 0 is either
      "direct"     : rank 0's SR buffer is mapped into every rank (CUDA IPC, `engine.PeerBuffer`) and the ranks' last
                     kernels store their crops straight into it over NVLink; the only collectives are two one-element
                     all-reduces that order the stores against rank 0's reads, or
      "collective" : the gather of the kept crops to rank 0 (`torch.distributed.gather`, NCCL over NVLink on the GPU box,
                     gloo in the CPU tests) followed by lft_integrate - the fallback when peer mapping is unavailable.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .engine import Engine

PATCH = 32    # option.py:16  patch_size_for_test
STRIDE = 16   # option.py:17  stride_for_test


def num_patches(h0: int, w0: int, patch: int = PATCH, stride: int = STRIDE) -> Tuple[int, int]:
    """numU, numV of LFdivide (utils/utils.py:95-104); Python floor division as in the reference, so a view smaller
    than a patch yields one zero-padded patch (the C side, `lft_lf_num_patches_ex`, must agree)."""
    bdr = (patch - stride) // 2
    h, w = h0 + 2 * bdr, w0 + 2 * bdr
    nu = (h - patch) // stride + (2 if (h - patch) % stride else 1)
    nv = (w - patch) // stride + (2 if (w - patch) % stride else 1)
    return nu, nv


def patch_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced blocks: the first n % world ranks get one extra patch
    (70 patches on 8 ranks -> 9,9,9,9,9,9,8,8)."""
    base, extra = divmod(n, world)
    out, p = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        out.append((p, p + c))
        p += c
    return out


def gather_crops(local: torch.Tensor, ranges: List[Tuple[int, int]], rank: int, world: int,
                 group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank crop slabs [n_r, A, A, c, c] to `dst` in patch order. Slabs are padded to the
    largest block so one fixed-size gather suffices (ragged blocks differ by at most one patch)."""
    if world == 1:
        return local
    nmax = max(b - a for a, b in ranges)
    shape = (nmax,) + tuple(local.shape[1:])
    send = local
    if local.shape[0] != nmax:
        send = torch.zeros(shape, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    send = send.contiguous()
    bufs = [torch.empty(shape, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: ranges[r][1] - ranges[r][0]] for r in range(world)], dim=0)


class LightFieldSR:
    """`sr = LightFieldSR(net_or_engine)(lr_sai)` with lr_sai [A*h0, A*w0] on the GPU ->
    sr_sai [A*h0*s, A*w0*s] (the `Sr_SAI_y` of test.py:100-101).

    `assemble` (multi-GPU only): "direct" (peer stores into rank 0's buffer), "collective" (gather + integrate) or "auto"
    (direct on an NCCL group if the peer mapping can be set up on every rank, else collective)."""

    def __init__(self, net_or_engine, max_ws_bytes: Optional[int] = None, patch: int = PATCH, stride: int = STRIDE,
                 assemble: str = "auto"):
        if assemble not in ("auto", "direct", "collective"):
            raise ValueError("assemble must be 'auto', 'direct' or 'collective'")
        self._src = net_or_engine
        self.max_ws_bytes = max_ws_bytes
        self.patch, self.stride = int(patch), int(stride)   # args.patch_size_for_test / args.stride_for_test
        self.assemble = assemble
        self._peer = {}        # (shape, world, rank) -> PeerBuffer or None (None: set-up failed, use the collective path)
        self._flag = None

    def _engine(self, device) -> Engine:
        if isinstance(self._src, Engine):
            return self._src
        return self._src.engine(device)

    def close(self) -> None:
        """Unmap / free the peer buffers (collective: call on every rank).  The peers unmap first, a barrier, then the owner
        frees - a mapping must not outlive the allocation it maps."""
        bufs = [b for b in self._peer.values() if b is not None]
        for b in bufs:
            if not b.owner:
                b.close()
        if bufs and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            torch.cuda.synchronize()
            dist.barrier()
        for b in bufs:
            if b.owner:
                b.close()
        self._peer = {}

    # -------------------------------------------------------------------------------------------- direct assembly
    def _peer_buffer(self, eng, shape, rank, world, group):
        """Rank 0 allocates the SR light field as a peer-visible buffer and broadcasts its IPC handle; every other rank maps
        it.  One all-reduce decides for all ranks whether the mapping stands (collective: every rank calls this)."""
        from .engine import PeerBuffer
        key = (tuple(shape), world, rank)
        if key in self._peer:
            return self._peer[key]
        dev = torch.device("cuda", eng.device)
        buf, ok = None, 1
        try:
            if rank == 0:
                buf = PeerBuffer.alloc(eng.device, shape)
        except Exception:
            ok = 0
        box = [buf.handle if buf is not None else None]
        dist.broadcast_object_list(box, src=0, group=group, device=dev)
        if rank != 0:
            try:
                if box[0] is None:
                    raise RuntimeError("no handle")
                buf = PeerBuffer.open(eng.device, box[0], shape)
            except Exception:
                ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        if int(t.item()) == 0:
            if buf is not None:
                buf.close()
            buf = None
        self._peer[key] = buf
        return buf

    def _fence(self, dev, group):
        """Stream-ordered cross-rank fence: a one-element all-reduce on the current stream (no host synchronisation)."""
        if self._flag is None or self._flag.device != dev:
            self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        dist.all_reduce(self._flag, group=group)

    @torch.no_grad()
    def __call__(self, lr_sai: torch.Tensor, rank: int = 0, world: int = 1, group=None) -> Optional[torch.Tensor]:
        eng = self._engine(lr_sai.device)
        A, s = eng.A, eng.s
        h0, w0 = lr_sai.shape[0] // A, lr_sai.shape[1] // A
        nu, nv = eng.num_patches(h0, w0, self.patch, self.stride)
        ranges = patch_ranges(nu * nv, world)
        p0, p1 = ranges[rank]
        shape = (A * h0 * s, A * w0 * s)
        lr_sai = lr_sai.contiguous()
        direct = hasattr(eng, "forward_lf_sr") and self.assemble != "collective"
        if world == 1 and direct:
            sr = torch.empty(shape, dtype=torch.float32, device=lr_sai.device)
            eng.forward_lf_sr(lr_sai, p0, p1, sr, max_ws_bytes=self.max_ws_bytes, patch=self.patch, stride=self.stride)
            return sr
        if world > 1 and direct and (self.assemble == "direct" or dist.get_backend(group) == "nccl"):
            buf = self._peer_buffer(eng, shape, rank, world, group)
            if buf is None and self.assemble == "direct":
                raise RuntimeError("assemble='direct': the peer mapping of rank 0's SR buffer could not be set up")
            if buf is not None:
                self._fence(lr_sai.device, group)   # rank 0 has consumed the previous light field: the buffer may be written
                eng.forward_lf_sr(lr_sai, p0, p1, buf, max_ws_bytes=self.max_ws_bytes, patch=self.patch, stride=self.stride)
                self._fence(lr_sai.device, group)   # every rank's stores have landed
                return buf.tensor.clone() if rank == 0 else None
        crops = eng.forward_lf_crops(lr_sai, p0, p1, max_ws_bytes=self.max_ws_bytes, patch=self.patch,
                                     stride=self.stride)
        allc = gather_crops(crops, ranges, rank, world, group)
        if allc is None:
            return None
        sr = torch.empty(shape, dtype=torch.float32, device=lr_sai.device)
        eng.integrate(allc, h0, w0, 0, nu * nv, sr, self.patch, self.stride)
        return sr


class HostPipeline:
    """Streams light fields pinned host -> device -> SR -> pinned host with the copies overlapped with compute:
    the H2D of light field i+1 and the D2H of light field i run on two side streams while the kernels of the other one
    run on the caller's stream (`depth` device slots).  What test.py's loop does per scene with `.to(device)` /
    `.cpu()` (test.py:94-95), without the stalls.

        pipe = HostPipeline(net)                       # net = lft_b200.model.get_model(args) or an Engine
        for lr_host, sr_host in zip(inputs, outputs):  # pinned CPU tensors [A*h0, A*w0] / [A*h0*s, A*w0*s]
            pipe.submit(lr_host, sr_host)
        pipe.drain()                                   # the current stream now waits for every output copy
    """

    def __init__(self, net_or_engine, depth: int = 2, max_ws_bytes: Optional[int] = None, patch: int = PATCH,
                 stride: int = STRIDE):
        # a ready LightFieldSR (e.g. one with a multi-GPU assembly mode chosen) is used as is
        self._sr = (net_or_engine if isinstance(net_or_engine, LightFieldSR)
                    else LightFieldSR(net_or_engine, max_ws_bytes, patch, stride))
        self._depth = depth
        self._slots: List[dict] = []
        self._copy: Optional[torch.cuda.Stream] = None   # device -> host
        self._up: Optional[torch.cuda.Stream] = None     # host -> device (own stream: an upload never queues behind a download)
        self._i = 0

    def _slot(self, device) -> dict:
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=device)
            self._up = torch.cuda.Stream(device=device)
            self._slots = [dict(lr=None, sr=None, h2d=torch.cuda.Event(), done=None, d2h=None) for _ in range(self._depth)]
        s = self._slots[self._i % self._depth]
        self._i += 1
        return s

    @torch.no_grad()
    def submit(self, lr_host: torch.Tensor, sr_host: Optional[torch.Tensor], device=None, rank: int = 0, world: int = 1,
               group=None) -> Optional[torch.cuda.Event]:
        """Multi-GPU (world > 1, collective: every rank submits the same light field from ITS pinned host copy): the patches
        are sharded over the ranks and only rank 0 receives the result (`sr_host` is ignored on the other ranks)."""
        need_out = rank == 0
        if lr_host.is_cuda or not lr_host.is_pinned() or (need_out and (sr_host is None or sr_host.is_cuda or not sr_host.is_pinned())):
            raise ValueError("HostPipeline.submit expects pinned host tensors")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        slot = self._slot(device)
        main = torch.cuda.current_stream(device)
        fresh = slot["lr"] is None or slot["lr"].shape != lr_host.shape
        if fresh:
            slot["lr"] = torch.empty(lr_host.shape, dtype=torch.float32, device=device)
        with torch.cuda.stream(self._up):
            if slot["done"] is not None:
                self._up.wait_event(slot["done"])       # the kernels that read this slot's LR buffer have finished
            if fresh:
                # a (re)allocated buffer comes from the caching allocator ON `main`: the block may have been released by
                # kernels of the OTHER slot that are still queued there, so the upload must wait for all of `main`
                self._up.wait_stream(main)
            slot["lr"].copy_(lr_host, non_blocking=True)
            slot["h2d"].record(self._up)
        main.wait_event(slot["h2d"])
        if slot["d2h"] is not None:
            main.wait_event(slot["d2h"])                # the slot's previous result has left the device
        sr = self._sr(slot["lr"], rank, world, group)
        slot["sr"] = sr                                  # keep alive until the slot is reused
        slot["done"] = torch.cuda.Event()
        slot["done"].record(main)
        if sr is None:                                   # not the assembling rank
            return None
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(slot["done"])
            sr_host.copy_(sr, non_blocking=True)
            slot["d2h"] = torch.cuda.Event()
            slot["d2h"].record(self._copy)
        return slot["d2h"]

    def drain(self) -> None:
        """Make the current stream wait for every outstanding device->host copy."""
        if self._copy is not None:
            torch.cuda.current_stream(self._copy.device).wait_stream(self._copy)
