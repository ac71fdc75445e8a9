"""The reference's own CPU / eager path for the benchmark's reference legs.  TEST / BENCH INFRASTRUCTURE (only bench.py's
`cpu_baseline` / `--impl reference` legs and tests/ import it), never the product.

If `baseline/_ref/` holds the staged reference sources (oracle/stage_reference.py) the UNMODIFIED reference runs:
`model/LFT.py::get_model` for the network and `utils/utils.py::LFdivide / LFintegrate` for the tiler, driven exactly like
test.py:83-101 (one net() call per patch, B = 1, `net.eval()` per call) - kind "reference".  Otherwise the oracle port
(oracle/lft_oracle.py, dense masked attention with the mask rebuilt per call, as the reference executes it) - kind "port"."""
from __future__ import annotations

import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
_ref = None


def staged() -> bool:
    return all(os.path.exists(os.path.join(STAGED, f)) for f in ("model/LFT.py", "utils/utils.py", "option.py"))


def load_reference():
    """(model module, utils module) of the staged reference, imported unmodified: utils.py:3,7 import skimage and parse
    sys.argv at import time, so a stub `skimage` and a clean argv are provided around the import (as in make_golden.py)."""
    global _ref
    if _ref is None:
        sys.path.insert(0, STAGED)
        argv, sys.argv = sys.argv, ["x"]
        try:
            if "skimage" not in sys.modules:
                sk = types.ModuleType("skimage")
                sk.metrics = types.ModuleType("skimage.metrics")
                sys.modules["skimage"] = sk
                sys.modules["skimage.metrics"] = sk.metrics
            import model.LFT as ref_model      # noqa
            import utils.utils as ref_utils    # noqa
        finally:
            sys.argv = argv
        _ref = (ref_model, ref_utils)
    return _ref


def make_net(sd, A: int, s: int, device="cpu"):
    ref_model, _ = load_reference()
    net = ref_model.get_model(types.SimpleNamespace(channels=64, angRes=A, scale_factor=s))
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def run_light_field(sd, lf: torch.Tensor, A: int, s: int, max_patches=None, device="cpu", net=None, batch: int = 1):
    """test.py:83-101 on `lf` [A*h0, A*w0] (CPU tensor): LFdivide -> net() per patch (B = `batch`, 1 as the reference does)
    -> LFintegrate -> SAI mosaic.  At most `max_patches` patches are evaluated (bounded sample; the rest of the SR light
    field stays zero).  Returns (sr_sai, patches_done, seconds, kind)."""
    if staged():
        _, U = load_reference()
        net = make_net(sd, A, s, device) if net is None else net
        h0, w0 = lf.shape[0] // A, lf.shape[1] // A
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        sub = U.LFdivide(lf, A, 32, 16)                                              # test.py:83
        nu, nv = sub.shape[:2]
        out = torch.zeros(nu, nv, A * 32 * s, A * 32 * s)                            # test.py:85
        flat_in, flat_out = sub.view(nu * nv, 1, A * 32, A * 32), out.view(nu * nv, A * 32 * s, A * 32 * s)
        n = nu * nv if max_patches is None else min(max_patches, nu * nv)
        for i in range(0, n, batch):
            j = min(n, i + batch)
            with torch.no_grad():
                net.eval()                                                           # test.py:92
                flat_out[i:j] = net(flat_in[i:j].to(device))[:, 0].cpu()             # test.py:94-95
        sr = U.LFintegrate(out, A, 32 * s, 16 * s, h0 * s, w0 * s)                   # test.py:96
        sai = sr.permute(0, 2, 1, 3).reshape(A * h0 * s, A * w0 * s)                 # test.py:100
        return sai, n, time.perf_counter() - t0, "reference"
    from oracle import lft_oracle as O
    t0 = time.perf_counter()
    with torch.no_grad():
        sai, n = O.infer_light_field(sd, lf, A, s, mode="dense", batch=batch, max_patches=max_patches)
    return sai, n, time.perf_counter() - t0, "port"
