"""GPU bring-up: tcgen05 GEMM self-test (descriptor variants) + stage parity vs the oracle.
Usage (on the GPU box): python tools/gpu_bringup.py [stages...]   -> prints max-abs errors."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lft_b200 import capi, synth            # noqa: E402
from lft_b200.engine import Engine          # noqa: E402
from oracle import lft_oracle as O          # noqa: E402


def selftest():
    lib = capi.load()
    rng = np.random.default_rng(0)
    for (M, N, K) in [(128, 64, 64), (256, 128, 128), (128, 16, 64), (128, 256, 128), (128, 192, 64)]:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        ref = A.astype(np.float64) @ W.astype(np.float64).T
        for prec in (0, 1):
            for variant in (0, 2):
                D = np.zeros((M, N), np.float32)
                aux = np.zeros((M, 16), np.float32)
                rc = lib.lft_gemm_selftest(A.ctypes.data, W.ctypes.data, D.ctypes.data, aux.ctypes.data, M, N, K, prec, variant)
                if rc != 0:
                    print("selftest rc", rc, lib.lft_last_error()); continue
                err = np.abs(D - ref).max()
                auxerr = np.abs(aux - (2 * A[:, :16] + 1)).max()
                print(f"selftest M{M} N{N} K{K} prec{prec} variant{variant}: max|D-ref|={err:.3e} (ref absmax {np.abs(ref).max():.2f}) aux_err={auxerr:.1e}", flush=True)


def stages(A=5, s=4, h=8, B=1, seed=10, prec="fp32"):
    sd = synth.synth_state_dict(A, s, seed)
    lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, h, h, seed))
    st = {}
    ref = O.forward(sd, lr, A, s, stages=st)
    eng = Engine(A, s, precision=prec)
    eng.load_state_dict(sd)
    lrd = lr.cuda()
    def rep(name, got, want):
        e = (got.cpu() - want).abs().max().item()
        print(f"[{prec} A{A} s{s} h{h} B{B}] {name}: max-abs err {e:.3e} (absmax {want.abs().max().item():.3f})", flush=True)
    t0 = time.time()
    try:
        x = eng.stage_conv_init(lrd); torch.cuda.synchronize(); rep("conv_init", x, st["conv_init"])
        xin = st["conv_init"].cuda()
        y = eng.stage_ang(0, xin); torch.cuda.synchronize(); rep("ang0", y, st["ang0"])
        z = eng.stage_spa(0, st["ang0"].cuda()); torch.cuda.synchronize(); rep("spa0", z, st["spa0"])
        feat = (st["spa3"] + st["conv_init"]).cuda()
        up = eng.stage_upsample(feat, lrd); torch.cuda.synchronize(); rep("upsample", up, ref)
        out = eng.forward(lrd); torch.cuda.synchronize(); rep("forward", out, ref)
    except Exception as e:  # keep going so one call reports as much as possible
        print("STAGE FAILURE:", repr(e), flush=True)
    print("elapsed", time.time() - t0)


if __name__ == "__main__":
    what = sys.argv[1:] or ["selftest", "stages"]
    print(torch.cuda.get_device_name(0))
    if "selftest" in what:
        selftest()
    if "stages" in what:
        stages()
        stages(prec="bf16")
    if "big" in what:
        stages(A=5, s=4, h=32, B=2, seed=0)
        stages(A=5, s=2, h=32, B=1, seed=1)
