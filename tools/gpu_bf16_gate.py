"""bf16-mode accuracy at the PSNR gate (north_star: |PSNR(ref,HR) - PSNR(bf16,HR)| <= 0.01 dB): prints, for a few weight
seeds / scales, the delta, PSNR(bf16, fp32 oracle), max-abs and the correlation of the bf16 error with (ref - HR)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from lft_b200 import synth
from lft_b200.engine import Engine
from oracle import lft_oracle as O
from test_gpu_parity import _hr_and_bicubic_lr, _psnr_views, _psnr

A, h = 5, 32
for (s, qk, seed) in [(4, 1.0, 3), (2, 1.0, 3), (4, 4.0, 3), (4, 1.0, 0), (4, 1.0, 7)]:
    sd = synth.synth_state_dict(A, s, seed, qk_gain=qk)
    hr, lr = _hr_and_bicubic_lr(A, h, s, 31)
    ref = O.forward(sd, lr, A, s)[0, 0].numpy()
    e = Engine(A, s, precision="bf16"); e.load_state_dict(sd)
    out = e.forward(lr.cuda())[0, 0].cpu().numpy()
    f = Engine(A, s, precision="fp32"); f.load_state_dict(sd)
    o32 = f.forward(lr.cuda())[0, 0].cpu().numpy()
    d, err = (ref - hr.numpy()).ravel().astype(np.float64), (out - ref).ravel().astype(np.float64)
    print(f"s={s} qk={qk} seed={seed}: dPSNR={abs(_psnr_views(ref, hr.numpy(), A) - _psnr_views(out, hr.numpy(), A)):.5f} dB "
          f"PSNR(ref,HR)={_psnr_views(ref, hr.numpy(), A):.2f} PSNR(bf16,ref)={_psnr(out, ref):.2f} maxabs={np.abs(out - ref).max():.2e} "
          f"err_rms={err.std():.2e} err_mean={err.mean():.2e} corr(err, ref-HR)={np.corrcoef(d, err)[0, 1]:.3f} "
          f"gain={np.dot(err, ref.ravel() - ref.mean()) / np.dot(ref.ravel() - ref.mean(), ref.ravel() - ref.mean()):.2e} "
          f"fp32 maxabs={np.abs(o32 - ref).max():.2e}", flush=True)
