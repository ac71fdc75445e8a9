"""Full light-field inference at test.py's --patch_size_for_test / --stride_for_test settings (option.py:16-17): the
reference default (32, 16) keeps the central 16x16 of every 32x32 patch, i.e. computes 4x the pixels it keeps; larger
strides trade border context for throughput; larger patches (48, 64: SURVEY 8f-3) keep the border and cut the redundancy.
Prints a markdown table (copied to profiles/r02_patch_stride.md):
patches per light field, ms, integrated SR MP/s, and the PSNR of each setting's output against the default's
(the outputs differ because each patch sees less context - this is the reference's own behaviour with the same flags,
checked against the oracle in tests/test_gpu_parity.py::test_full_light_field_patch_stride_vs_oracle)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import LightFieldSR

A, s, h0, w0 = 5, 4, 128, 128
eng = Engine(A, s, precision=sys.argv[1] if len(sys.argv) > 1 else "fp32")
eng.load_state_dict(synth.synth_state_dict(A, s, 0))
lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, 2)).cuda()
mp = A * A * h0 * s * w0 * s / 1e6


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = None
print("| patch | stride | border | patches / LF | ms / LF | integrated SR MP/s | PSNR vs (32,16) output [dB] |\n|---|---|---|---|---|---|---|")
for patch, stride in ((32, 16), (32, 20), (32, 24), (32, 28), (32, 32), (16, 8), (24, 16), (48, 32), (64, 32), (64, 40), (64, 48), (64, 56), (64, 64)):
    sr = LightFieldSR(eng, patch=patch, stride=stride)
    out = sr(lf)
    nu, nv = eng.num_patches(h0, w0, patch, stride)
    ms = timed(lambda: sr(lf))
    if base is None:
        base = out.clone()
        ps = "-"
    else:
        mse = float(((out - base).double() ** 2).mean())
        ps = f"{10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item():.1f}"
    print(f"| {patch} | {stride} | {(patch - stride) // 2} | {nu * nv} | {ms:.2f} | {mp / ms * 1e3:.0f} | {ps} |", flush=True)
