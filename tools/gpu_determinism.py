"""Repeat every stage / the whole forward on the same inputs (warm caching allocator, no implicit syncs) and count
bit-level mismatches against the first run."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
big = torch.full((2 * 1024 ** 3,), 123.0, device="cuda"); del big
A, s = 5, int(sys.argv[1]) if len(sys.argv) > 1 else 2
P, B, R = 32, 12, int(sys.argv[2]) if len(sys.argv) > 2 else 30
sd = synth.synth_state_dict(A, s, 8)
eng = Engine(A, s); eng.load_state_dict(sd)
lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, P, P, 3)).cuda()
def rep(name, fn):
    ref = fn().clone(); bad = 0; mx = 0.0
    for _ in range(R):
        o = fn()
        d = (o - ref).abs().max().item()
        if d != 0: bad += 1; mx = max(mx, d)
    print(f"{name:12s} mismatching runs {bad}/{R}  max diff {mx:.3e}")
feat = eng.stage_conv_init(lr).clone()
rep("conv_init", lambda: eng.stage_conv_init(lr))
for i in range(4):
    rep(f"ang{i}", lambda: eng.stage_ang(i, feat))
    rep(f"spa{i}", lambda: eng.stage_spa(i, feat))
rep("upsample", lambda: eng.stage_upsample(feat, lr))
rep("forward", lambda: eng.forward(lr))
lf = torch.from_numpy(synth.synth_light_field(A, 40, 56, 20)).cuda()
rep("lf_crops", lambda: eng.forward_lf_crops(lf, 0, 12))
