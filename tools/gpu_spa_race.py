"""Localise the rare nondeterminism of the spa stage: snapshot Q/K/V (embed), O (attn), tok=Y1 (ffn) and the stage
output after every run and compare with the first run."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
big = torch.full((2 * 1024 ** 3,), 123.0, device="cuda"); del big
A, s, P, B = 5, 4, 32, 12
R = int(sys.argv[1]) if len(sys.argv) > 1 else 400
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
sd = synth.synth_state_dict(A, s, 8)
eng = Engine(A, s, precision=prec); eng.load_state_dict(sd)
lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, P, P, 3)).cuda()
feat = eng.stage_conv_init(lr).clone()
T = B * A * A * P * P
def snap():
    out = eng.stage_spa(3, feat)
    ws = eng._workspace(B, P).view(torch.float32)
    parts = {"tokY1": ws[256 * T:384 * T], "q": ws[384 * T:512 * T], "k": ws[512 * T:640 * T], "v": ws[640 * T:768 * T],
             "o": ws[768 * T:896 * T]}
    d = {k: v.clone() for k, v in parts.items()}; d["out"] = out.clone()
    return d
ref = snap()
nbad = 0
for r in range(R):
    cur = snap()
    bad = {k: float((cur[k] - ref[k]).abs().max()) for k in ref if not torch.equal(cur[k], ref[k])}
    if bad:
        nbad += 1
        info = {}
        for k in bad:
            idx = (cur[k] != ref[k]).nonzero().flatten()
            info[k] = (bad[k], int(idx.numel()), int(idx.min()), int(idx.max()))
        print("run", r, info)
print("mismatching runs", nbad, "/", R)
