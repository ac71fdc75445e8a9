"""Smallest end-to-end invocation (for compute-sanitizer): one 8x8 patch forward + a tiny light field."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import LightFieldSR
from oracle import lft_oracle as O
A, s, h = 5, 4, 8
sd = synth.synth_state_dict(A, s, 0)
eng = Engine(A, s); eng.load_state_dict(sd)
lr = torch.from_numpy(synth.synth_lr_mosaic(1, A, h, h, 0))
out = eng.forward(lr.cuda()); torch.cuda.synchronize()
print("forward err", (out.cpu() - O.forward(sd, lr, A, s)).abs().max().item())
lf = torch.from_numpy(synth.synth_light_field(A, 32, 48, 1)).cuda()
sr = LightFieldSR(eng)(lf); torch.cuda.synchronize()
print("lf ok", tuple(sr.shape), bool(torch.isfinite(sr).all()))
