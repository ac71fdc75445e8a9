"""Which stage's bf16 rounding costs the output accuracy?  The forward is chained from the stage entry points with ONE stage
(or stage family) in bf16 mode and all others in fp32 mode; prints the output error (rms / max / correlation with the output)
per choice.  Decides where a mixed mode would have to keep three passes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine

A, s, h = 5, 4, 32
sd = synth.synth_state_dict(A, s, 3)
lr = torch.from_numpy(synth.synth_lr_mosaic(1, A, h, h, 5)).cuda()
e = Engine(A, s); e.load_state_dict(sd)


def chain(bf):
    def p(name):
        e.set_precision("bf16" if name in bf else "fp32")
    p("conv"); x = e.stage_conv_init(lr); res = x
    for i in range(4):
        p(f"ang{i}"); x = e.stage_ang(i, x)
        p(f"spa{i}"); x = e.stage_spa(i, x)
    p("up"); return e.stage_upsample(x + res, lr)[0, 0].double().cpu().numpy()


ref = chain(set())
full = e.forward(lr)[0, 0].double().cpu().numpy()
print("chained fp32 vs forward fp32:", np.abs(ref - full).max())
groups = {"all": {"conv", "up"} | {f"ang{i}" for i in range(4)} | {f"spa{i}" for i in range(4)}, "conv": {"conv"}, "up": {"up"},
          "ang*": {f"ang{i}" for i in range(4)}, "spa*": {f"spa{i}" for i in range(4)}}
for i in range(4):
    groups[f"ang{i}"] = {f"ang{i}"}
    groups[f"spa{i}"] = {f"spa{i}"}
for name, g in groups.items():
    out = chain(g)
    err = out - ref
    rc = ref - ref.mean()
    print(f"{name:6s} rms={err.std():.2e} mean={err.mean():+.2e} max={np.abs(err).max():.2e} gain={np.dot(err.ravel(), rc.ravel()) / np.dot(rc.ravel(), rc.ravel()):+.2e}", flush=True)
