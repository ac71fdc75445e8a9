import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from oracle import lft_oracle as O
A, P = 5, 8
for s in (2, 4):
    sd = synth.synth_state_dict(A, s, 0)
    lr = torch.zeros(1, 1, A * P, A * P)
    feat = torch.ones(1, A * A, P, P, 64)
    w0 = torch.zeros(64 * s * s, 64, 1, 1)
    for c in range(64):
        for ij in range(s * s): w0[c * s * s + ij, c, 0, 0] = 1.0 + ij     # H[c] at sub-pixel ij = (1 + ij) * feat[c]
    w3 = torch.zeros(1, 64, 3, 3); w3.view(64, 9)[7, 4] = 1.0
    sd2 = dict(sd); sd2["upsampling.0.weight"] = w0; sd2["upsampling.3.weight"] = w3
    eng = Engine(A, s); eng.load_state_dict(sd2)
    out = eng.stage_upsample(feat.cuda(), lr.cuda()).cpu()[0, 0]
    ref = O.upsample_mosaic(feat, sd2, A, s)[0, 0]
    print("s =", s, "expected block:\n", ref[s * 3:s * 4, s * 3:s * 4], "\ngot:\n", out[s * 3:s * 4, s * 3:s * 4])
