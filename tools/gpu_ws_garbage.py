"""Does any kernel read workspace memory it did not write?  Run the same forwards with the workspace pre-filled
with zeros, with large finite values and with NaN: the outputs must be bit-identical."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
for (A, s, h0, w0, B, P) in [(5, 2, 40, 56, 3, 32), (5, 4, 128, 128, 2, 8), (3, 2, 33, 47, 2, 12)]:
    sd = synth.synth_state_dict(A, s, 8)
    eng = Engine(A, s); eng.load_state_dict(sd)
    lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, 20)).cuda()
    lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, P, P, 3)).cuda()
    nu, nv = eng.num_patches(h0, w0)
    res = {}
    for name, val in (("zero", 0.0), ("big", 1e3), ("nan", float("nan")), ("zero2", 0.0)):
        ws = eng._workspace(nu * nv, 32); ws[: ws.numel() // 4 * 4].view(torch.float32).fill_(val)
        c = eng.forward_lf_crops(lf, 0, nu * nv).clone()
        ws2 = eng._workspace(B, P); ws2[: ws2.numel() // 4 * 4].view(torch.float32).fill_(val)
        f = eng.forward(lr).clone()
        res[name] = (c, f)
    for name in ("big", "nan", "zero2"):
        dc = (res[name][0] - res["zero"][0]).abs(); df = (res[name][1] - res["zero"][1]).abs()
        print((A, s, h0, w0, B, P), name, "lf crops max diff", float(dc.nan_to_num(9e9).max()), "n", int((dc.nan_to_num(1) > 0).sum()),
              "| forward max diff", float(df.nan_to_num(9e9).max()), "n", int((df.nan_to_num(1) > 0).sum()))
