"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference`) on the CPU.

Run in the authoring container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs and weights come from `lft_b200.synth` (counter-based, reproducible anywhere), so only
the reference OUTPUTS need to be committed. The reference is imported, never copied:
  * model/LFT.py   -> get_model(args).forward, plus forward hooks on conv_init / altblock.i.*
  * utils/utils.py -> LFdivide / LFintegrate (needs a stub `skimage` and a clean sys.argv,
                      because utils.py:3,7 import skimage and parse argv at import time)
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from lft_b200 import synth  # noqa: E402

REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)
    argv, sys.argv = sys.argv, ["x"]
    sk = types.ModuleType("skimage")
    sk.metrics = types.ModuleType("skimage.metrics")
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.metrics", sk.metrics)
    import model.LFT as ref_model          # noqa
    import utils.utils as ref_utils        # noqa
    sys.argv = argv
    return ref_model, ref_utils


def run_case(ref_model, name, A, s, h, B, seed, want_stages, qk_gain=1.0, ln_wide=False):
    args = types.SimpleNamespace(channels=64, angRes=A, scale_factor=s)
    net = ref_model.get_model(args)
    sd = synth.synth_state_dict(A, s, seed, qk_gain=qk_gain, ln_wide=ln_wide)
    net.load_state_dict(sd, strict=True)
    net.eval()
    lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, h, h, seed))
    stages = {}

    def cl(t):  # [B,C,N,h,w] -> channels-last [B,N,h,w,C]
        return t.permute(0, 2, 3, 4, 1).contiguous().numpy()

    hooks = []
    if want_stages:
        first = {}
        hooks.append(net.conv_init.register_forward_hook(
            lambda m, i, o: first.__setitem__("conv_init_in", i[0]) or first.__setitem__("conv_init_out", o)))
        for i in (0, 3):
            hooks.append(net.altblock[i].ang_trans.register_forward_hook(
                lambda m, inp, o, i=i: stages.__setitem__(f"ang{i}", cl(o.detach()))))
            hooks.append(net.altblock[i].spa_trans.register_forward_hook(
                lambda m, inp, o, i=i: stages.__setitem__(f"spa{i}", cl(o.detach()))))
    with torch.no_grad():
        out = net(lr)
    for hk in hooks:
        hk.remove()
    if want_stages:
        stages["conv_init"] = cl((first["conv_init_out"] + first["conv_init_in"]).detach())  # LFT.py:66
    path = os.path.join(HERE, f"{name}.npz")
    meta = [A, s, h, B, seed] + ([int(qk_gain), int(ln_wide)] if (qk_gain != 1.0 or ln_wide) else [])
    np.savez(path, out=out.numpy(), meta=np.array(meta), **stages)
    print(name, "out", tuple(out.shape), "absmax", float(out.abs().max()), "->", os.path.getsize(path) // 1024, "KiB")


def run_tiler(ref_utils, name, A, h0, w0, s, seed, store_full):
    lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, seed))
    sub = ref_utils.LFdivide(lf, A, 32, 16)
    numU, numV = sub.shape[:2]
    # stand-in SR patches: integer-valued so that integrate is checked exactly, independent of a net
    g = torch.arange(numU * numV * (A * 32 * s) ** 2, dtype=torch.float32).remainder(65521.0)
    fake_sr = g.view(numU, numV, A * 32 * s, A * 32 * s)
    integ = ref_utils.LFintegrate(fake_sr, A, 32 * s, 16 * s, h0 * s, w0 * s)
    rec = {
        "meta": np.array([A, h0, w0, s, seed, numU, numV]),
        "divide_sha256": np.frombuffer(hashlib.sha256(sub.numpy().tobytes()).digest(), dtype=np.uint8),
        "integrate_sha256": np.frombuffer(hashlib.sha256(integ.numpy().tobytes()).digest(), dtype=np.uint8),
    }
    if store_full:
        rec["divide"] = sub.numpy()
        rec["integrate"] = integ.numpy()
    # identity property noted in SURVEY section 4: integrate(divide(x)) == x at scale 1
    ident = ref_utils.LFintegrate(sub, A, 32, 16, h0, w0)
    rec["identity_ok"] = np.array([bool(torch.equal(ident.permute(0, 2, 1, 3).reshape(A * h0, A * w0), lf))])
    np.savez(os.path.join(HERE, f"{name}.npz"), **rec)
    print(name, "numU,numV", numU, numV, "identity", rec["identity_ok"][0])


# test.py:83,96 pass args.patch_size_for_test / args.stride_for_test (option.py:16-17) to the tiler: hash-only goldens
# for non-default values, incl. an odd patch-stride difference (LR border 5 but SR crop offset 11*s//2), no overlap
# (stride == patch) and a view smaller than one patch (h0 + 2*bdr < patch: one zero-padded patch row).
# Patches larger than 32 x 32 (--patch_size_for_test 48 / 64): the reference's PositionEncoding and gen_mask are
# size-generic (LFT.py:91-115,147-162); dense 4096 x 4096 attention limits the golden to a few views on the CPU.
BIG_PATCH_CASES = [  # name, A, s, h, B, seed, stages
    ("fwd_A2_s2_h64_B1", 2, 2, 64, 1, 21, False),
    ("fwd_A3_s4_h48_B1", 3, 4, 48, 1, 22, False),
]

TILER_PS_CASES = [  # name, A, h0, w0, s, seed, patch, stride
    ("tilerps_A3_40x56_s2_p32_s24", 3, 40, 56, 2, 5, 32, 24),
    ("tilerps_A3_40x56_s2_p32_s32", 3, 40, 56, 2, 5, 32, 32),
    ("tilerps_A5_44x60_s4_p16_s8", 5, 44, 60, 4, 7, 16, 8),
    ("tilerps_A3_40x56_s4_p32_s21", 3, 40, 56, 4, 8, 32, 21),
    ("tilerps_A3_12x50_s2_p32_s16", 3, 12, 50, 2, 9, 32, 16),
    ("tilerps_A2_33x47_s2_p24_s10", 2, 33, 47, 2, 10, 24, 10),
    ("tilerps_A3_80x96_s2_p64_s32", 3, 80, 96, 2, 11, 64, 32),
    ("tilerps_A3_80x96_s2_p64_s48", 3, 80, 96, 2, 11, 64, 48),
    ("tilerps_A5_128x128_s4_p64_s48", 5, 128, 128, 4, 2, 64, 48),
    ("tilerps_A5_108x156_s4_p48_s32", 5, 108, 156, 4, 3, 48, 32),
]


def run_tiler_ps(ref_utils, name, A, h0, w0, s, seed, patch, stride):
    lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, seed))
    sub = ref_utils.LFdivide(lf, A, patch, stride)
    numU, numV = sub.shape[:2]
    g = torch.arange(numU * numV * (A * patch * s) ** 2, dtype=torch.float32).remainder(65521.0)
    fake_sr = g.view(numU, numV, A * patch * s, A * patch * s)
    integ = ref_utils.LFintegrate(fake_sr, A, patch * s, stride * s, h0 * s, w0 * s)
    np.savez(os.path.join(HERE, f"{name}.npz"),
             meta=np.array([A, h0, w0, s, seed, numU, numV, patch, stride]),
             divide_sha256=np.frombuffer(hashlib.sha256(sub.numpy().tobytes()).digest(), dtype=np.uint8),
             integrate_sha256=np.frombuffer(hashlib.sha256(integ.numpy().tobytes()).digest(), dtype=np.uint8))
    print(name, "numU,numV", numU, numV)


# BASELINE configs 3 / 4 end to end through the reference's own test loop (test.py:83-101: LFdivide -> one net() call per
# patch -> LFintegrate -> SAI mosaic): the full 4x SR light field is 26 / 27 MB, so every 7th pixel of every 7th row is
# committed (7 is coprime to the 64-pixel crop pitch, the 128-pixel patch pitch and the view size: all patches, crop
# borders and views are sampled).
LF_CASES = [  # name, A, h0, w0, s, weight seed, light-field seed, qk_gain
    ("lf_A5_128x128_s4", 5, 128, 128, 4, 0, 2, 1.0),
    ("lf_A5_108x156_s4", 5, 108, 156, 4, 0, 3, 1.0),
]
LF_SUB = 7


def run_lf(ref_model, ref_utils, name, A, h0, w0, s, wseed, seed, qk_gain):
    import time
    args = types.SimpleNamespace(channels=64, angRes=A, scale_factor=s)
    net = ref_model.get_model(args)
    net.load_state_dict(synth.synth_state_dict(A, s, wseed, qk_gain=qk_gain), strict=True)
    net.eval()
    lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, seed))
    t0 = time.time()
    sub = ref_utils.LFdivide(lf, A, 32, 16)
    numU, numV = sub.shape[:2]
    out = torch.zeros(numU, numV, A * 32 * s, A * 32 * s)
    with torch.no_grad():
        for u in range(numU):
            for v in range(numV):
                out[u, v] = net(sub[u:u + 1, v:v + 1]).squeeze()          # test.py:88-95, B = 1 per call
    sr = ref_utils.LFintegrate(out, A, 32 * s, 16 * s, h0 * s, w0 * s)
    sai = sr.permute(0, 2, 1, 3).reshape(A * h0 * s, A * w0 * s)          # test.py:100
    np.savez(os.path.join(HERE, f"{name}.npz"), meta=np.array([A, h0, w0, s, wseed, seed, int(qk_gain), LF_SUB, numU, numV]),
             sub=sai[::LF_SUB, ::LF_SUB].contiguous().numpy(),
             sha256=np.frombuffer(hashlib.sha256(sai.numpy().tobytes()).digest(), dtype=np.uint8))
    print(name, tuple(sai.shape), "patches", numU * numV, f"{time.time() - t0:.0f} s")


# Sharpened attention (VERDICT r1 "weak" item 1): the default synthetic weights leave every soft-max near uniform (max
# probability 0.08).  qk_gain = 4 / 6 scales Wq and Wk only: max |logit| 22 / 50, max probability > 0.99, mean row
# maximum 0.37 / 0.6 in all eight attention calls; ln_wide draws LayerNorm gamma from [0.2, 3].
SHARP_CASES = [  # name, A, s, h, B, seed, stages, qk_gain, ln_wide
    ("fwd_sharp4_A5_s4_h8_B1", 5, 4, 8, 1, 13, True, 4.0, False),
    ("fwd_sharp6_A5_s4_h8_B1", 5, 4, 8, 1, 14, True, 6.0, False),
    ("fwd_sharp4_lnwide_A5_s2_h8_B1", 5, 2, 8, 1, 15, True, 4.0, True),
    ("fwd_sharp4_A3_s2_h12_B1", 3, 2, 12, 1, 16, False, 4.0, False),
    ("fwd_sharp4_A5_s4_h32_B1", 5, 4, 32, 1, 17, False, 4.0, False),
]


def main():
    torch.manual_seed(0)
    ref_model, ref_utils = import_reference()
    if "--big-patch-only" in sys.argv[1:]:
        for c in BIG_PATCH_CASES:
            run_case(ref_model, *c)
        for c in TILER_PS_CASES[6:]:
            run_tiler_ps(ref_utils, *c)
        return
    if "--lf-only" in sys.argv[1:]:         # the two full light fields (about 10 minutes of CPU time)
        for c in LF_CASES:
            run_lf(ref_model, ref_utils, *c)
        return
    if "--sharp-only" in sys.argv[1:]:      # add the sharpened-attention goldens without regenerating the others
        for c in SHARP_CASES:
            run_case(ref_model, *c)
        return
    if "--tiler-ps-only" in sys.argv[1:]:   # add the patch/stride goldens without regenerating the others
        for c in TILER_PS_CASES:
            run_tiler_ps(ref_utils, *c)
        return
    # name, A, s, h(=w), B, seed, stages
    run_case(ref_model, "fwd_A5_s4_h8_B1", 5, 4, 8, 1, 10, True)
    run_case(ref_model, "fwd_A5_s2_h8_B2", 5, 2, 8, 2, 11, False)
    run_case(ref_model, "fwd_A3_s2_h12_B1", 3, 2, 12, 1, 12, False)
    run_case(ref_model, "fwd_A5_s2_h32_B1", 5, 2, 32, 1, 1, False)
    run_case(ref_model, "fwd_A5_s4_h32_B1", 5, 4, 32, 1, 0, False)
    for c in SHARP_CASES:
        run_case(ref_model, *c)
    for c in LF_CASES:
        run_lf(ref_model, ref_utils, *c)
    for c in BIG_PATCH_CASES:
        run_case(ref_model, *c)
    run_tiler(ref_utils, "tiler_A3_40x56_s2", 3, 40, 56, 2, 5, True)
    run_tiler(ref_utils, "tiler_A5_108x156_s4", 5, 108, 156, 4, 3, False)
    run_tiler(ref_utils, "tiler_A5_128x128_s4", 5, 128, 128, 4, 2, False)
    for c in TILER_PS_CASES:
        run_tiler_ps(ref_utils, *c)


if __name__ == "__main__":
    main()
