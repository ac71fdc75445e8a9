"""Print the phase timeline (cycles) of the middle CTA of the last k_spa_ffn launch (needs LFT_TIMELINE build)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import capi, synth
from lft_b200.engine import Engine
A, s = 5, 4
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
eng = Engine(A, s, precision=prec); eng.load_state_dict(synth.synth_state_dict(A, s, 0))
lr = torch.from_numpy(synth.synth_lr_mosaic(64, A, 32, 32, 0)).cuda()
for _ in range(2): eng.forward(lr)
torch.cuda.synchronize()
for which, name in ((0, "k_spa_ffn"), (1, "k_ang"), (2, "k_spa_embed_qkv")):
    buf = (C.c_int64 * 64)()
    capi.check(eng.lib.lft_debug_timeline(which, buf))
    row = [buf[i] for i in range(32)]; mma = [buf[32 + i] for i in range(32)]
    if not any(x > 0 for x in row + mma):
        print(prec, name, ": no marks in this build"); continue
    t0 = min(x for x in row + mma if x > 0)
    print(prec, name, "row :", {i: x - t0 for i, x in enumerate(row) if x > 0})
    print(prec, name, "mma :", {i: x - t0 for i, x in enumerate(mma) if x > 0})

buf = (C.c_int64 * 120)()
capi.check(eng.lib.lft_debug_timeline(4, buf))
v = [buf[i] for i in range(120)]
t0 = v[0]
if t0: print("embed producer (middle CTA): per slab [before empty-wait, after wait, after copy issue] relative cycles")
for i in range(26):
    a, b, c = v[3 * i:3 * i + 3]
    if a: print(f"  slab {i:2d}: {a - t0:7d} {b - t0:7d} {c - t0:7d}   wait {b - a:6d}  issue {c - b:5d}")
