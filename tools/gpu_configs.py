"""Timings of the BASELINE.json configurations that bench.py does not cover (configs 1, 2 and 5); config 3 is bench.py,
config 4 is tools/gpu_lf_dist.py.  Prints a markdown table (copied to profiles/r01_configs.md)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine

WS = 8 << 30   # workspace cap: lft_forward chunks the batch to it (bit-identical results)

def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rows = []
def run(name, A, s, B, prec, reps, graphed=False):
    eng = Engine(A, s, precision=prec); eng.load_state_dict(synth.synth_state_dict(A, s, 4))
    lr = torch.rand(B, 1, A * 32, A * 32, device="cuda", generator=torch.Generator("cuda").manual_seed(B))
    ms = timed((lambda: eng.forward_graphed(lr)) if graphed else (lambda: eng.forward(lr, max_ws_bytes=WS)), reps)
    raw_mp = B * (A * 32 * s) ** 2 / 1e6
    flop = {(5, 4): 61.733e9, (5, 2): 58.853e9, (9, 4): 204.773e9}[(A, s)] * B
    rows.append(f"| {name} | {A}x{A} | {s}x | {B} | {prec} | {ms:.3f} | {B / ms * 1e3:.1f} | {raw_mp / ms * 1e3:.0f} | {flop / ms / 1e9:.0f} |")
    print(rows[-1], flush=True)
    eng.close()

print("| config | angRes | scale | patches B | path | ms / forward | patches/s | raw SR MP/s | algorithmic TFLOP/s |\n|---|---|---|---|---|---|---|---|---|")
run("1: one 32x32 patch", 5, 4, 1, "fp32", 50)
run("1: one 32x32 patch, CUDA graph replay", 5, 4, 1, "fp32", 50, graphed=True)
run("2: batch of 64", 5, 2, 64, "fp32", 10)
run("2: batch of 64", 5, 2, 64, "bf16", 10)
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    run("5: 9x9 batch sweep", 9, 4, B, "fp32", 3 if B >= 64 else 10)
