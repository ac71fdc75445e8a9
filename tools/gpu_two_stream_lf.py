"""Tail filling at small per-GPU batches (the N = 8 strong-scaling regime: 8 patches per rank = 5.5 waves of 296 persistent CTAs
per kernel): the patches of one rank as ONE call vs k parts on k streams (k engines = k workspaces), whose kernels can fill
each other's last waves.  Prints ms per call for every split."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import patch_ranges

A, s, h0, w0 = 5, 4, 128, 128
sd = synth.synth_state_dict(A, s, 0)
K = 4
engs = [Engine(A, s) for _ in range(K)]
for e in engs: e.load_state_dict(sd)
lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, 2)).cuda()
sr = torch.empty(A * h0 * s, A * w0 * s, device="cuda")
streams = [torch.cuda.Stream() for _ in range(K)]


def run(n, k):
    if k == 1:
        engs[0].forward_lf_sr(lf, 0, n, sr)
        return
    main = torch.cuda.current_stream()
    for i, (a, b) in enumerate(patch_ranges(n, k)):
        streams[i].wait_stream(main)
        with torch.cuda.stream(streams[i]):
            engs[i].forward_lf_sr(lf, a, b, sr)
    for i in range(k):
        main.wait_stream(streams[i])


def timed(fn, reps=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n in (8, 16, 32, 64):
    run(n, 1); torch.cuda.synchronize(); ref = sr.clone()
    res = []
    for k in (1, 2, 3, 4):
        sr.zero_(); run(n, k); torch.cuda.synchronize()
        ok = bool(torch.equal(sr, ref))
        res.append(f"{k} stream(s) {timed(lambda: run(n, k)):.3f} ms{'' if ok else ' MISMATCH'}")
    print(f"{n} patches: " + " | ".join(res), flush=True)
