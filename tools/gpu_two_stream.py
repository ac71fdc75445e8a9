"""Experiment: one light field as 64/n chunks of n patches issued alternately on two CUDA streams (own workspace
each), so that CTAs of different kernels (tensor-heavy / LSU-heavy) share the SMs. Compared with one stream."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
A, s = 5, 4
sd = synth.synth_state_dict(A, s, 0)
engs = [Engine(A, s) for _ in range(2)]
for e in engs: e.load_state_dict(sd)
lf = torch.from_numpy(synth.synth_light_field(A, 128, 128, 2)).cuda()
streams = [torch.cuda.Stream() for _ in range(2)]
out = torch.empty(64, A, A, 16 * s, 16 * s, device="cuda")
ref = engs[0].forward_lf_crops(lf, 0, 64).clone()

def run(n, nstreams):
    cur = torch.cuda.current_stream()
    for st in streams[:nstreams]: st.wait_stream(cur)
    for i, p0 in enumerate(range(0, 64, n)):
        k = i % nstreams
        with torch.cuda.stream(streams[k]):
            engs[k].forward_lf_crops(lf, p0, p0 + n, out=out[p0:p0 + n])
    for st in streams[:nstreams]: cur.wait_stream(st)

for n in (32, 16, 8, 4):
    for ns in (1, 2):
        for _ in range(3): run(n, ns)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run(n, ns)
        e1.record(); torch.cuda.synchronize()
        ok = torch.equal(out, ref)
        print(f"chunk {n:3d} patches, {ns} stream(s): {e0.elapsed_time(e1)/10:.3f} ms/LF  bit-identical={ok}")
