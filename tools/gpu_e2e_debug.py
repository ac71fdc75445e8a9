"""Per-step timing of the end-to-end path (HostPipeline) to look at its variance."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import HostPipeline
A, S = 5, 4
eng = Engine(A, S); eng.load_state_dict(synth.synth_state_dict(A, S, 0))
lf_host = torch.from_numpy(synth.synth_light_field(A, 128, 128, 2)).pin_memory()
outs = [torch.empty(A * 512, A * 512).pin_memory() for _ in range(2)]
print("pinned:", lf_host.is_pinned(), [o.is_pinned() for o in outs])
pipe = HostPipeline(eng)
for trial in range(4):
    for i in range(2): pipe.submit(lf_host, outs[i & 1])
    pipe.drain(); torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(10):
        pipe.submit(lf_host, outs[i & 1]); evs[i + 1].record()
    t_cpu = time.perf_counter() - t0
    pipe.drain(); end = torch.cuda.Event(enable_timing=True); end.record(); torch.cuda.synchronize()
    steps = [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    print(f"trial {trial}: total {evs[0].elapsed_time(end) / 10:.2f} ms/step, cpu enqueue {1e3 * t_cpu / 10:.2f} ms/step, steps {[round(x, 1) for x in steps]}")
    # plain D2H bandwidth of the 26 MB result
    x = torch.empty(A * 512, A * 512, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); outs[0].copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print(f"   D2H 26 MB: {e0.elapsed_time(e1):.2f} ms")
