"""Time one HCInew-shape light field with the 64-patch batch processed in chunks of n patches (workspace-limited)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import LightFieldSR
A, s = 5, 4
eng = Engine(A, s); eng.load_state_dict(synth.synth_state_dict(A, s, 0))
lf = torch.from_numpy(synth.synth_light_field(A, 128, 128, 2)).cuda()
for n in (64, 32, 16, 8, 4, 2):
    ws = eng.workspace_bytes(n, 32)
    sr = LightFieldSR(eng, max_ws_bytes=ws)
    for _ in range(3): sr(lf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): sr(lf)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {n:3d} patches: {e0.elapsed_time(e1)/10:.3f} ms/LF  (workspace {ws/2**20:.0f} MiB)")
