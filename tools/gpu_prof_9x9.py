import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
A, s, B = 9, 4, 16
eng = Engine(A, s); eng.load_state_dict(synth.synth_state_dict(A, s, 4))
lr = torch.rand(B, 1, A * 32, A * 32, device="cuda")
for _ in range(2): eng.forward(lr, max_ws_bytes=8 << 30)
eng.profile_enable(True)
for _ in range(3): eng.forward(lr, max_ws_bytes=8 << 30)
torch.cuda.synchronize()
p = eng.profile_read()
tot = sum(v["ms"] for v in p.values())
for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
    if v["launches"]: print(f"{k:16s} {v['ms'] / 3:8.3f} ms  {100 * v['ms'] / tot:5.1f}%")
