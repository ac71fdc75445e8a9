"""Multi-GPU check of the sharded light-field path (torchrun, NCCL): an EPFL-shape light field
(5x5 views of 108x156 -> 7x10 = 70 ragged patches) is split over the ranks, crops are gathered to rank 0 and
integrated; rank 0 compares with its own single-GPU result (must be bit-identical) and times both."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth
from lft_b200.engine import Engine
from lft_b200.lightfield import LightFieldSR, patch_ranges

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A, s, h0, w0 = 5, 4, 108, 156
eng = Engine(A, s, device=local)
eng.load_state_dict(synth.synth_state_dict(A, s, 0))
lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, 3)).cuda()
pipe = LightFieldSR(eng)
for _ in range(2):
    sr = pipe(lf, rank, world)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sr = pipe(lf, rank, world)
e1.record(); dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    one = pipe(lf)  # all 70 patches on this GPU
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        one = pipe(lf)
    f1.record(); torch.cuda.synchronize()
    mp = A * A * h0 * s * w0 * s / 1e6
    print(f"EPFL-shape 70 patches, ranges {patch_ranges(70, world)}: identical={bool(torch.equal(sr, one))} "
          f"finite={bool(torch.isfinite(sr).all())} shape={tuple(sr.shape)} "
          f"sharded {t.item():.2f} ms ({mp / t.item() * 1e3:.1f} MP/s) vs 1 GPU {f0.elapsed_time(f1) / 5:.2f} ms "
          f"({mp / (f0.elapsed_time(f1) / 5) * 1e3:.1f} MP/s)", flush=True)
dist.destroy_process_group()
