"""Worker of tests/test_gpu_parity.py::test_multi_gpu_sharded_light_field_bit_identical (launched with torchrun, one rank
per GPU, NCCL): an HCInew-shape (64 patches) and a ragged EPFL-shape (70 patches) 4x light field are sharded patch-wise
over the ranks (BASELINE configs 3 / 4) and re-assembled on rank 0 by (a) direct peer stores into rank 0's buffer and
(b) the NCCL gather + lft_integrate; both must equal rank 0's own single-GPU result bit for bit, twice in a row (the second
pass reuses the peer buffer).  Prints DIST_LF_OK on success."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import synth  # noqa: E402
from lft_b200.engine import Engine  # noqa: E402
from lft_b200.lightfield import LightFieldSR, patch_ranges  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    A, s = 5, 4
    eng = Engine(A, s, device=local)
    eng.load_state_dict(synth.synth_state_dict(A, s, 0))
    ok = True
    for (h0, w0, seed) in ((128, 128, 2), (108, 156, 3)):
        lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, seed)).to(dev)
        one = LightFieldSR(eng)(lf) if rank == 0 else None            # all patches on this GPU
        for mode in ("direct", "collective"):
            pipe = LightFieldSR(eng, assemble=mode)
            for rep in range(2):
                sr = pipe(lf, rank, world)
                torch.cuda.synchronize()
                if rank == 0:
                    same = bool(torch.equal(sr, one))
                    print(f"{h0}x{w0} world={world} ranges={patch_ranges(eng.num_patches(h0, w0)[0] * eng.num_patches(h0, w0)[1], world)} "
                          f"{mode} pass {rep}: identical={same}", flush=True)
                    ok = ok and same
                else:
                    ok = ok and sr is None
            pipe.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0:
        print("DIST_LF_OK" if int(flag.item()) == 1 else "DIST_LF_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
