"""North-star target check (BASELINE.json: ">= 20x the reference's single-GPU eager-PyTorch LFT forward at 1 B200"):
times the oracle -- the eager-PyTorch restatement of the reference forward -- on the same GPU, on the
same light field, and compares it with the CUDA path through the C ABI.  SURVEY.md section 8(d) "(ii) Eager GPU".

When the unmodified reference is staged under baseline/_ref (oracle/stage_reference.py) two more arms run the REAL
`model/LFT.py::get_model` (nn.MultiheadAttention -> fused SDPA, its own gen_mask loop) on the GPU - `ref_b1` with test.py's
B = 1 loop and `ref_b8` batched by 8 - and the >= 20x assertion is made against them as well.

Three eager arms of the oracle port, all fp32 with torch defaults:
  b1_dense  : test.py:83-99 semantics -- one patch per call, dense masked 1024x1024 attention, mask rebuilt per call
  b8_dense  : the same forward on batches of 8 patches (most the dense score tensors allow comfortably)
  b8_window : the 5x5-window formulation batched by 8 (the most favourable eager formulation, not what the reference runs)
The measured numbers are written to gpurun_out/eager_baseline.json (copied to profiles/ by hand)."""
import json
import os
import time

import pytest
import torch

from lft_b200 import synth
from oracle import lft_oracle as O

pytestmark = pytest.mark.gpu

A, S, SEED = 5, 4, 2


def _time_eager(fn, n_warm, n_iter):
    for _ in range(n_warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()           # wall clock on purpose: the eager path is host-bound (python mask loop)
    for _ in range(n_iter):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n_iter


def test_speedup_vs_eager_gpu_forward():
    from lft_b200.engine import Engine
    from lft_b200.lightfield import LightFieldSR, num_patches

    sd = synth.synth_state_dict(A, S, SEED)
    lf = torch.from_numpy(synth.synth_light_field(A, 128, 128, SEED)).cuda()          # [A*128, A*128] LR mosaic
    nU, nV = num_patches(128, 128)
    n_patches = nU * nV
    assert n_patches == 64

    eng = Engine(A, S)
    eng.load_state_dict(sd)
    sr = LightFieldSR(eng)
    for _ in range(3):
        sr(lf)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        out = sr(lf)
    ev1.record()
    torch.cuda.synchronize()
    ours_ms_per_lf = ev0.elapsed_time(ev1) / 10
    ours_ms_per_patch = ours_ms_per_lf / n_patches

    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    patches = O.lf_divide(lf.cpu(), A, 32, 16).reshape(n_patches, 1, A * 32, A * 32).cuda()
    with torch.no_grad():
        b1 = _time_eager(lambda: O.forward(sd_gpu, patches[:1], A, S, mode="dense"), 2, 6)
        b8d = _time_eager(lambda: O.forward(sd_gpu, patches[:8], A, S, mode="dense"), 1, 3) / 8
        b8w = _time_eager(lambda: O.forward(sd_gpu, patches[:8], A, S, mode="window"), 1, 3) / 8
        # parity reference without TF32 convolutions (torch's cuDNN default on GPU; the timed arms above keep the
        # defaults because that is what a user of the reference gets -- and it alone costs ~2e-4 of accuracy)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ref = O.forward(sd_gpu, patches[:1], A, S, mode="dense")
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
    # same result, so the ratio compares like with like
    got = eng.forward(patches[:1])
    assert (got - ref).abs().max().item() < 1e-4

    from oracle import reference_arm as R
    real = {}
    if R.staged():
        net = R.make_net(sd, A, S, "cuda")
        with torch.no_grad():
            real["ref_b1_test_py_semantics"] = _time_eager(lambda: net(patches[:1]), 2, 6)
            real["ref_b8"] = _time_eager(lambda: net(patches[:8]), 1, 3) / 8
            tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            try:
                rref = net(patches[:1])
            finally:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        assert (got - rref).abs().max().item() < 1e-4      # the CUDA path against the REAL reference on the same GPU

    res = {
        "workload": "HCInew-shape 5x5x128x128 LR light field, 4x, 64 patches of 32x32",
        "ours_ms_per_lf": ours_ms_per_lf, "ours_ms_per_patch": ours_ms_per_patch,
        "eager_gpu_ms_per_patch": {"b1_dense_test_py_semantics": 1e3 * b1, "b8_dense": 1e3 * b8d, "b8_window": 1e3 * b8w},
        "speedup": {"b1_dense_test_py_semantics": 1e3 * b1 / ours_ms_per_patch, "b8_dense": 1e3 * b8d / ours_ms_per_patch,
                    "b8_window": 1e3 * b8w / ours_ms_per_patch},
        "note": "eager arms: oracle/lft_oracle.py on cuda:0, fp32, torch defaults; ours: LightFieldSR (divide+forward+integrate)",
    }
    if real:
        res["reference_gpu_ms_per_patch"] = {k: 1e3 * v for k, v in real.items()}
        res["speedup_vs_reference"] = {k: 1e3 * v / ours_ms_per_patch for k, v in real.items()}
        res["note"] += "; reference arms: the unmodified model/LFT.py get_model (baseline/_ref) on cuda:0"
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/eager_baseline.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    assert out.shape == (A * 512, A * 512)
    assert res["speedup"]["b1_dense_test_py_semantics"] >= 20.0
    assert res["speedup"]["b8_dense"] >= 20.0
    if real:   # the north-star denominator is the reference as test.py drives it (B = 1 per call); ref_b8 is recorded only
        assert res["speedup_vs_reference"]["ref_b1_test_py_semantics"] >= 20.0, res["speedup_vs_reference"]
