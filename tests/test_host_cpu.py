"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the drop-in
module reproduces the reference state_dict contract, the product refuses to run without a GPU, and
the sharding/gather host logic (world_size 2, gloo)."""
import ctypes as C
import os
import re
import types

import numpy as np
import pytest
import torch

from lft_b200 import capi, synth
from lft_b200 import lightfield as LF
from oracle import lft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "lft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lft_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lft_b200.h but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(capi.SIGNATURES) == set(names)
    assert lib.lft_version() >= 100


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    cfg = capi.LftConfig(5, 4, 64, 0, 0)
    h = C.c_void_p()
    rc = lib.lft_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.lft_last_error()
    from lft_b200.engine import Engine
    with pytest.raises(capi.LftError):
        Engine(5, 4)


def test_bad_config_rejected():
    lib = capi.load()
    h = C.c_void_p()
    for cfg in (capi.LftConfig(5, 3, 64, 0, 0), capi.LftConfig(5, 4, 32, 0, 0), capi.LftConfig(12, 4, 64, 0, 0)):
        assert lib.lft_create(C.byref(cfg), C.byref(h)) == -1


@pytest.mark.parametrize("A,s", [(5, 4), (5, 2), (9, 4)])
def test_dropin_state_dict_contract(A, s, tmp_path):
    from lft_b200.model import get_model
    net = get_model(types.SimpleNamespace(channels=64, angRes=A, scale_factor=s))
    sd = synth.synth_state_dict(A, s, 3)
    assert list(net.state_dict().keys()) == list(sd.keys())
    assert [tuple(v.shape) for v in net.state_dict().values()] == [tuple(v.shape) for v in sd.values()]
    for prefix in (False, True):  # test.py:39-51 tries 'module.'-prefixed keys first
        p = tmp_path / f"ck{int(prefix)}.pth"
        synth.save_checkpoint(str(p), sd, module_prefix=prefix)
        ck = torch.load(str(p), map_location="cpu")
        net.load_state_dict(ck["state_dict"])  # strict
        assert torch.equal(net.state_dict()["altblock.2.spa_trans.MLP.weight"], sd["altblock.2.spa_trans.MLP.weight"])
    bad = dict(sd)
    bad.pop("upsampling.3.weight")
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad)


def test_forward_refuses_cpu_tensor():
    from lft_b200.model import get_model
    net = get_model(types.SimpleNamespace(channels=64, angRes=5, scale_factor=4))
    with pytest.raises(capi.LftError):
        net(torch.zeros(1, 1, 40, 40))


def test_patch_partition_and_counts():
    assert LF.num_patches(128, 128) == (8, 8)
    assert LF.num_patches(108, 156) == (7, 10)
    lf = torch.zeros(5 * 44, 5 * 60)
    assert tuple(O.lf_divide(lf, 5, 32, 16).shape[:2]) == LF.num_patches(44, 60)
    r = LF.patch_ranges(70, 8)
    assert [b - a for a, b in r] == [9, 9, 9, 9, 9, 9, 8, 8] and r[0][0] == 0 and r[-1][1] == 70
    assert all(r[i][1] == r[i + 1][0] for i in range(7))
    assert LF.patch_ranges(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]


def test_c_tiling_counts_match_reference_formula():
    """lft_lf_num_patches_ex (C, truncating division) against LFdivide's count formula (Python floor division,
    utils.py:93-104) over a sweep of view sizes / patch sizes / strides, incl. views smaller than one patch."""
    lib = capi.load()
    nu, nv = C.c_int32(), C.c_int32()
    checked = 0
    for patch in (4, 8, 16, 24, 31, 32, 48, 64):
        for stride in sorted({1, 2, patch // 2, patch - 3, patch - 1, patch}):
            if stride < 1:
                continue
            bdr = (patch - stride) // 2
            for h0, w0 in ((patch - 2 * bdr, 3 * patch + 1), (bdr, 57), (19, 19), (40, 56), (108, 156), (5, 128)):
                want = LF.num_patches(h0, w0, patch, stride)
                rc = lib.lft_lf_num_patches_ex(h0, w0, patch, stride, C.byref(nu), C.byref(nv))
                uncovered = want[0] * stride < h0 or want[1] * stride < w0   # the reference's LFintegrate fails (utils.py:155)
                if h0 < max(bdr, 1) or w0 < max(bdr, 1) or min(want) < 1 or uncovered:
                    assert rc == -1, (h0, w0, patch, stride)
                    continue
                assert rc == 0, (h0, w0, patch, stride, lib.lft_last_error())
                assert (nu.value, nv.value) == want, (h0, w0, patch, stride)
                lf = torch.zeros(2 * h0, 2 * w0)
                if patch * stride <= 256 or stride >= patch // 2:  # keep the oracle calls cheap
                    assert tuple(O.lf_divide(lf, 2, patch, stride).shape[:2]) == want
                checked += 1
    assert checked > 100
    # the two-argument entry point is the reference default (32, 16)
    assert lib.lft_lf_num_patches(108, 156, C.byref(nu), C.byref(nv)) == 0 and (nu.value, nv.value) == (7, 10)
    for bad in ((64, 64, 65, 16), (64, 64, 3, 1), (64, 64, 32, 0), (64, 64, 16, 17), (3, 64, 32, 16)):
        assert lib.lft_lf_num_patches_ex(*bad, C.byref(nu), C.byref(nv)) == -1, bad


def test_tiling_rejects_uncovered_last_row():
    """an odd (patch - stride) can leave numU*stride == h0 - 1 (patch 32, stride 21, h0 43): the reference's
    LFintegrate raises a shape mismatch there (utils.py:155); the C side must refuse instead of leaving the last SR rows
    unwritten.  Neighbouring sizes that do cover the view are accepted."""
    lib = capi.load()
    nu, nv = C.c_int32(), C.c_int32()
    assert lib.lft_lf_num_patches_ex(43, 43, 32, 21, C.byref(nu), C.byref(nv)) == -1
    assert b"LFintegrate" in lib.lft_last_error()
    assert lib.lft_lf_num_patches_ex(43, 64, 32, 21, C.byref(nu), C.byref(nv)) == -1
    assert lib.lft_lf_num_patches_ex(42, 40, 32, 21, C.byref(nu), C.byref(nv)) == 0 and (nu.value, nv.value) == (2, 2)
    assert lib.lft_lf_num_patches_ex(44, 44, 32, 21, C.byref(nu), C.byref(nv)) == 0 and nu.value * 21 >= 44


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, s, n = 3, 2, 7  # ragged: 4 + 3 patches
        ranges = LF.patch_ranges(n, world)
        p0, p1 = ranges[rank]
        c = 16 * s
        local = torch.stack([torch.full((A, A, c, c), float(p)) + torch.arange(c * c).view(c, c) / 4096.0
                             for p in range(p0, p1)]) if p1 > p0 else torch.zeros(0, A, A, c, c)
        got = LF.gather_crops(local, ranges, rank, world)
        if rank == 0:
            want = torch.stack([torch.full((A, A, c, c), float(p)) + torch.arange(c * c).view(c, c) / 4096.0
                                for p in range(n)])
            q.put(bool(torch.equal(got, want)))
        else:
            q.put(got is None)
    finally:
        dist.destroy_process_group()


def test_gather_crops_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(res)


def test_crop_slab_integrates_like_reference_tiler():
    """The [n,A,A,16s,16s] crop slab + patch-order placement equals LFintegrate on whole SR patches."""
    A, s, h0, w0 = 3, 2, 40, 56
    nu, nv = LF.num_patches(h0, w0)
    g = torch.Generator().manual_seed(0)
    sr_patches = torch.rand(nu, nv, A * 32 * s, A * 32 * s, generator=g)
    want = O.lf_integrate(sr_patches, A, 32 * s, 16 * s, h0 * s, w0 * s)
    c, b = 16 * s, 8 * s
    crops = sr_patches.view(nu * nv, A, 32 * s, A, 32 * s)[:, :, b:b + c, :, b:b + c].permute(0, 1, 3, 2, 4)
    full = crops.reshape(nu, nv, A, A, c, c).permute(2, 3, 0, 4, 1, 5).reshape(A, A, nu * c, nv * c)
    assert torch.equal(full[:, :, :h0 * s, :w0 * s], want)


def test_psnr_per_view_matches_reference_definition():
    """utils.py:79 (skimage PSNR of non-negative float images = 10 log10(1 / MSE)) per view, utils.py:85 mean over views."""
    from lft_b200.evalloop import psnr_per_view
    A, H, W = 3, 20, 28
    g = torch.Generator().manual_seed(3)
    hr = torch.rand(A * H, A * W, generator=g)
    sr = (hr + 0.05 * torch.randn(A * H, A * W, generator=g)).clamp(0, 1)
    got = psnr_per_view(sr, hr, A).numpy()
    for u in range(A):
        for v in range(A):
            a = hr[u * H:(u + 1) * H, v * W:(v + 1) * W].double().numpy()
            b = sr[u * H:(u + 1) * H, v * W:(v + 1) * W].double().numpy()
            assert abs(got[u, v] - 10 * np.log10(1.0 / np.mean((a - b) ** 2))) < 1e-9


def test_ssim_per_view_against_direct_convolution():
    """ssim_per_view (scipy separable filter) against an independent direct 11 x 11 evaluation of the same definition
    (Gaussian sigma 1.5, reflect borders, sample covariances = x 121 / 120 as scikit-image's default, data range 2) on the
    interior pixels."""
    from lft_b200.evalloop import ssim_per_view
    A, H, W = 2, 24, 30
    g = torch.Generator().manual_seed(4)
    hr = torch.rand(A * H, A * W, generator=g)
    sr = (hr + 0.1 * torch.randn(A * H, A * W, generator=g)).clamp(0, 1)
    got = ssim_per_view(sr, hr, A).numpy()
    k1 = np.exp(-0.5 * (np.arange(-5, 6) / 1.5) ** 2)
    k1 /= k1.sum()
    k2 = np.outer(k1, k1)
    c1, c2 = (0.01 * 2.0) ** 2, (0.03 * 2.0) ** 2
    for u in range(A):
        for v in range(A):
            a = hr[u * H:(u + 1) * H, v * W:(v + 1) * W].double().numpy()
            b = sr[u * H:(u + 1) * H, v * W:(v + 1) * W].double().numpy()
            vals = []
            for y in range(5, H - 5):          # interior: the window never touches the border, so no padding rule is involved
                for x in range(5, W - 5):
                    wa, wb = a[y - 5:y + 6, x - 5:x + 6], b[y - 5:y + 6, x - 5:x + 6]
                    ua, ub = (k2 * wa).sum(), (k2 * wb).sum()
                    n = 121.0 / 120.0
                    va, vb = n * ((k2 * wa * wa).sum() - ua * ua), n * ((k2 * wb * wb).sum() - ub * ub)
                    vab = n * ((k2 * wa * wb).sum() - ua * ub)
                    vals.append((2 * ua * ub + c1) * (2 * vab + c2) / ((ua * ua + ub * ub + c1) * (va + vb + c2)))
            assert abs(got[u, v] - np.mean(vals)) < 1e-9
    same = ssim_per_view(hr, hr, A)
    assert torch.allclose(same, torch.ones(A, A, dtype=torch.float64))


class _StubEngine:
    """CPU stand-in with the Engine methods LightFieldSR uses: 'SR' = nearest-neighbour upsampling of each LR patch view, so
    that the sharded path (patch_ranges -> per-rank crops -> gather -> integrate) can run under gloo without a GPU.  The
    tiling itself comes from the oracle (pinned against the reference in test_oracle_golden.py)."""

    def __init__(self, A, s):
        self.A, self.s = A, s

    def engine(self, device):
        return self

    def num_patches(self, h0, w0, patch=32, stride=16):
        return LF.num_patches(h0, w0, patch, stride)

    def forward_lf_crops(self, lr_lf, p0, p1, out=None, max_ws_bytes=None, patch=32, stride=16):
        A, s = self.A, self.s
        sub = O.lf_divide(lr_lf, A, patch, stride)
        nu, nv = sub.shape[:2]
        flat = sub.view(nu * nv, A, patch, A, patch)[p0:p1]
        sr = flat.repeat_interleave(s, dim=2).repeat_interleave(s, dim=4)            # [n, A, P*s, A, P*s]
        c, b = stride * s, ((patch - stride) * s) // 2
        return sr[:, :, b:b + c, :, b:b + c].permute(0, 1, 3, 2, 4).contiguous()      # [n, A, A, c, c]

    def integrate(self, crops, h0, w0, p0, p1, sr_lf, patch=32, stride=16):
        A, s = self.A, self.s
        nu, nv = LF.num_patches(h0, w0, patch, stride)
        c = stride * s
        full = crops.reshape(nu, nv, A, A, c, c).permute(2, 0, 4, 3, 1, 5).reshape(A, nu * c, A, nv * c)
        sr_lf.copy_(full[:, :h0 * s, :, :w0 * s].reshape(A * h0 * s, A * w0 * s))
        return sr_lf


def _sharded_worker(rank, world, port, q, patch, stride):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, s, h0, w0 = 3, 2, 40, 56
        lf = torch.from_numpy(synth.synth_light_field(A, h0, w0, 5))
        got = LF.LightFieldSR(_StubEngine(A, s), patch=patch, stride=stride)(lf, rank=rank, world=world)
        if rank == 0:
            want = lf.view(A, h0, A, w0).repeat_interleave(s, dim=1).repeat_interleave(s, dim=3).reshape(A * h0 * s, A * w0 * s)
            q.put(bool(torch.equal(got, want)))
        else:
            q.put(got is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("patch,stride", [(32, 16), (32, 24), (16, 8)])
def test_light_field_sharding_world2_gloo(patch, stride):
    """LightFieldSR's rank logic end to end on two gloo ranks: ragged patch ranges (12 / 6 / 35 patches over 2 ranks),
    one gather, integrate on rank 0.  With a nearest-neighbour stand-in for the network the assembled light field must be
    the upsampled input exactly (the divide -> crop -> integrate identity of the reference tiler, utils.py:91-157)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + patch + stride
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q, patch, stride)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(res)


def test_h5_test_set_reader_with_stand_in_h5py(tmp_path, monkeypatch):
    """evalloop.TestSetDataLoader / MultiTestSetDataLoader (utils_datasets.py:40-98): directory layout, the column-major
    transpose and the item shapes, with a stand-in `h5py` (the real one is absent here) that serves .npz files."""
    import sys
    import types as _t
    from lft_b200 import evalloop as E
    A, s, h0, w0 = 3, 2, 6, 10
    root = tmp_path / "data_for_test"
    for ds, n in (("HCI_new", 2), ("EPFL", 1)):
        d = root / f"SR_{A}x{A}_{s}x" / ds
        d.mkdir(parents=True)
        for i in range(n):
            lr = np.arange(A * h0 * A * w0, dtype=np.float32).reshape(A * h0, A * w0) + i
            hr = np.arange(A * h0 * s * A * w0 * s, dtype=np.float32).reshape(A * h0 * s, A * w0 * s) - i
            np.savez(d / f"scene{i}.h5", Lr_SAI_y=lr.T, Hr_SAI_y=hr.T)           # stored transposed, like MATLAB's h5write
            (d / f"scene{i}.h5.npz").rename(d / f"scene{i}.h5")

    class _File:
        def __init__(self, path, mode):
            self._z = np.load(path)
        def __enter__(self):
            return self
        def __exit__(self, *a):
            self._z.close()
        def get(self, k):
            return self._z[k]
    fake = _t.ModuleType("h5py")
    fake.File = _File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    args = _t.SimpleNamespace(path_for_test=str(root) + "/", angRes=A, scale_factor=s, data_name="ALL", num_workers=0)
    names, loaders, n = E.MultiTestSetDataLoader(args)
    assert names == ["EPFL", "HCI_new"] and n == 3 and [len(l) for l in loaders] == [1, 2]
    lr, hr = next(iter(loaders[1]))
    assert tuple(lr.shape) == (1, 1, A * h0, A * w0) and tuple(hr.shape) == (1, 1, A * h0 * s, A * w0 * s)
    assert float(lr[0, 0, 0, 1]) in (1.0, 2.0) and float(lr[0, 0, 1, 0]) in (A * w0, A * w0 + 1.0)   # row-major again
    assert len(E.TestSetDataLoader(args, "EPFL")) == 1
