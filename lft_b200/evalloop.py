"""test.py-compatible evaluation loop: the body of `test(test_loader, device, net)`
(test.py:73-111) with the per-patch Python loop replaced by the batched device path.

    for Lr_SAI_y, Hr_SAI_y in loader:             # [1, A*h0, A*w0] / [1, A*h0*s, A*w0*s] like TestSetDataLoader
        Sr_SAI_y = sr(Lr_SAI_y.squeeze().cuda())   # LFdivide -> forward(all patches) -> LFintegrate on the GPU

Metrics are not part of the hot path (skimage PSNR/SSIM, out of scope); a per-view PSNR with the
reference's definition (utils.py:79,85: 10 log10(1/MSE) on [0,1] images, mean over views) is provided because the
bf16 gate is stated in PSNR, and `ssim_per_view` restates the SSIM of utils.py:82-84 (CPU, scipy; unpinned - see there).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.utils.data

from .lightfield import LightFieldSR


def psnr_per_view(sr_sai: torch.Tensor, hr_sai: torch.Tensor, angRes: int) -> torch.Tensor:
    """[A, A] PSNR of every view of two SAI mosaics [A*H, A*W] (utils.py:56-88 without SSIM)."""
    A = angRes
    H, W = sr_sai.shape[0] // A, sr_sai.shape[1] // A
    d = (sr_sai.double() - hr_sai.double().to(sr_sai.device)).view(A, H, A, W).permute(0, 2, 1, 3)
    mse = (d * d).mean(dim=(2, 3)).clamp_min(1e-20)
    return 10.0 * torch.log10(1.0 / mse)


def ssim_per_view(sr_sai: torch.Tensor, hr_sai: torch.Tensor, angRes: int, data_range: float = 2.0) -> torch.Tensor:
    """[A, A] SSIM of every view, restating what `metrics.structural_similarity(label, out, gaussian_weights=True)` of
    utils.py:82-84 computes: Gaussian window sigma 1.5 truncated at 3.5 sigma (11 x 11), SAMPLE covariances (the call leaves
    scikit-image's `use_sample_covariance=True`, so vx, vy and vxy carry the factor NP / (NP - 1) = 121 / 120), K1 = 0.01,
    K2 = 0.03, borders of (11 - 1) / 2 pixels dropped before the mean.  `data_range`: the reference passes none; the
    scikit-image of its era (<= 0.18; README.md:15 names python 3.6 / PyTorch 1.3) then takes the dtype range of float
    images, -1..1, i.e. 2.0 - later versions refuse float images without it.  UNPINNED: scikit-image is not installed in the
    authoring container, so this function is checked only against an independent direct-convolution evaluation of the same
    definition (tests/test_host_cpu.py), not against the reference's library.  CPU-side (scipy), not part of the hot path."""
    import numpy as np
    from scipy.ndimage import gaussian_filter
    A = angRes
    H, W = sr_sai.shape[0] // A, sr_sai.shape[1] // A
    x = hr_sai.detach().cpu().double().numpy().reshape(A, H, A, W).transpose(0, 2, 1, 3)   # label first, as in utils.py:82
    y = sr_sai.detach().cpu().double().numpy().reshape(A, H, A, W).transpose(0, 2, 1, 3)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    pad = 5                                                                                # (win_size - 1) // 2, win_size = 11
    cov_norm = 121.0 / 120.0                                                               # NP / (NP - 1), NP = win_size ** 2
    out = np.zeros((A, A))
    f = lambda im: gaussian_filter(im, sigma=1.5, truncate=3.5)                            # mode='reflect', scipy's default
    for u in range(A):
        for v in range(A):
            a, b = x[u, v], y[u, v]
            ux, uy = f(a), f(b)
            vx, vy, vxy = (cov_norm * (f(a * a) - ux * ux), cov_norm * (f(b * b) - uy * uy),
                           cov_norm * (f(a * b) - ux * uy))
            smap = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
            out[u, v] = smap[pad:H - pad, pad:W - pad].mean()
    return torch.from_numpy(out)


class TestSetDataLoader(torch.utils.data.Dataset):
    """The h5 test-set reader of utils_datasets.py:67-98: every file of `<path_for_test>SR_<A>x<A>_<s>x/<data_name>/` holds
    `Lr_SAI_y` / `Hr_SAI_y` (written by the MATLAB preparation column-major, hence the transpose at :86-87); items are
    `([1, A*h0, A*w0], [1, A*h0*s, A*w0*s])` float tensors like `ToTensor()` of a 2-D array gives.  Needs `h5py`, which the
    authoring image lacks: the import is deferred to the first item and fails loudly there (tests feed a stand-in module)."""

    def __init__(self, args, data_name: str = "ALL"):
        import os
        self.dataset_dir = f"{args.path_for_test}SR_{args.angRes}x{args.angRes}_{args.scale_factor}x/"
        self.file_list = [data_name + "/" + f for f in os.listdir(self.dataset_dir + data_name)]
        self.item_num = len(self.file_list)

    def __getitem__(self, index):
        import numpy as np
        try:
            import h5py
        except ImportError as e:   # no fallback format: the reference's test sets are h5 files
            raise ImportError("lft_b200.evalloop.TestSetDataLoader needs h5py to read the reference's test sets") from e
        with h5py.File(self.dataset_dir + self.file_list[index], "r") as hf:
            lr = np.transpose(np.array(hf.get("Lr_SAI_y")), (1, 0))
            hr = np.transpose(np.array(hf.get("Hr_SAI_y")), (1, 0))
        return torch.from_numpy(lr.copy()).float()[None], torch.from_numpy(hr.copy()).float()[None]

    def __len__(self):
        return self.item_num


def MultiTestSetDataLoader(args):
    """utils_datasets.py:40-64: one batch-size-1 loader per test set found under the data directory (every sub-directory,
    like the reference's `os.listdir`; sorted here so that the order does not depend on the file system)
    -> (names, loaders, number of scenes)."""
    import os
    from torch.utils.data import DataLoader
    dataset_dir = f"{args.path_for_test}SR_{args.angRes}x{args.angRes}_{args.scale_factor}x/"
    data_list = sorted(os.listdir(dataset_dir))
    loaders, n = [], 0
    for name in data_list:
        ds = TestSetDataLoader(args, name)
        n += len(ds)
        loaders.append(DataLoader(dataset=ds, num_workers=getattr(args, "num_workers", 0), batch_size=1, shuffle=False))
    return data_list, loaders, n


def cal_metrics(angRes: int, label: torch.Tensor, out: torch.Tensor) -> Tuple[float, float]:
    """utils.py:56-88 for two SAI mosaics: per-view PSNR / SSIM averaged over the views with a NON-ZERO value
    (`PSNR.sum() / np.sum(PSNR > 0)`, utils.py:85-86)."""
    import numpy as np
    ps = psnr_per_view(out, label, angRes).cpu().numpy().astype("float32")
    ss = ssim_per_view(out, label, angRes).numpy().astype("float32")
    return float(ps.sum() / max(int((ps > 0).sum()), 1)), float(ss.sum() / max(int((ss > 0).sum()), 1))


@torch.no_grad()
def test(test_loader: Iterable, device, net, angRes: Optional[int] = None, patch_size_for_test: int = 32,
         stride_for_test: int = 16, outputs: Optional[List[torch.Tensor]] = None) -> Tuple[float, float]:
    """Mirror of test.py:73-111: returns (psnr_epoch_test, ssim_epoch_test) like the reference; the SR SAI mosaics are
    appended (on the CPU) to `outputs` if a list is passed.
    patch_size_for_test / stride_for_test are the reference's `args` of the same names (option.py:16-17; test.py:83,96)."""
    A = angRes if angRes is not None else net.angRes
    sr = LightFieldSR(net, patch=patch_size_for_test, stride=stride_for_test)
    psnrs, ssims = [], []
    for Lr_SAI_y, Hr_SAI_y in test_loader:
        lr = Lr_SAI_y.squeeze().to(device, torch.float32).contiguous()   # test.py:77
        Sr_SAI_y = sr(lr)                                                 # test.py:83-101
        if outputs is not None:
            outputs.append(Sr_SAI_y.cpu())
        if Hr_SAI_y is not None:
            p, q = cal_metrics(A, Hr_SAI_y.squeeze(), Sr_SAI_y.cpu())    # test.py:103
            psnrs.append(p)
            ssims.append(q)
    nan = float("nan")
    return (sum(psnrs) / len(psnrs) if psnrs else nan), (sum(ssims) / len(ssims) if ssims else nan)
