import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_case(golden_dir, name):
    """A reference-generated forward golden (tests/golden/make_golden.py) with the synthetic weights / input it was made
    from: meta = [A, s, h, B, seed(, qk_gain, ln_wide)]."""
    import numpy as np
    import torch
    from lft_b200 import synth
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m = [int(x) for x in g["meta"]]
    A, s, h, B, seed = m[:5]
    qk, lnw = (float(m[5]), bool(m[6])) if len(m) > 5 else (1.0, False)
    sd = synth.synth_state_dict(A, s, seed, qk_gain=qk, ln_wide=lnw)
    lr = torch.from_numpy(synth.synth_lr_mosaic(B, A, h, h, seed))
    return g, A, s, sd, lr
