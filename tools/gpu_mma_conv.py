"""tcgen05 issue pattern of the 3x3 conv (108 MMAs per 'rep') on resident operands."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import capi
lib = capi.load()
reps = 50
for ctas, smem in ((1, 150 * 1024), (2, 100 * 1024)):
    for N in (64, 128):
        for mode, name in ((5, "201-row planes + tap shifts"), (6, "208-row planes + tap shifts"), (7, "201-row planes, no shifts")):
            grid = 148 * ctas
            buf = (C.c_int64 * grid)()
            capi.check(lib.lft_mma_bench(N, 64, reps, mode, grid, smem, buf))
            cyc = np.array(list(buf), dtype=np.float64).mean() / (reps * 108)
            print(f"ctas/SM {ctas} N={N:3d} {name:30s}: {cyc:6.1f} cycles/MMA per CTA -> per-SM {cyc / ctas:6.1f}", flush=True)
