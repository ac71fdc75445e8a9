"""tcgen05.commit cost: K/16 MMAs per group, an extra commit after every n-th group (n = 0: none)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import capi
lib = capi.load()
reps, grid = 400, 148
for N in (64, 128):
    for K in (64, 128):
        for every in (0, 1, 2):
            buf = (C.c_int64 * grid)()
            capi.check(lib.lft_mma_bench(N, K, reps, 0 + 16 * every, grid, 150 * 1024, buf))
            cyc = np.array(list(buf), dtype=np.float64).mean() / (reps * K // 16)
            print(f"SS N={N:3d}: groups of {K // 16} MMAs, commit every {every} group(s): {cyc:6.1f} cycles/MMA", flush=True)
