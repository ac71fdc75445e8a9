"""tcgen05 issue/operand micro-benchmark (see lft_mma_bench in include/lft_b200.h)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lft_b200 import capi
lib = capi.load()
reps = 200
for ctas_per_sm, smem in ((1, 150 * 1024), (2, 100 * 1024)):
    for mode in (0, 1, 2, 3, 4):
        for N in (64, 128, 256):
            K = 128
            grid = 148 * ctas_per_sm
            buf = (C.c_int64 * grid)()
            capi.check(lib.lft_mma_bench(N, K, reps, mode, grid, smem, buf))
            cyc = np.array(list(buf), dtype=np.float64)
            n_mma = reps * K // 16
            print(f"ctas/SM {ctas_per_sm} mode {('SS', 'TS', 'SS+1row', 'SS sw128', 'SS sw128+1row')[mode]} N={N:3d}: {cyc.mean() / n_mma:7.1f} cycles/MMA per CTA "
                  f"(math floor {128 * N / 256:.0f}) -> per-SM {cyc.mean() / n_mma / ctas_per_sm:6.1f}", flush=True)
