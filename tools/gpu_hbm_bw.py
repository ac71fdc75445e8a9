"""HBM write-only / read-only / copy bandwidth with plain torch ops (context for the store-heavy kernels)."""
import torch
n = 1 << 30  # 4 GiB of fp32
x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
tw = t(lambda: x.fill_(1.0)); print(f"write-only  (fill_)  : {4 * n / tw / 1e12:.2f} TB/s")
tr = t(lambda: x.sum());      print(f"read-only   (sum)    : {4 * n / tr / 1e12:.2f} TB/s")
tc = t(lambda: y.copy_(x));   print(f"copy        (copy_)  : {8 * n / tc / 1e12:.2f} TB/s (read + write bytes)")
