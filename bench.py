#!/usr/bin/env python
"""Headline benchmark: SR output megapixels/s for 5x5, 4x full light-field inference (BASELINE.json configs[2]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]
  (N > 1: launched by torchrun, one rank per GPU, NCCL)

A step = one pass of the hot path over ONE synthetic light field of HCInew shape (5x5 views of 128x128 LR -> 512x512 SR,
64 overlapping 32x32 patches).  At N > 1 the 64 patches of that ONE light field are sharded over the ranks
(`lightfield.patch_ranges(64, N)`, 8 per GPU at N = 8) and the SR light field is re-assembled on rank 0 ("strong"
scaling: total work fixed) - by peer stores into rank 0's buffer over NVLink (default) or the NCCL gather
(`--assemble collective`).  `value` counts INTEGRATED output pixels (after LFintegrate) / max-over-ranks device time,
inputs resident in HBM.  `e2e` is the same metric through the public API (lightfield.HostPipeline) from pinned host memory
to pinned host memory, every copy inside the timed region.  The old weak-scaling number (one whole light field per GPU)
is kept as the extra key `weak`.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A, S, H0, W0 = 5, 4, 128, 128
PATCHES = 64
MP_PER_LF = A * A * H0 * S * W0 * S / 1e6          # 6.5536
TOKENS_PER_LF = PATCHES * A * A * 32 * 32          # 1,638,400
# algorithmic FLOP per LR token and launch (window attention at its mean 23.16 keys)
FLOP_PER_TOKEN = {
    "conv3x3_64": 73728,   # + conv_init0 (1,152) fused into the first launch
    "ang_fused": 71936, "spa_embed_qkv": 147456 + 98304,
    "spa_attn": 11858, "spa_ffn": 180224, "up_gemm": 131072 + 18432,
}
FLOP_PER_LF = 2411464 * TOKENS_PER_LF              # 3.951 TFLOP: the reference's work per light field (every token of every patch)
# algorithmic HBM bytes per unit for the bandwidth-bound kernels (fp32, each operand once): window attention reads Q, K, V and
# writes O (4 x 128 floats per token); the gather reads 9 tap sums and writes one SR pixel (+ bicubic taps from L2)
BYTES_PER_UNIT = {"spa_attn": 4 * 128 * 4, "up_gather": (9 + 1) * 4, "lf_divide": 8, "lf_integrate": 8}
TENSOR_KINDS = set(FLOP_PER_TOKEN) - {"spa_attn"}


def ncu_traffic(kind):
    """DRAM bytes per launch of `kind` from the newest committed ncu --set full capture (profiles/r*_traffic.json)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if files:
        return json.load(open(files[-1])).get(kind), os.path.basename(files[-1])
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload_config(world):
    """`config` of the JSON line - the same dict for both arms (the reference arm times a bounded sample of this workload;
    what the sample was is in its `cpu_baseline.sample`)."""
    return {"workload": "LFT 5x5 4x full-LF inference, ONE HCInew-shape light field per step: 5x5x128x128 LR -> 512x512 SR, 64 "
                        "patches of 32x32 (stride 16) sharded over the GPUs, SR light field assembled on rank 0",
            "weights": "seeded synthetic checkpoint in the reference format (shipped pth absent)",
            "l2": "working set ~6.8 GB per step at N=1 (0.85 GB per GPU at N=8) >> 126 MB L2 (no flush needed)",
            "parallelism": f"patch-sharded dp{world}"}


REF_PATCHES_PER_STEP = 3


def cpu_baseline(sd, lf, n_patches, threads=None):
    """The reference's CPU path (test.py:83-101: LFdivide, one net() call per patch, LFintegrate) on a bounded sample: the
    UNMODIFIED reference staged under baseline/_ref when present (kind "reference"), else the oracle port (kind "port")."""
    import torch
    from oracle import reference_arm as R
    if threads:
        torch.set_num_threads(threads)
    _, n, dt, kind = R.run_light_field(sd, lf, A, S, max_patches=n_patches)
    mp = n * A * A * (16 * S) ** 2 / 1e6
    return mp / dt, dt, n, torch.get_num_threads(), kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads,
    REF_PATCHES_PER_STEP patches of the benchmark's light field per step (the whole light field is 64)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from lft_b200 import synth
    sd = synth.synth_state_dict(A, S, 0)
    lf = torch.from_numpy(synth.synth_light_field(A, H0, W0, 2))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "port"
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(sd, lf, 1)
    times = []
    for _ in range(args.steps):
        v, dt, n, th, kind = cpu_baseline(sd, lf, REF_PATCHES_PER_STEP)
        times.append(dt)
    mp_per_step = REF_PATCHES_PER_STEP * A * A * (16 * S) ** 2 / 1e6
    val = mp_per_step * len(times) / sum(times)
    what = ("the unmodified reference (model/LFT.py get_model + utils.py LFdivide / LFintegrate, staged in baseline/_ref)"
            if kind == "reference" else "oracle port of test.py:83-101 (dense masked attention, mask rebuilt per call)")
    line = {
        "impl": "reference", "metric": "SR output megapixels/sec (5x5 4x full LF)", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": cores, "kind": kind,
                         "sample": f"{len(times)} steps x {REF_PATCHES_PER_STEP} of the 64 patches of the light field per step "
                                   f"(LFdivide of the whole light field + B=1 net() per patch + LFintegrate, test.py:83-101); {what}"},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) write to fd 1 from C, so fd 1 is
    pointed at stderr for the rest of the process and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--impl", default="lft_b200", choices=["lft_b200", "reference"])
    ap.add_argument("--assemble", default="auto", choices=["auto", "direct", "collective"],
                    help="N > 1: peer stores into rank 0's SR buffer (direct) or NCCL gather + integrate (collective)")
    ap.add_argument("--cpu-patches", type=int, default=32, help="patches in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling leg at N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # keep stdout to the single JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"

    import torch
    import torch.distributed as dist
    from lft_b200 import synth
    from lft_b200.engine import Engine
    from lft_b200.lightfield import HostPipeline, LightFieldSR, patch_ranges

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd = synth.synth_state_dict(A, S, 0)
    eng = Engine(A, S, precision=args.precision, device=local)
    eng.load_state_dict(sd)
    lf_host = torch.from_numpy(synth.synth_light_field(A, H0, W0, 2)).pin_memory()   # the SAME light field on every rank
    lf = lf_host.to(dev)
    sr_pipe = LightFieldSR(eng, assemble=args.assemble)

    def step():
        return sr_pipe(lf, rank, world)    # patches [p0, p1) of this rank; SR light field assembled on rank 0

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    sync()
    assembled = "single GPU" if world == 1 else ("direct peer stores (CUDA IPC over NVLink) + 2 one-element all-reduces"
                                                   if any(v is not None for v in sr_pipe._peer.values()) else "NCCL gather + lft_integrate")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        step()
    e1.record()
    sync()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - n0
    # per-kernel durations: the same K steps once more with the library's per-launch CUDA events (an event between two
    # kernels serialises them, so this pass runs without programmatic dependent launch; its step time is reported too)
    eng.profile_enable(True)
    p0_, p1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    p0_.record()
    for i in range(args.steps):
        step()
    p1_.record()
    sync()
    profiled_ms = max_over_ranks(p0_.elapsed_time(p1_))
    prof = eng.profile_read()
    eng.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = MP_PER_LF * args.steps / (total_ms / 1e3)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    launches_all = int(lt.item())

    # ---- e2e: public API (HostPipeline), pinned host -> device -> SR -> pinned host, every step; at N > 1 every rank uploads
    # the light field from ITS pinned host copy and rank 0 downloads the assembled SR light field
    sr_hosts = [torch.empty(A * H0 * S, A * W0 * S, dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else [None, None]
    pipe = HostPipeline(sr_pipe)   # copies of step i overlap the kernels of step i+1 (all inside the timed region)
    e2e_i = [0]

    def e2e_step():
        pipe.submit(lf_host, sr_hosts[e2e_i[0] & 1], rank=rank, world=world)
        e2e_i[0] += 1

    for _ in range(6):   # untimed: every pipeline slot reused twice, so the caching allocator has reached its steady state
        e2e_step()
    pipe.drain()
    sync()
    n_e2e = max(3, min(args.steps, 10))
    blocks = []
    for _ in range(3):   # three timed blocks of n_e2e steps; the median block is reported, all three are listed
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        b0.record()
        for _ in range(n_e2e):
            e2e_step()
        pipe.drain()   # every result is in host memory before the closing event
        b1.record()
        sync()
        blocks.append(max_over_ranks(b0.elapsed_time(b1)))
    e2e_blocks = [MP_PER_LF * n_e2e / (b / 1e3) for b in blocks]
    e2e_val = sorted(e2e_blocks)[1]

    # ---- extra: weak scaling (one WHOLE light field per GPU per step, assembled where it was computed)
    weak = None
    if world > 1 and not args.no_weak:
        own = LightFieldSR(eng)
        lf_own = torch.from_numpy(synth.synth_light_field(A, H0, W0, 2 + rank)).to(dev)
        for _ in range(3):
            own(lf_own)
        w0_, w1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        w0_.record()
        for _ in range(args.steps):
            own(lf_own)
        w1_.record()
        sync()
        wms = max_over_ranks(w0_.elapsed_time(w1_))
        weak = {"value": world * MP_PER_LF * args.steps / (wms / 1e3), "unit": "MP/s", "ms_per_step": wms / args.steps,
                "what": "one whole light field per GPU per step (per-GPU work fixed), no cross-GPU traffic"}

    if rank == 0:
        pk = peaks()
        kinds = {k: v for k, v in prof.items() if v["launches"] > 0}
        step_kernel_ms = sum(v["ms"] for v in kinds.values()) / args.steps
        top = max(kinds, key=lambda k: kinds[k]["ms"])
        avg_ms = kinds[top]["ms"] / kinds[top]["launches"]
        units_per_launch = kinds[top]["units"] / kinds[top]["launches"]
        if top in TENSOR_KINDS:
            ach = FLOP_PER_TOKEN[top] * units_per_launch / (avg_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"]}
        else:
            ach = BYTES_PER_UNIT.get(top, 0) * units_per_launch / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
        traffic, traffic_src = ncu_traffic(top)
        roof.update({"kernel": top, "avg_launch_ms": avg_ms, "units_per_launch": units_per_launch,
                     "units": "LR tokens (pixels x views) the launch processes, averaged over the layers (the last layers of the "
                              "light-field path run on the pixels the kept crop depends on only)",
                     "share_of_step": kinds[top]["ms"] / args.steps / step_kernel_ms,
                     "traffic": traffic if world == 1 else None,
                     "traffic_unit": f"bytes/launch at N=1 (dram read+write, ncu --set full, profiles/{traffic_src}); the captured launch is "
                                     f"the first block's, which processes all {TOKENS_PER_LF} tokens of the light field (units_per_launch above "
                                     "is the mean over the four blocks)",
                     "peak_source": pk["src"],
                     "note": ("fp32 path issues 3 bf16 MMAs per product (hi*hi+lo*hi+hi*lo): attainable frac <= 1/3"
                              if args.precision == "fp32" else "single bf16 MMA per product")})
        executed = sum(FLOP_PER_TOKEN.get(k, 0) * v["units"] for k, v in kinds.items())   # rank 0's share
        whole_ref = FLOP_PER_LF * args.steps / (total_ms * 1e-3) / 1e12
        line = {
            "metric": "SR output megapixels/sec (5x5 4x full LF)", "value": value, "unit": "MP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32 (bf16x3 split on tcgen05, fp32 accumulate)" if args.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_val, "unit": "MP/s", "h2d_bytes_per_step": int(lf_host.numel() * 4) * world,
                    "d2h_bytes_per_step": int(A * H0 * S * A * W0 * S * 4), "steps": n_e2e,
                    "blocks": [round(b, 1) for b in e2e_blocks], "reported": "median of 3 blocks",
                    "what": "lightfield.HostPipeline: every rank uploads the LR light field from pinned host memory, rank 0 "
                            "downloads the assembled SR light field into pinned host memory; copies overlap the next step's kernels"},
            "gpu_launches": launches_all, "clocks": clocks, "roofline": roof, "assembly": assembled,
            "kernel_timing": {"ms_per_step_with_events": profiled_ms / args.steps,
                              "what": "roofline / kernels come from a second pass of the same K steps with per-launch CUDA events on the "
                                      "launching stream; `value` / `ms_per_step` from the first pass without them (kernels chained by "
                                      "programmatic dependent launch)"},
            "patches_per_rank": [b - a for a, b in patch_ranges(PATCHES, world)],
            "whole_step": {"reference_work_tflops": whole_ref, "frac_of_bf16_peak": whole_ref / (pk["tensor"] * world),
                           "executed_tflops_rank0": executed / (total_ms * 1e-3) / 1e12,
                           "note": "reference_work = 2,411,464 FLOP x every token of every patch (SURVEY 8d) / step time, divided by "
                                   "the peak of all N GPUs; executed = what rank 0's kernels really ran after skipping the pixels "
                                   "LFintegrate discards"},
            "kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps, "units_per_step": v["units"] / args.steps,
                            "tflops": (FLOP_PER_TOKEN.get(k, 0) * v["units"] / max(v["ms"], 1e-9) / 1e9)}
                        for k, v in kinds.items()},
        }
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_cpu_baseline:
            v, dt, n, th, kind = cpu_baseline(sd, lf_host.clone(), args.cpu_patches)
            line["cpu_baseline"] = {"value": v, "unit": "MP/s", "cores": th, "kind": kind,
                                    "sample": f"{n} of 64 patches of the same light field, B=1 per call, {dt:.1f} s "
                                              + ("(the unmodified reference staged in baseline/_ref: LFdivide + get_model per patch + LFintegrate, test.py:83-101)"
                                                 if kind == "reference" else "(oracle port of test.py:83-101, dense masked attention)")}
        _emit(line)
    sr_pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
