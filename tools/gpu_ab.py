"""A/B timing of experimental library variants (LFT_VARIANT builds, see lft_b200/build.py): runs the headline bench with
each variant in its own process and prints ms per step and per kernel kind.   python tools/gpu_ab.py '' st1 st2 ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
extra = os.environ.get("AB_BENCH_ARGS", "--steps 10 --warmup 3 --no-cpu-baseline").split()
for v in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if v:
        env["LFT_VARIANT"] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        rows[v or "default"] = d
        ks = "  ".join(f"{k}={x['ms_per_step']:.3f}" for k, x in d["kernels"].items() if x["ms_per_step"] > 0.1)
        print(f"{v or 'default':10s} step {d['ms_per_step']:.3f} ms  clk {d['clocks']['sm_mhz']}  {ks}", flush=True)
    except Exception as e:
        print(v, "FAILED", e, r.stderr[-800:], flush=True)
