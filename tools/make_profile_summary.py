"""Build profiles/<tag>_ncu_summary.md from ncu reports + a launch-list CSV + a bench JSON (CPU side).
usage: make_profile_summary.py <tag> <bench.json> <launches.csv> <rep1> [rep2 ...]"""
import csv, io, json, subprocess, sys
from collections import defaultdict
tag, bench, launches = sys.argv[1:4]
reps = sys.argv[4:]
out = [f"# {tag}: ncu --set full --clock-control none (one launch per kernel kind, fp32 mode) + launch list + live CUDA-event shares",
       "# commands: see the bottom of this file"]
seen = set()
for rep in reps:
    txt = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
    for block in txt.split("\n## ")[1:]:
        name = block.split("\n", 1)[0]
        if name in seen:
            continue
        seen.add(name)
        out.append("## " + block.rstrip())
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[4].split("(")[0].replace("void ", "")
    tot[k][0] += 1; tot[k][1] += float(r[-1]) / 1e6
s = sum(v[1] for k, v in tot.items() if k.startswith("k_"))
out.append(f"\n## ncu launch list (profiles/{tag}_ncu_launches.csv): launches and share of the library's kernels' time (cold-cache, serialised)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("k_"):
        out.append(f"{k:22s} launches {v[0]:4d}  total {v[1]:8.3f} ms  share {100 * v[1] / s:5.1f}%")
d = json.load(open(bench))
ks = d["kernels"]; st = sum(v["ms_per_step"] for v in ks.values())
out.append(f"\n## live CUDA-event shares from bench.py (same build, {bench}), ms per step / share; whole step {d['ms_per_step']:.2f} ms")
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_step"]):
    out.append(f"{k:22s} {v['ms_per_step']:7.3f} ms  share {100 * v['ms_per_step'] / st:5.1f}%  algorithmic {v['tflops']:7.1f} TFLOP/s")
out.append(f"""
## commands (one gpurun call; every ncu pass after the same bench command had exited 0 without ncu)
python bench.py --steps 20 --warmup 5
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/{tag.split('_')[-1]}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
ncu --set full --clock-control none --import-source on -k regex:'k_spa_ffn|k_spa_attn|k_ang|k_spa_embed|k_conv3x3|k_up_gemm|k_up_gather' -s 40 -c 9 -o gpurun_out/prof_{tag.split('_')[-1]} python bench.py --steps 2 --warmup 3 --no-cpu-baseline
(the launch list is cut at 400 launches)""")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[-40:]))
