"""Per-kernel counter / stall table (markdown) from an `ncu --set full` report (CPU side): tools/ncu_counters.py <rep> [<title>]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__cycles_active.avg']
stalls = ['long_scoreboard', 'not_selected', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'lg_throttle', 'selected',
          'no_instruction', 'mio_throttle', 'barrier', 'membar']
names = [r[idx['Kernel Name']].split('(')[0].replace('void ', '') for r in rows[2:]]
cnt = {}
cols = []
for n in names:
    cnt[n] = cnt.get(n, 0) + 1
    cols.append(n if names.count(n) == 1 else f"{n} #{cnt[n]}")
print(f"# {title}\n")
print("Stall columns: average number of warps per scheduler stalled for that reason per issued instruction "
      "(`smsp__average_warps_issue_stalled_*_per_issue_active`).\n")
print("| metric | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for w in want:
    if w in idx:
        print(f"| {w} [{units[idx[w]]}] | " + " | ".join(r[idx[w]] for r in rows[2:]) + " |")
for s in stalls:
    k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
    if k in idx:
        print(f"| stall {s} | " + " | ".join(f"{float(r[idx[k]]):.2f}" for r in rows[2:]) + " |")
