"""Print the key ncu raw-page metrics per kernel of a .ncu-rep (CPU side)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_barriers', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_alu.sum']
for r in rows[2:]:
    print("\n##", r[idx['Kernel Name']].split('(')[0])
    for w in want:
        if w in idx: print(f"{w}: {r[idx[w]]} {units[idx[w]]}")
