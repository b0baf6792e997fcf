#!/usr/bin/env python3
"""Prints the handful of ncu metrics we quote (reads a .ncu-rep here, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum']
want += [c for c in h if 'issue_stalled' in c and c.endswith('per_issue_active.ratio')]
for w in want:
    if w in h:
        i = h.index(w)
        vals = [r[i] for r in rows[1:]]
        if w.endswith('ratio') and 'stalled' in w:
            try:
                if all(float(v.replace(',', '')) < 0.05 for v in vals[1:]): continue
            except ValueError: pass
        print(f"{w:95s} {vals}")
