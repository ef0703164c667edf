#!/usr/bin/env python
"""Text summary of one ncu --set full report: key raw metrics + per-line / per-function table (tools/ncu_lines.py).
usage: ncu_summary.py <report.ncu-rep> <lib.so> <kernel-substring> > profiles/xxx.txt"""
import csv, subprocess, sys, os
rep, lib, kname = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u, v = r[0], r[1], r[2]
keys = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum"]
print(f"# ncu summary of {os.path.basename(rep)} (ncu --set full --clock-control none; cold-cache, replayed: use shares, not absolutes)")
for a, b, c in zip(h, u, v):
    if a in keys or (a.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in a and float(c or 0) > 0.05):
        print(f"{a:75s} {c} {b}")
print()
sys.stdout.flush()
subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), rep, lib, kname, "25"])
