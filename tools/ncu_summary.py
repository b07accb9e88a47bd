#!/usr/bin/env python3
"""Condense an .ncu-rep into the text summary committed under profiles/ (raw metrics + hot SASS regions).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_kernel.txt "<command that was profiled>" """
import csv, subprocess, sys

rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
lines = [f"# ncu summary of {rep}", f"# command: {cmd}", f"# kernel: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}", ""]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append(f"{w:75s} {vals[i]:>18s} {units[i]}")
hot = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hot.py"), rep, "2"], capture_output=True, text=True).stdout
lines += ["", "# hot SASS regions (instruction index range, executions per instruction, share of warp instructions,",
          "# share of stall samples, average active lanes, dominant opcodes)", hot]
open(out, "w").write("\n".join(lines))
print("wrote", out)
