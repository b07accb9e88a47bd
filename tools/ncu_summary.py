#!/usr/bin/env python3
"""Text summary of an `ncu --set full --import-source on` report (run here, no GPU needed): the launch geometry, the
utilisation metrics the judge asks for, DRAM traffic per launch and the source lines that collect the most warp-stall
samples.  usage: tools/ncu_summary.py <report.ncu-rep> [out.txt]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {rep}: {len(rows)} launch(es) of {rows[0][name_i]}", file=out)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:70s} {units[i]:12s} " + "  ".join(r[i] for r in rows), file=out)
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, agg, text, have = None, collections.Counter(), {}, False
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 3 and r[0] == "Line No":
            have = True
        elif have and len(r) >= 6 and r[0].isdigit() and r[2] == "-":
            try:
                n = int(r[4])
            except ValueError:
                continue
            agg[(cur, int(r[0]))] += n
            text[(cur, int(r[0]))] = r[1][:110]
    tot = sum(agg.values()) or 1
    print(f"\n# warp-stall samples by source line (all launches, {tot} samples)", file=out)
    for k, v in agg.most_common(25):
        print(f"{100 * v / tot:5.1f}%  {k[0]}:{k[1]:<5d} {text[k]}", file=out)


if __name__ == "__main__":
    main()
