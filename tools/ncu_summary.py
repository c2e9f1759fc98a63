#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel time, instructions, DRAM bytes, stall
reasons and (with --src KERNEL) the SASS lines with the most stall samples.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--src fwd_tile_kernel] [--top 25]"""
import csv, subprocess, sys, io

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]

def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        print("=====", r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, r[i], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stalls per issue:", ", ".join("%s %.2f" % (h, v) for v, h in sorted(st, reverse=True)[:7]))
    if "--src" in sys.argv:
        kn = sys.argv[sys.argv.index("--src") + 1]
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kn], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        ia, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        seen = set()
        data = []
        for i, r in enumerate(rows[2:]):
            if len(r) > isamp and r[isamp].isdigit() and r[0] not in seen:   # the page may list the function twice
                seen.add(r[0])
                data.append((int(r[isamp]), i, r[ia], r[iex]))
        tot = sum(d[0] for d in data)
        print("total samples", tot, "instructions", len(data))
        for s_, i, src, ex in sorted(data, reverse=True)[:top]:
            print("%6d %5.1f%% line %5d exec %9s  %s" % (s_, 100.0 * s_ / tot, i, ex, src[:100]))

main()
