#!/usr/bin/env python
"""Summarise ncu reports into profiles/: key raw metrics per kernel + top source lines.
usage: ncu_summary.py [--select | --what "command ..."] OUT.md report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "gpc__cycles_elapsed.avg.per_second",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        res.append((d.get("Kernel Name", "?"), {k: (d.get(k, ""), units[hdr.index(k)] if k in hdr else "") for k in KEYS}))
    return res


def main():
    args = sys.argv[1:]
    what = ("`python bench.py --steps 1 --warmup 3` (config 2: 10,000 windows x 8 haplotypes x 64 reads)")
    if args and args[0] == "--select":   # the haplotype selection stage (tools/gpu_select_profile.sh)
        args = args[1:]
        what = ("`python tools/select_bench.py 10000 1 --pin` (synth-select-v1: 10,000 windows x 8 candidate variants, one "
                "launch of a late round: ~50 trial haplotypes x 11 sampled reads per window, half of the windows)")
    if args and args[0] == "--what":    # free text: the command the reports were captured under
        what = args[1]
        args = args[2:]
    out_md, reps = args[0], args[1:]
    assert out_md.endswith(".md"), "usage: ncu_summary.py [--select] OUT.md report.ncu-rep ..."
    with open(out_md, "w") as f:
        f.write("# ncu summaries (`ncu --set full --clock-control none --import-source on`, one launch per kernel)\n\n")
        f.write("Captured on a B200 under gpurun while running " + what + ".  Times under the profiler are "
                "cold-cache and serialised; bench.py's numbers come from CUDA events outside the profiler.\n\n")
        for rep in reps:
            for name, vals in raw(rep):
                f.write("## %s\n\nreport: `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % (name.split("(")[0], rep))
                for k in KEYS:
                    v, u = vals[k]
                    if v != "":
                        f.write("| %s | %s | %s |\n" % (k, v, u))
                top = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "12"],
                                     capture_output=True, text=True).stdout
                f.write("\nTop source lines by executed instructions:\n\n```\n%s```\n\n" % top)


if __name__ == "__main__":
    main()
