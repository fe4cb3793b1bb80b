#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_sdf.ncu-rep profiles/r01_sdf_tiles_kernel.md [launches.csv]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (see scripts/gpu_check.sh);",
             "per-launch times under ncu are cold-cache and serialised — compare shares, not absolutes.", ""]
    lines.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(vals))) + " |")
    lines.append("|---|---|" + "---|" * len(vals))
    lines.append("| kernel | | " + " | ".join(v[name_col].split("(")[0] for v in vals) + " |")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {units[i]} | " + " | ".join(v[i] for v in vals) + " |")
    if len(sys.argv) > 3:
        lines += ["", "## launch list (`ncu --metrics gpu__time_duration.sum`, same command)", "", "```"]
        tot = {}
        for r in csv.DictReader(l for l in open(sys.argv[3]) if not l.startswith("==")):
            n = r["Kernel Name"].split("(")[0].split("<")[0]
            a = tot.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"]) / 1e3
        whole = sum(v[1] for v in tot.values())
        for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{us:12.1f} us  {100 * us / whole:5.1f} %  x{c:<4d} {n}")
        lines.append("```")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
