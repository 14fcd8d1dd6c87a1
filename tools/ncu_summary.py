#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture (.ncu-rep) into a markdown file under profiles/: headline metrics of the kernel and
the source lines that execute the most instructions.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title" """
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = ["# " + title, "", "kernel: `%s`" % d.get("Kernel Name", ("?", ""))[0], "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, h = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    text = {}
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            h = r
            continue
        if h and cur and len(r) == len(h):
            x = dict(zip(h, r))
            try:
                ln, ie, sm, te = int(x["Line No"]), int(x["Instructions Executed"] or 0), int(x["# Samples"] or 0), int(x["Thread Instructions Executed"] or 0)
            except ValueError:
                continue
            a = agg[(cur, ln)]
            a[0] += ie
            a[1] += sm
            a[2] += te
            text[(cur, ln)] = r[1].strip()[:100]
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    lines += ["", "## source lines by executed warp instructions (top 25)", "", "| file:line | inst % | stall samples % | active lanes | source |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append("| %s:%d | %.2f | %.2f | %.1f | `%s` |" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots, v[2] / max(v[0], 1), text[k].replace("|", "\\|")))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
