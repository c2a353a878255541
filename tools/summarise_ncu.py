#!/usr/bin/env python3
"""Summarise `ncu --set full` captures (gpurun_out/<name>.ncu-rep) into profiles/<out>.md and refresh profiles/svr_traffic.json.
    python tools/summarise_ncu.py <out.md> <kernel>=<report> [<kernel>=<report> ...]      (run where ncu is installed, no GPU needed)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e-6, 'ms': 1e-3, 's': 1, 'ns': 1e-9}


def load(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def num(d, m):
    v, u = d[m]
    return float(v.replace(',', '')) * MULT.get(u, 1)


def main():
    import bench
    out_md = sys.argv[1]
    lines = ["# ncu captures (`--set full --clock-control none`, one launch each; times under ncu are cold-cache and serialised)", "",
             "Library build id (bench.py `library_build`): `%s`" % bench.library_build_id(), ""]
    for spec in sys.argv[2:]:
        kernel, rep = spec.split("=")
        d = load(rep)
        lines += ["## %s  (%s)" % (kernel, os.path.basename(rep)), "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in d:
                lines.append("| `%s` | %s | %s |" % (w, d[w][0], d[w][1]))
        r, w_ = num(d, 'dram__bytes_read.sum'), num(d, 'dram__bytes_write.sum')
        lines += ["", "DRAM per launch: %.1f MB read + %.1f MB written; duration %.3f ms." % (r / 1e6, w_ / 1e6, num(d, 'gpu__time_duration.sum') * 1e3), ""]
        if kernel == "k_svr_fact":
            json.dump({"kernel": "k_svr_fact", "dram_bytes_per_launch": int(r + w_), "dram_bytes_read": int(r), "dram_bytes_written": int(w_),
                       "candidates_per_launch": 2531484, "library_build": bench.library_build_id(),
                       "how": "ncu --set full --clock-control none -k regex:k_svr_fact -s 1 -c 1 python tools/profile_step.py 60 2 0 (%s): "
                              "dram__bytes_read.sum + dram__bytes_write.sum of the one launch that covers the 2.53 M-candidate bench panel" % os.path.basename(out_md),
                       "note": "reads are the work items' row tables (K-feat wrote them for the same launch) and states, writes are the scores"},
                      open(os.path.join(ROOT, "profiles", "svr_traffic.json"), "w"), indent=1)
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("\n".join(l for l in lines if l.startswith("DRAM") or l.startswith("##")))


if __name__ == "__main__":
    main()
