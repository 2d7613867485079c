#!/usr/bin/env python
"""Summarise one `ncu --set full` report: key counters + instruction mix of the first profiled launch.
usage: python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/rNN_x_ncu.txt   (needs ncu on PATH, no GPU)"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## launch {d.get('ID')}: {d.get('Kernel Name', '')[:110]}")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"{k:92s} {units[hdr.index(k)]:16s} {d[k]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = list(csv.reader(io.StringIO(src)))
    try:
        h = lines[1]
        ie, so = h.index("Instructions Executed"), h.index("Source")
        ops = collections.Counter()
        for r in lines[2:]:
            if len(r) <= max(ie, so) or not r[ie].isdigit():
                continue
            t = r[so].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] += int(r[ie])
        tot = sum(ops.values())
        print(f"\n## SASS mix of the first launch ({tot} warp instructions executed)")
        for k, v in ops.most_common(16):
            print(f"{k:12s} {v / tot * 100:5.1f}%")
    except Exception as e:  # noqa: BLE001
        print("no source page:", e)


if __name__ == "__main__":
    main(sys.argv[1])
