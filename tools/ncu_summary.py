#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key counters per kernel + hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print("==", d.get("Kernel Name", "?")[:90])
        for k in KEYS:
            if k in d:
                print("   %-82s %s %s" % (k, d[k], units[hdr.index(k)]))


def _num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def source(rep, top):
    """Aggregate the per-SASS-instruction samples by CUDA source line (needs -lineinfo)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, func, hdr = None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] in ("", "-"):
            continue  # per-SASS rows repeat what the per-line rows already sum
        key = (func, fname, r[0], r[1].strip())
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += _num(r[isamp])
        a[1] += _num(r[iinst])
    funcs = sorted({k[0] for k in agg})
    for f in funcs:
        items = [(k, v) for k, v in agg.items() if k[0] == f]
        tot = sum(v[0] for _, v in items) or 1.0
        toti = sum(v[1] for _, v in items) or 1.0
        items.sort(key=lambda kv: -kv[1][0])
        print("-- %s: stall-sample share | warp-instruction share | file:line | source" % f[:60])
        for k, v in items[:top]:
            print("   %5.1f%% %5.1f%%  %s:%s  %s" % (100 * v[0] / tot, 100 * v[1] / toti, k[1], k[2], k[3][:110]))


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))
