#!/usr/bin/env python
"""Tabulate an `ncu --csv --metrics ...` launch list: one row per kernel name (mean over launches).

    python tools/launch_table.py gpurun_out/launches.csv [n_lightcurves_per_launch] [--json out.json]

With --json the per-kernel FP64 operation counts (FMA = 2, add = mul = 1; ncu counters
smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on), DRAM bytes and shared-memory wavefronts per
light curve are written next to a hash of the kernel sources they were measured on; bench.py quotes its
roofline fractions on that file and says so when the sources have changed since.
"""
import csv
import hashlib
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = ("cv_kernels.cuh", "roche_device.cuh", "gp_device.cuh", "lfit_cabi.cu", "sampler.cuh", "peer.cuh", "angle_table.inc")


def csrc_sha():
    """Hash of the kernel sources (what the counters belong to)."""
    h = hashlib.sha256()
    for name in CSRC:
        with open(os.path.join(ROOT, "lfit_python_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def short(name):
    name = name.split('(')[0].replace("void ", "").replace("lfb::", "")
    return name.strip()


def main(path, n_lc=None, json_out=None):
    rows = list(csv.reader(open(path)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr_i]
    ik, im, iv, iid = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
    per = OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) < len(H):
            continue
        per.setdefault(r[iid], {"name": r[ik]})[r[im]] = float(r[iv].replace(',', ''))
    agg = OrderedDict()
    for d in per.values():
        a = agg.setdefault(short(d["name"]), {"n": 0})
        a["n"] += 1
        for k, v in d.items():
            if k != "name":
                a[k] = a.get(k, 0.0) + v
    tot_t = sum(a.get('gpu__time_duration.sum', 0) / a["n"] for a in agg.values())
    print("%-34s %4s %10s %6s %12s %9s %7s %10s %11s %7s" % ("kernel", "n", "time_us", "share", "fp64_flop", "TFLOP/s", "lanes",
                                                             "dram_MB", "warp_inst", "issue%"))
    tot_f = 0.0
    tot_i = 0.0
    kern = OrderedDict()
    for name, a in agg.items():
        n = a["n"]
        t = a.get('gpu__time_duration.sum', 0) / n * 1e-3
        fl = (2 * a.get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 0)
              + a.get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 0)
              + a.get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 0)) / n
        tot_f += fl
        lanes = a.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0) / n
        dram = (a.get('dram__bytes_read.sum', 0) + a.get('dram__bytes_write.sum', 0)) / n
        smem_wf = a.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0) / n
        winst = a.get('smsp__inst_executed.sum', 0) / n
        issue = a.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0) / n
        tot_i += winst
        kern[name] = {"launches": n, "time_us": t, "fp64_flop": fl, "dram_bytes": dram, "smem_wavefronts": smem_wf,
                      "lanes": lanes, "warp_inst": winst, "issue_active_pct": issue}
        print("%-34s %4d %10.1f %5.1f%% %12.4g %9.2f %7.1f %10.2f %11.4g %7.1f" % (
            name[:34], n, t, 100 * t * 1e3 / tot_t, fl, fl / (t * 1e-6) * 1e-12 if t else 0, lanes, dram * 1e-6, winst, issue))
    print("sum of kernel times %.1f us, FP64 flop per pass %.4g, warp instructions per pass %.4g" % (tot_t * 1e-3, tot_f, tot_i))
    if n_lc:
        print("FP64 flop per light curve: %.4g; warp instructions per light curve: %.4g" % (tot_f / n_lc, tot_i / n_lc))
    if json_out and n_lc:
        def grp(pred, key):
            return sum(k[key] for nm, k in kern.items() if pred(nm)) / n_lc
        out = {
            "csrc_sha": csrc_sha(), "source": os.path.basename(path), "lightcurves_per_launch": n_lc,
            "convention": "FMA = 2, DADD = DMUL = 1 (executed, predicated-on thread instructions); per light curve",
            "per_lightcurve": {
                "elements_flop": grp(lambda nm: nm.startswith("elements_kernel") or nm.startswith("stage1_kernel"), "fp64_flop"),
                "flux_flop": grp(lambda nm: nm.startswith("flux_kernel"), "fp64_flop"),
                "all_flop": tot_f / n_lc,
                "all_warp_inst": tot_i / n_lc,
                "elements_warp_inst": grp(lambda nm: nm.startswith("elements_kernel"), "warp_inst"),
                "flux_warp_inst": grp(lambda nm: nm.startswith("flux_kernel"), "warp_inst"),
                "elements_dram_bytes": grp(lambda nm: nm.startswith("elements_kernel") or nm.startswith("stage1_kernel"), "dram_bytes"),
                "flux_dram_bytes": grp(lambda nm: nm.startswith("flux_kernel"), "dram_bytes"),
                "flux_smem_wavefronts": grp(lambda nm: nm.startswith("flux_kernel"), "smem_wavefronts"),
                "all_dram_bytes": sum(k["dram_bytes"] for k in kern.values()) / n_lc,
            },
            "kernels": {nm: {k: (v / n_lc if k in ("fp64_flop", "dram_bytes", "smem_wavefronts") else v)
                             for k, v in d.items()} for nm, d in kern.items()},
        }
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", json_out)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:]]
    jo = None
    if "--json" in args:
        i = args.index("--json")
        jo = args[i + 1]
        del args[i:i + 2]
    if "--sha" in args:
        print(csrc_sha())
        sys.exit(0)
    main(args[0], float(args[1]) if len(args) > 1 else None, jo)
