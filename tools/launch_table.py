#!/usr/bin/env python
"""Tabulate an `ncu --csv --metrics ...` launch list: one row per kernel name (mean over launches).

    python tools/launch_table.py gpurun_out/launches.csv [n_lightcurves_per_launch]
"""
import csv
import sys
from collections import OrderedDict


def main(path, n_lc=None):
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
        a = agg.setdefault(d["name"].split('(')[0], {"n": 0})
        a["n"] += 1
        for k, v in d.items():
            if k != "name":
                a[k] = a.get(k, 0.0) + v
    tot_t = sum(a.get('gpu__time_duration.sum', 0) / a["n"] for a in agg.values())
    print("%-34s %4s %10s %6s %12s %9s %7s" % ("kernel", "n", "time_us", "share", "fp64_flop", "TFLOP/s", "lanes"))
    tot_f = 0.0
    for name, a in agg.items():
        n = a["n"]
        t = a.get('gpu__time_duration.sum', 0) / n * 1e-3
        fl = (2 * a.get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 0)
              + a.get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 0)
              + a.get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 0)) / n
        tot_f += fl
        lanes = a.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0) / n
        print("%-34s %4d %10.1f %5.1f%% %12.4g %9.2f %7.1f" % (name[:34], n, t, 100 * t * 1e3 / tot_t, fl,
                                                              fl / (t * 1e-6) * 1e-12 if t else 0, lanes))
    print("sum of kernel times %.1f us, FP64 flop per pass %.4g" % (tot_t * 1e-3, tot_f))
    if n_lc:
        print("FP64 flop per light curve: %.4g" % (tot_f / n_lc))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
