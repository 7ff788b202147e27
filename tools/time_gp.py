#!/usr/bin/env python
"""Time ln_prob passes judged by the Gaussian process (useGP = 1 trees) on a bench workload.

    [LFB_LANES=n] python tools/time_gp.py [config] [walkers] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lfit_python_b200 import _cabi, workloads

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = workloads.config(cfg)
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl.n_walkers
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
eng = _cabi.Engine(0, **wl.grid)
wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
wl.apply_gp(eng)
theta = wl.walkers(n, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR), seed=2024)
td = torch.from_numpy(theta).cuda()
out = torch.empty(n, dtype=torch.float64, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
ms = []
with torch.cuda.stream(st):
    for i in range(reps + 5):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        eng.log_prob_device(td.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
        b.record(st)
        st.synchronize()
        if i >= 5:
            ms.append(a.elapsed_time(b))
ms = np.array(ms)
print("GP pass (lanes %s): %.4f ms median (%d walkers, %.3f M lc/s); finite %d" % (
    os.environ.get("LFB_LANES", "default"), np.median(ms), n, n * wl.n_ecl / np.median(ms) / 1e3, int(torch.isfinite(out).sum())))
eng.close()
