#!/usr/bin/env python
"""What the ncu launch list / --set full captures profile: plain ln_prob passes of the bench workload (C2 by default,
4096 walkers in ONE batch on one lane, device-resident theta), three warm passes and two more
(bracketed by cudaProfilerStart / Stop: run ncu with --profile-from-start off).

    python tools/profile_pass.py [config] [walkers]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LFB_LANES", "1")
import torch
from lfit_python_b200 import _cabi, workloads

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = workloads.config(cfg)
n = int(sys.argv[2]) if len(sys.argv) > 2 else wl.n_walkers
eng = _cabi.Engine(0, **wl.grid)
wl.make_noise_only_data()          # (no calc_flux launches in the list: the data only have to exist)
wl.apply(eng)
theta = wl.walkers(n, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR), seed=2024)   # the bench's ensemble
td = torch.from_numpy(theta).cuda()
out = torch.empty(n, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for i in range(5):
        if i == 3:
            torch.cuda.profiler.start()      # ncu --profile-from-start off: only the last two passes are profiled
        eng.log_prob_device(td.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
        st.synchronize()
    torch.cuda.profiler.stop()
print("passes done; finite:", int(torch.isfinite(out).sum()), "of", n, "launches", eng.launch_count)
eng.close()
