"""Stress the multi-stream pass for races: many back-to-back passes in every calling pattern, results compared
with a reference pass (a race shows up as a changed result or a CUDA error)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lfit_python_b200 import _cabi, workloads, mcmc_utils


def main(reps=200):
    eng = _cabi.Engine(0)
    wl = workloads.config(1)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(4096, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR))
    ref = eng.log_prob(theta)
    td = torch.from_numpy(theta).cuda()
    out = torch.empty(4096, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    bad = 0
    phase = "device pointers, back to back"
    try:
        with torch.cuda.stream(st):
            for i in range(reps):
                eng.log_prob_device(td.data_ptr(), 4096, out.data_ptr(), stream=st.cuda_stream)
                if i % 10 == 9:
                    st.synchronize()
                    bad += int(not np.array_equal(out.cpu().numpy(), ref))
        st.synchronize()
        print(phase, "mismatches", bad)
        phase = "host pointers"
        bad = 0
        for i in range(reps):
            bad += int(not np.array_equal(eng.log_prob(theta), ref))
        print(phase, "mismatches", bad)
        phase = "half ensembles (two lanes of 1024)"
        bad = 0
        ref2 = eng.log_prob(theta[:2048])
        for i in range(reps):
            bad += int(not np.array_equal(eng.log_prob(theta[:2048]), ref2))
        print(phase, "mismatches", bad)
        phase = "device sampler"
        s = mcmc_utils.DeviceSampler(eng, 4096, seed=3)
        s.set_state(theta)
        s.run(reps // 2)
        pos, lnp = s.get_state()
        print(phase, "consistent", bool(np.array_equal(lnp, eng.log_prob(pos))))
        s.close()
        phase = "one lane, traced"
        os.environ["LFB_LANES"] = "1"
        e1 = _cabi.Engine(0)
        os.environ.pop("LFB_LANES")
        wl.apply(e1)
        e1.set_trace(True)
        bad = 0
        with torch.cuda.stream(st):
            for i in range(reps // 2):
                e1.log_prob_device(td.data_ptr(), 4096, out.data_ptr(), stream=st.cuda_stream)
                st.synchronize()
                bad += int(not np.array_equal(out.cpu().numpy(), ref))
                e1.last_trace_ms()
        print(phase, "mismatches", bad)
        e1.close()
    except Exception as exc:
        print("FAILED in phase:", phase, "--", str(exc)[:300])
        raise
    eng.close()
    print("stress ok")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200)
