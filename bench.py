#!/usr/bin/env python
"""bench.py -- lightcurve evals/s of the LFIT CV eclipse-model hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement (oracle) arm

A "step" is one pass of the hot path over one batch: ln_prob (priors + model +
chi-squared, mcmcfit.py:37-41) for every walker of the ensemble.  The workload is
BASELINE.json configs[1]: one complex-BS eclipse, 2000 phase points, 4096 walkers,
exposure-width smearing.  Under torchrun each rank owns its own 4096 walkers (weak
scaling) and the ranks all-gather positions and log-probs over NCCL, as an emcee
half-step would.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e`
goes through the public host API (pinned H2D of theta, D2H of ln_prob inside the
timed region).  `roofline` is the lightcurve kernel against the FP64 FMA rate measured
on the same device in the same run; `cpu_baseline` is the CPU oracle on the box's host
cores over a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lightcurve_evals_per_s"
UNIT = "lightcurve evals/s"

# FP64 operations per light-curve evaluation of the bench workload (FMA = 2, add = mul = 1,
# compares / conversions 0), from ncu instruction counters of this very command
# (smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on, tabulated by tools/launch_table.py);
# see DESIGN.md "Roofline accounting".  "elements" = the four elements_kernel launches (stage 1),
# "all" = every kernel of a log-probability pass.
#   algorithmic: the solver with every Newton step in FP64 -- the fixed per-unit figure that
#                roofline.achieved is quoted on (it does not move when the kernels get cleverer);
#   executed:    what the committed kernels issue today (the first Newton steps run in FP32, the
#                white-dwarf tiles start from the centre's solution).
FLOPS_PER_LIGHTCURVE = {
    "r1": {"elements": 2.458e6, "all": 2.962e6, "source": "profiles/r01_launches_fp64solver.csv",
           "executed": {"elements": 1.396e6, "all": 1.811e6, "source": "profiles/r01_launches.csv"},
           # dram__bytes_read.sum + dram__bytes_write.sum of the four elements_kernel launches of one batch of
           # 2048 light curves (ncu --set full, profiles/r01_elements_kernel.txt), per light curve
           "dram_bytes_per_lightcurve": 8.58e6 / 2048},
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def oracle_layout(O, wl):
    return O.FlatLayout(wl.ndim, wl.npars, wl.gather, wl.consts, wl.prior_src, wl.prior_type, wl.prior_p1,
                        wl.prior_p2, wl.prior_norm, wl.prior_isvar, wl.lc_off, wl.lc_phase, wl.lc_width,
                        wl.lc_y, wl.lc_ye)


def cpu_baseline(wl, theta, target_s=12.0):
    """The CPU oracle (port of the path) on all host threads over a bounded sample of the walkers."""
    from oracle import oracle as O
    cfg = O.config(**{k: v for k, v in getattr(wl, "grid", {}).items()})
    lay = oracle_layout(O, wl)
    cores = O.max_threads()
    n0 = min(theta.shape[0], 2 * cores)
    t0 = time.perf_counter()
    O.log_prob(lay, theta[:n0], what=2, cfg=cfg)
    dt = max(time.perf_counter() - t0, 1e-3)
    n1 = int(min(theta.shape[0], max(n0, n0 * target_s / dt)))
    n1 = max(cores, (n1 // cores) * cores)
    t0 = time.perf_counter()
    O.log_prob(lay, theta[:n1], what=2, cfg=cfg)
    dt = time.perf_counter() - t0
    return {"value": n1 * wl.n_ecl / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d walkers x %d eclipses, ln_prob, %.1f s" % (n1, theta.shape[0], wl.n_ecl, dt)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  lfit / trm.roche are
    not under /root/reference, so this is the oracle port (oracle/), all host threads."""
    if rank != 0:
        return
    from oracle import oracle as O
    from lfit_python_b200 import workloads
    wl = workloads.config(args.config)
    if args.n_ph:
        wl = workloads.config(args.config, n_ph=args.n_ph)
    cfg = O.config(**wl.grid)
    wl.make_data(lambda p, x, w: O.calc_flux(p, x, w, cfg=cfg)[1])
    lay = oracle_layout(O, wl)
    cores = O.max_threads()
    n_sample = args.ref_sample or 8 * cores
    theta = wl.walkers(n_sample, ln_prior_fn=lambda t: O.log_prob(lay, t, what=0, cfg=cfg))
    for _ in range(args.warmup):
        O.log_prob(lay, theta[: 2 * cores], what=2, cfg=cfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.log_prob(lay, theta, what=2, cfg=cfg)
    dt = time.perf_counter() - t0
    value = n_sample * wl.n_ecl * args.steps / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "walkers_per_step": n_sample, "eclipses": wl.n_ecl, "n_phase": wl.n_ph,
                   "note": "CPU restatement of the lfit path (oracle/), OpenMP over walkers; lfit itself is not "
                           "vendored in the reference tree; each step is a bounded sample of the 4096-walker batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d walkers x %d eclipses per step" % (n_sample, wl.n_ecl)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args, rank, local_rank, world):
    import torch
    from lfit_python_b200 import _cabi, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = workloads.config(args.config)
    if args.n_ph:
        wl = workloads.config(args.config, n_ph=args.n_ph)
    n = args.walkers or wl.n_walkers
    eng = _cabi.Engine(local_rank, **wl.grid)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    # every rank draws its own shard of the ensemble (weak scaling: n walkers per GPU)
    theta_h = wl.walkers(n, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR), seed=2024 + rank)
    theta_d = torch.from_numpy(theta_h).cuda()
    lnp_d = torch.empty(n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 2x the 126 MB L2
    if world > 1:
        all_pos = torch.empty(world * n, wl.ndim, dtype=torch.float64, device="cuda")
        all_lnp = torch.empty(world * n, dtype=torch.float64, device="cuda")
    # a dedicated non-default stream: the C ABI reads a NULL stream as "the handle's own stream",
    # and torch's default stream has handle 0
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        eng.log_prob_device(theta_d.data_ptr(), n, lnp_d.data_ptr(), what=_cabi.LN_PROB, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(all_lnp, lnp_d)
            dist.all_gather_into_tensor(all_pos, theta_d)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    fp64_peak = eng.measure_fp64_peak()
    launches0 = eng.launch_count

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, stage_ms = [], []
    for i in range(args.steps):
        flush.fill_(float(i))  # evict L2 between timed iterations (outside the event pair)
        ev[i][0].record()
        step()
        ev[i][1].record()
        ev[i][1].synchronize()
        kernel_ms.append(eng.last_kernel_ms())
        stage_ms.append(eng.last_stage_ms())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())

    # end to end through the public host API: pinned H2D of theta + D2H of ln_prob every step
    e2e_steps = args.steps
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lnp_h = eng.log_prob(theta_h, what=_cabi.LN_PROB)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert np.array_equal(np.isfinite(lnp_h), np.isfinite(lnp_d.cpu().numpy()))

    # emcee steps/s: the stretch move of mcmc_utils.EnsembleSampler (host side, emcee's algebra) driving the
    # vectorised CUDA log-probability -- one step = two half-steps = n log-probability evaluations per GPU
    from lfit_python_b200 import mcmc_utils
    sampler = mcmc_utils.EnsembleSampler(n, wl.ndim, lambda t: eng.log_prob(t, what=_cabi.LN_PROB), vectorize=True,
                                         rng=np.random.default_rng(99 + rank))
    state = sampler.run_mcmc(theta_h, 2, store=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_mc = min(max(5, args.steps), 50)
    sampler.run_mcmc(state[0], n_mc, log_prob0=state[1], store=False)
    mc_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mc_s, op=dist.ReduceOp.MAX)
    mc_steps_per_s = n_mc / float(mc_s.item())
    acc_frac = float(sampler.acceptance_fraction.mean())

    # the same move with the ensemble resident in HBM (no PCIe traffic per step)
    class _Vec:
        engine, ndim = eng, wl.ndim
    dsampler = mcmc_utils.DeviceEnsembleSampler(n, _Vec, seed=7 + rank)
    dsampler.run_mcmc(theta_h, 2)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dsampler.run_mcmc(None, n_mc)
    dmc_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dmc_s, op=dist.ReduceOp.MAX)
    dmc_steps_per_s = n_mc / float(dmc_s.item())

    # clean per-stage device times for the roofline: the same pass with the two batch lanes
    # serialised (LFB_LANES=1), so that the element solves are timed alone, by CUDA events
    # recorded on the stream they run on
    os.environ["LFB_LANES"] = "1"
    eng1 = _cabi.Engine(local_rank, **wl.grid)
    os.environ.pop("LFB_LANES")
    wl.apply(eng1)
    eng1.set_trace(True)
    serial_stage, serial_trace = [], []
    for i in range(3 + min(args.steps, 20)):
        flush.fill_(float(i))
        eng1.log_prob_device(theta_d.data_ptr(), n, lnp_d.data_ptr(), what=_cabi.LN_PROB, stream=stream)
        torch.cuda.synchronize()
        if i >= 3:
            serial_stage.append(eng1.last_stage_ms())
            serial_trace.append(eng1.last_trace_ms())
    eng1.close()

    # the same pass judged by the Gaussian process (useGP = 1 trees): residuals + gp_kernel instead of chi-squared
    gp = None
    if rank == 0 and not args.no_gp:
        os.environ["LFB_LANES"] = "1"
        eng_gp = _cabi.Engine(local_rank, **wl.grid)
        os.environ.pop("LFB_LANES")
        wl.apply_gp(eng_gp)
        eng_gp.set_trace(True)
        gp_ms, gp_tr = [], []
        for i in range(3 + min(args.steps, 20)):
            flush.fill_(float(i))
            eng_gp.log_prob_device(theta_d.data_ptr(), n, lnp_d.data_ptr(), what=_cabi.LN_PROB, stream=stream)
            torch.cuda.synchronize()
            if i >= 3:
                gp_ms.append(eng_gp.last_stage_ms()["total"])
                gp_tr.append(eng_gp.last_trace_ms()["gp_kernel"])
        eng_gp.close()
        gp = {"ms_per_pass_one_lane": float(np.mean(gp_ms)), "gp_kernel_ms": float(np.mean(gp_tr)),
              "lightcurve_evals_per_s": n * wl.n_ecl / (float(np.mean(gp_ms)) * 1e-3),
              "note": "GPLCModel likelihood (4-state Kalman filter per eclipse) on the same walkers, one lane"}

    if rank == 0:
        evals_per_step = world * n * wl.n_ecl
        value = evals_per_step * args.steps / (total_ms * 1e-3)
        k_ms = float(np.mean(kernel_ms))
        fl = FLOPS_PER_LIGHTCURVE.get(args.kernel_rev)
        stages = {k: float(np.mean([d[k] for d in stage_ms])) for k in stage_ms[0]}
        serial = {k: float(np.mean([d[k] for d in serial_stage])) for k in serial_stage[0]}
        trace = {k: float(np.mean([d[k] for d in serial_trace])) for k in serial_trace[0]}
        el_ms = sum(v for k, v in trace.items() if k.startswith("elements_kernel"))  # the four stage-1 launches
        roof = {"bound": "fp64", "kernel": "elements_kernel<wd,disc,spot,donor> (stage 1: Roche ingress/egress solves)",
                "kernel_ms": el_ms, "kernel_share_of_step": el_ms / serial["total"],
                "stage_ms_serial": serial, "kernel_ms_serial": trace, "stage_ms_overlapped": stages, "pipeline_ms": k_ms,
                "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "DFMA probe on this device, this run "
                "(MEASURED_PEAKS.json has no FP64 vector figure)", "traffic": None}
        if fl and wl.name.startswith("C2") and not args.n_ph and not wl.grid:
            per_rank = n * wl.n_ecl
            roof["flops_per_lightcurve"] = fl
            roof["traffic"] = fl["dram_bytes_per_lightcurve"] * per_rank  # bytes per step of the elements stage (ncu)
            roof["achieved"] = fl["elements"] * per_rank / (el_ms * 1e-3) * 1e-12
            roof["frac"] = roof["achieved"] / fp64_peak
            roof["whole_pass"] = {"achieved": fl["all"] * per_rank / (k_ms * 1e-3) * 1e-12,
                                  "frac": fl["all"] * per_rank / (k_ms * 1e-3) * 1e-12 / fp64_peak}
            ex = fl["executed"]
            roof["executed_fp64"] = {"achieved": ex["elements"] * per_rank / (el_ms * 1e-3) * 1e-12,
                                     "frac": ex["elements"] * per_rank / (el_ms * 1e-3) * 1e-12 / fp64_peak,
                                     "note": "FP64 flops the kernels issue (FP32 warm-up steps not counted)"}
        else:
            roof["achieved"] = None
            roof["frac"] = None
        # HBM side of the same kernel, for the record: theta in, chi-squared out, light curve re-read per CTA
        alg_bytes = n * wl.n_ecl * (18 * 8 + 8 + 2 * (16 * 903 + 8 * 200 + 32 * 103) + 2 * 16 * 2012)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback"
        roof["hbm"] = {"achieved": alg_bytes / (k_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": alg_bytes / (k_ms * 1e-3) * 1e-9 / hbm_peak, "peak_source": hbm_src,
                       "note": "algorithmic bytes of the whole pass: theta in, chi-squared out, element and event records written and read once"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "walkers_per_gpu": n, "eclipses": wl.n_ecl, "n_phase": wl.n_ph,
                       "ndim": wl.ndim, "grid": eng.config, "l2": "flushed between timed iterations (256 MB fill)",
                       "parallelism": "walkers sharded, %d rank(s)" % world,
                       "collective": "nccl all_gather of ln_prob + positions" if world > 1 else "none"},
            "ensemble_passes_per_s": args.steps / (total_ms * 1e-3),
            "gp_likelihood": gp,
            "emcee_steps_per_s": mc_steps_per_s,
            "emcee": {"walkers_per_gpu": n, "steps_timed": n_mc, "acceptance_fraction": acc_frac,
                      "note": "host stretch move (numpy) + one CUDA ln_prob call per half-step, host buffers",
                      "device_resident_steps_per_s": dmc_steps_per_s},
            "e2e": {"value": evals_per_step * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * wl.ndim * 8), "d2h_bytes_per_step": int(n * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(wl, theta_h, target_s=args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json config index (default 1: the metric's config)")
    ap.add_argument("--walkers", type=int, default=0, help="override walkers per GPU")
    ap.add_argument("--n-ph", type=int, default=0, help="override phase points per eclipse")
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: walkers per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel-rev", default="r1")
    ap.add_argument("--no-gp", action="store_true", help="skip the Gaussian-process likelihood leg")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
