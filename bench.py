#!/usr/bin/env python
"""bench.py -- lightcurve evals/s (and emcee steps/s) of the LFIT CV eclipse-model hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--config i]   # the CUDA engine
    python bench.py --impl reference --gpus N --steps K ...       # the CPU restatement (oracle) arm

A "step" is one pass of the hot path over one batch: ln_prob (priors + model + chi-squared,
mcmcfit.py:37-41) for every walker of the ensemble.  The default workload is BASELINE.json configs[1]
(C2): one complex-BS eclipse, 2000 phase points, 4096 walkers, exposure-width smearing; --config 0..4
selects the others.  Under torchrun the ranks shard the walkers and, per step, all-gather ONE packed
buffer of positions + log-probs (the engine's own exchange over NVLink peer memory; NCCL as the fallback),
as a stretch-move half-step does:
C1-C3 weak scaling (the config's walkers per GPU), C4 / C5 strong scaling (BASELINE names their total).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the
public host API (pinned H2D of theta, D2H of ln_prob inside the timed region).  `roofline` is the
stage-1 element solve (the FP64-bound kernels) against the FP64 FMA rate measured on the same device in
the same run, on FP64 operations the kernels EXECUTE (ncu counters of profiles/r02_flops.json);
`roofline.kernels` carries the flux kernel against FP64 / HBM / shared memory.  `cpu_baseline` is the CPU
oracle on the box's host cores over a bounded sample.
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
STRONG_CONFIGS = (3, 4)   # BASELINE.json names a total ensemble for these ("sharded across 8xB200", "1/2/4/8 GPUs")
# SURVEY.md section 8d context figure: the reference's pure-Python tree walk alone (lfit / roche stubbed out,
# 6 eclipses, ndim 84) costs 15.4 ms per ln_prob = at most 390 light-curve evaluations/s per core before any physics
REFERENCE_TREE_OVERHEAD = {"ms_per_ln_prob": 15.4, "eclipses": 6, "lightcurve_evals_per_s_per_core_ceiling": 390,
                           "source": "SURVEY.md section 6 / 8d (measured with lfit and trm.roche stubbed)"}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads():
    """Cores this process may run on -- NOT OpenMP's default, which torchrun pins to 1 (OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def load_flops():
    """FP64 operations / DRAM bytes per light curve from the ncu launch list of the C2 pass
    (tools/launch_table.py --json), and whether the kernel sources changed since it was taken."""
    path = os.path.join(ROOT, "profiles", "r02_flops.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        fl = json.load(f)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import launch_table
        fl["stale"] = launch_table.csrc_sha() != fl.get("csrc_sha")
    except Exception:
        fl["stale"] = None
    if fl["stale"]:
        print("bench.py: WARNING profiles/r02_flops.json was measured on other kernel sources (csrc changed since): "
              "roofline fractions use outdated operation counts -- re-run the ncu launch list", file=sys.stderr)
    return fl


def cpu_baseline(wl, theta, target_s=12.0):
    """The CPU oracle (port of the path; built -O3 -march=native on this box) on all host cores over a bounded
    sample of the walkers."""
    from oracle import oracle as O
    native = O.use_native()
    cfg = O.config(**{k: v for k, v in getattr(wl, "grid", {}).items()})
    lay = oracle_layout(O, wl)
    cores = host_threads()
    n0 = min(theta.shape[0], 2 * cores)
    t0 = time.perf_counter()
    O.log_prob(lay, theta[:n0], what=2, cfg=cfg, nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-3)
    n1 = int(min(theta.shape[0], max(n0, n0 * target_s / dt)))
    n1 = max(min(cores, theta.shape[0]), (n1 // cores) * cores)
    t0 = time.perf_counter()
    O.log_prob(lay, theta[:n1], what=2, cfg=cfg, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": n1 * wl.n_ecl / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "build": "gcc -O3 -march=native -ffp-contract=off" if native else "gcc -O3 (generic x86-64)",
            "sample": "%d of %d walkers x %d eclipses, ln_prob, %.1f s" % (n1, theta.shape[0], wl.n_ecl, dt)}


def pick_workload(args, world):
    from lfit_python_b200 import workloads
    wl = workloads.config(args.config, n_ph=args.n_ph) if args.n_ph else workloads.config(args.config)
    scaling = args.scaling or ("strong" if args.config in STRONG_CONFIGS else "weak")
    total = args.walkers or wl.n_walkers
    if scaling == "strong":
        n = -(-total // world)
        n += n & 1
    else:
        n = total
    return wl, scaling, n


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  lfit / trm.roche are not under
    /root/reference, so this is the oracle port (oracle/), all host cores, on a bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    native = O.use_native()
    wl, scaling, n = pick_workload(args, world)
    cfg = O.config(**wl.grid)
    cores = host_threads()
    wl.make_data(lambda p, x, w: O.calc_flux(p, x, w, cfg=cfg)[1])
    lay = oracle_layout(O, wl)
    prior = lambda t: O.log_prob(lay, t, what=0, cfg=cfg, nthreads=cores)
    # size the per-step sample so that warmup + steps fit the budget: calibrate on one light curve per core
    theta_cal = wl.walkers(max(cores, 2), ln_prior_fn=prior)
    t0 = time.perf_counter()
    O.log_prob(lay, theta_cal, what=2, cfg=cfg, nthreads=cores)
    per_walker = (time.perf_counter() - t0) / theta_cal.shape[0] * min(cores, theta_cal.shape[0]) / cores
    budget = args.ref_seconds / max(args.steps + args.warmup, 1)
    full = n * world if scaling == "weak" else n * world
    n_sample = args.ref_sample or int(min(full, max(cores, budget / max(per_walker, 1e-9))))
    n_sample = max(cores, (n_sample // cores) * cores) if n_sample >= cores else n_sample
    theta = wl.walkers(n_sample, ln_prior_fn=prior)
    for _ in range(args.warmup):
        O.log_prob(lay, theta, what=2, cfg=cfg, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.log_prob(lay, theta, what=2, cfg=cfg, nthreads=cores)
    dt = time.perf_counter() - t0
    value = n_sample * wl.n_ecl * args.steps / dt
    same = n_sample == full
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "walkers_per_step": n_sample, "walkers_of_the_config": full, "eclipses": wl.n_ecl,
                   "n_phase": wl.n_ph, "same_config": same,
                   "note": "CPU restatement of the lfit path (oracle/: literal element x sample sums, the same Newton "
                           "solver), OpenMP over walkers on every host core; lfit itself is not vendored in the "
                           "reference tree.  The metric is a per-evaluation rate: each step is a bounded sample of the "
                           "config's ensemble (%d of %d walkers) so that the run ends within minutes" % (n_sample, full)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "build": "gcc -O3 -march=native -ffp-contract=off" if native else "gcc -O3 (generic x86-64)",
                         "sample": "%d walkers x %d eclipses per step" % (n_sample, wl.n_ecl)},
        "reference_tree_overhead": REFERENCE_TREE_OVERHEAD,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args, rank, local_rank, world):
    import torch
    from lfit_python_b200 import _cabi, mcmc_utils, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl, scaling, n = pick_workload(args, world)
    eng = _cabi.Engine(local_rank, **wl.grid)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    # every rank draws its own shard of the ensemble
    theta_h = wl.walkers(n, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR), seed=2024 + rank)
    theta_d = torch.from_numpy(theta_h).cuda()
    lnp_d = torch.empty(n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 2x the 126 MB L2
    peer = None
    if world > 1:
        # ONE packed buffer per rank -- positions and ln_prob side by side -- gathered once per step: stored
        # straight into every peer's HBM over NVLink by the engine's own kernel (csrc/peer.cuh); NCCL's
        # all-gather where the peers' windows cannot be mapped
        packed = torch.empty(n, wl.ndim + 1, dtype=torch.float64, device="cuda")
        packed[:, : wl.ndim] = theta_d
        gathered = torch.empty(world * n, wl.ndim + 1, dtype=torch.float64, device="cuda")
        if not os.environ.get("LFB_BENCH_NCCL"):
            peer = parallel.PeerExchange(eng, n * (wl.ndim + 1) * 8)
            if not peer.available:
                if rank == 0:
                    print("bench.py: peer exchange unavailable (%s): NCCL all-gather" % peer.why, file=sys.stderr)
                peer = None
    gathered_ptr = [0]
    # a dedicated non-default stream: the C ABI reads a NULL stream as "the handle's own stream",
    # and torch's default stream has handle 0
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        eng.log_prob_device(theta_d.data_ptr(), n, lnp_d.data_ptr(), what=_cabi.LN_PROB, stream=stream)
        if peer is not None:
            gathered_ptr[0] = peer.allgather(theta_d.data_ptr(), wl.ndim, lnp_d.data_ptr(), 1, n, stream)
        elif world > 1:
            packed[:, wl.ndim].copy_(lnp_d)
            dist.all_gather_into_tensor(gathered, packed)

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
    # the K timed steps are queued back to back -- [L2 flush][event][step][event] -- with no host synchronisation in
    # between, so that the ranks stay in step on the device (a host-side pause on one rank would be waited for
    # inside the timed region of all the others)
    for i in range(args.steps):
        flush.fill_(float(i))  # evict L2 between timed iterations (outside the event pair)
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    # per-stage device times of the overlapped pipeline: a few more steps, read back one by one
    kernel_ms, stage_ms = [], []
    for i in range(min(args.steps, 10)):
        flush.fill_(float(i))
        step()
        torch.cuda.synchronize()
        kernel_ms.append(eng.last_kernel_ms())
        stage_ms.append(eng.last_stage_ms())
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    gather_ok = None
    if world > 1:
        # what was gathered is what was computed: this rank's slice holds its own positions and ln_probs, and
        # every rank holds the same gathered buffer
        if peer is not None:
            assert not peer.timed_out(), "a rank never arrived at an exchange"
            gathered = peer.view(gathered_ptr[0], n, wl.ndim + 1).reshape(world * n, wl.ndim + 1).clone()
        mine = gathered[rank * n:(rank + 1) * n]
        gather_ok = bool(torch.equal(mine[:, wl.ndim], lnp_d) and torch.equal(mine[:, : wl.ndim], theta_d))
        chk = torch.nan_to_num(gathered[:, wl.ndim], neginf=-1e300).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        gather_ok = gather_ok and bool((lo == hi).item())
        flag = torch.tensor([1.0 if gather_ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(flag.item() == 1.0)
        assert gather_ok, "the all-gathered positions / log-probs differ from what the ranks computed"

    peer_used = peer is not None
    if peer is not None:
        peer.close()     # (every rank is past its last exchange: the barrier above; the sharded sampler opens its own)
        peer = None
    # end to end through the public host API: pinned H2D of theta + D2H of ln_prob every step
    e2e_steps = args.steps
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # (inputs in page-locked host memory, as the contract asks: the engine copies from it as it stands; pageable
    # arrays are staged through its own pinned buffers at the price of a host memcpy)
    theta_pin = torch.from_numpy(theta_h).pin_memory()
    lnp_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    theta_hp, lnp_hp = theta_pin.numpy(), lnp_pin.numpy()
    eng.log_prob(theta_hp, what=_cabi.LN_PROB, out=lnp_hp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lnp_h = eng.log_prob(theta_hp, what=_cabi.LN_PROB, out=lnp_hp)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert np.array_equal(lnp_h, lnp_d.cpu().numpy())

    # ---- emcee steps/s (one step = two half-steps = one ln_prob per walker of the ensemble) ----
    n_mc = min(max(5, args.steps), 50)
    emcee = {"steps_timed": n_mc}
    # (a) host stretch move (numpy, emcee's call shape) driving the vectorised CUDA log-probability, per GPU
    hs = mcmc_utils.EnsembleSampler(n, wl.ndim, lambda t: eng.log_prob(t, what=_cabi.LN_PROB), vectorize=True,
                                    rng=np.random.default_rng(99 + rank))
    state = hs.run_mcmc(theta_h, 2, store=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    hs.run_mcmc(state[0], n_mc, log_prob0=state[1], store=False)
    mc_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mc_s, op=dist.ReduceOp.MAX)
    emcee["host_sampler_steps_per_s"] = n_mc / float(mc_s.item())
    emcee["host_sampler_note"] = "numpy stretch move + one CUDA ln_prob call per half-step, host buffers, %d walkers per GPU" % n
    emcee["acceptance_fraction"] = float(hs.acceptance_fraction.mean())
    # (b) the ensemble resident in HBM: native stretch-move kernels, nothing crosses PCIe.  One GPU: the
    # config's ensemble on this GPU; N GPUs: ONE ensemble sharded over the ranks, one exchange of packed rows per
    # half-step (weak configs: N x walkers in total; strong configs: the config's total)
    total_walkers = n * world
    if world == 1:
        ds = mcmc_utils.DeviceSampler(eng, n, seed=7)
        ds.set_state(theta_h)
        ds.run(3)
        ds.get_state()
        t0 = time.perf_counter()
        ds.run(n_mc)
        ds.get_state()
        dt = time.perf_counter() - t0
        emcee["device_resident_steps_per_s"] = n_mc / dt
        emcee["sharded_steps_per_s"] = n_mc / dt
        emcee["device_acceptance_fraction"] = float(ds.acceptance_fraction.mean())
        ds.close()
    else:
        # the same start ensemble on every rank: all ranks' shards, gathered
        allpos = gathered[:, : wl.ndim].contiguous().cpu().numpy()
        ss = parallel.ShardedDeviceSampler(eng, total_walkers, seed=7)
        ss.set_state(allpos)
        ss.run(3)
        ss.get_state()
        dist.barrier()
        t0 = time.perf_counter()
        ss.run(n_mc)
        pos_s, lnp_s = ss.get_state()
        dt_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        emcee["sharded_steps_per_s"] = n_mc / float(dt_t.item())
        # every rank ends with the same ensemble
        cs = torch.tensor([float(np.nan_to_num(lnp_s, neginf=-1e300).sum()), float(pos_s.sum())], dtype=torch.float64, device="cuda")
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        emcee["sharded_ranks_agree"] = bool(torch.equal(lo, hi))
        assert emcee["sharded_ranks_agree"], "the ranks of the sharded sampler hold different ensembles"
        emcee["device_acceptance_fraction"] = float((ss.naccepted / max(ss.iterations, 1)).mean())
        ss_exchange = "peer stores over NVLink, csrc/peer.cuh" if ss.exchange == "peer" else "nccl all_gather_into_tensor"
        ss.close()
    emcee["sharded_ensemble_walkers"] = total_walkers
    emcee["sharded_note"] = ("device-resident stretch move (csrc/sampler.cuh), ONE ensemble of %d walkers over %d GPU(s), "
                             "one packed all-gather of [rows, ndim + 2] per half-step%s" % (
                                 total_walkers, world, "" if world == 1 else " (exchange: %s)" % ss_exchange))
    emcee["sharded_lightcurve_evals_per_s"] = emcee["sharded_steps_per_s"] * total_walkers * wl.n_ecl

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
        stages = {k: float(np.mean([d[k] for d in stage_ms])) for k in stage_ms[0]}
        serial = {k: float(np.mean([d[k] for d in serial_stage])) for k in serial_stage[0]}
        trace = {k: float(np.mean([d[k] for d in serial_trace])) for k in serial_trace[0]}
        # the stage-1 launches of the main stream (the donor's tiles are solved beside them on a side stream)
        el_ms = sum(v for k, v in trace.items() if k.startswith("elements_kernel") and "side stream" not in k)
        fx_ms = sum(v for k, v in trace.items() if k.startswith("flux_kernel"))
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
        smem_peak = 148 * 128 * sm_mhz * 1e6 * 1e-9  # GB/s: 128 B per clock per SM
        roof = {"bound": "fp64", "kernel": "elements_kernel<disc, white dwarf, donor, strip> (stage 1: Roche ingress/egress solves)",
                "kernel_ms": el_ms, "kernel_share_of_step": el_ms / serial["total"],
                "stage_ms_serial": serial, "kernel_ms_serial": trace, "stage_ms_overlapped": stages, "pipeline_ms": k_ms,
                "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "DFMA probe on this device, this run "
                "(MEASURED_PEAKS.json has no FP64 vector figure)", "achieved": None, "frac": None, "traffic": None}
        fl = load_flops()
        per_rank = n * wl.n_ecl
        if fl and args.config == 1 and not args.n_ph and not wl.grid:
            pl = fl["per_lightcurve"]
            tf = lambda flop, ms: flop * per_rank / (ms * 1e-3) * 1e-12
            roof["achieved"] = tf(pl["elements_flop"], el_ms)
            roof["frac"] = roof["achieved"] / fp64_peak
            roof["traffic"] = pl["elements_dram_bytes"] * per_rank
            roof["counters"] = {"file": "profiles/r02_flops.json", "source": fl.get("source"), "csrc_sha": fl.get("csrc_sha"),
                                "stale": fl.get("stale"), "flop_per_lightcurve": {"elements": pl["elements_flop"],
                                                                                  "flux": pl["flux_flop"], "all": pl["all_flop"]},
                                "note": "FP64 operations the kernels EXECUTE (ncu, FMA = 2; the FP32 warm-up of the Newton "
                                        "iterations is not counted)"}
            roof["whole_pass"] = {"achieved": tf(pl["all_flop"], k_ms), "frac": tf(pl["all_flop"], k_ms) / fp64_peak}
            if "all_warp_inst" in pl:
                # what actually bounds these kernels: instruction issue (one warp instruction per clock per SM sub-partition)
                n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
                issue_peak = n_sm * 4 * sm_mhz * 1e6 * 1e-9
                gi = lambda winst, ms: winst * per_rank / (ms * 1e-3) * 1e-9
                roof["issue"] = {
                    "unit": "G warp-instructions/s", "peak": issue_peak,
                    "peak_source": "%d SMs x 4 schedulers x %d MHz (the SM clock sampled during the timed region)" % (n_sm, sm_mhz),
                    "elements": {"achieved": gi(pl["elements_warp_inst"], el_ms), "frac": gi(pl["elements_warp_inst"], el_ms) / issue_peak},
                    "flux_kernel": {"achieved": gi(pl["flux_warp_inst"], fx_ms), "frac": gi(pl["flux_warp_inst"], fx_ms) / issue_peak},
                    "whole_pass": {"achieved": gi(pl["all_warp_inst"], k_ms), "frac": gi(pl["all_warp_inst"], k_ms) / issue_peak},
                    "warp_inst_per_lightcurve": {"elements": pl["elements_warp_inst"], "flux": pl["flux_warp_inst"], "all": pl["all_warp_inst"]},
                    "note": "warp instructions the kernels execute (ncu smsp__inst_executed.sum) / CUDA-event time; an FP64 "
                            "instruction holds its issue port for two clocks, so the FP64-heavy kernels saturate below 1.0"}
            roof["kernels"] = {
                "flux_kernel": {
                    "ms": fx_ms, "share_of_step": fx_ms / serial["total"],
                    "fp64": {"achieved": tf(pl["flux_flop"], fx_ms), "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": tf(pl["flux_flop"], fx_ms) / fp64_peak},
                    "hbm": {"achieved": pl["flux_dram_bytes"] * per_rank / (fx_ms * 1e-3) * 1e-9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": pl["flux_dram_bytes"] * per_rank / (fx_ms * 1e-3) * 1e-9 / hbm_peak,
                            "note": "DRAM bytes of the kernel (ncu) / its CUDA-event time"},
                    "shared_memory": {"achieved": pl["flux_smem_wavefronts"] * 128 * per_rank / (fx_ms * 1e-3) * 1e-9,
                                      "peak": smem_peak, "unit": "GB/s",
                                      "frac": pl["flux_smem_wavefronts"] * 128 * per_rank / (fx_ms * 1e-3) * 1e-9 / smem_peak,
                                      "note": "shared-memory wavefronts x 128 B against 128 B/clk/SM x 148 SMs"},
                    "bound": "latency / instruction issue: see profiles/r02_flux_kernel.txt"},
            }
        # HBM side of the whole pass, for the record: theta in, chi-squared out, element and event records
        G = eng.config
        n_wd, n_disc = 4 * G["n_wd_rings"] ** 2, G["n_disc_r"] * G["n_disc_th"]
        alg_bytes = n * wl.n_ecl * (18 * 8 + 8 + 2 * (16 * (n_wd // 2 + n_disc // 2 + G["n_bs"]) + 8 * G["n_bs"] + 32 * 103)
                                    + 2 * 16 * (n_wd + n_disc + G["n_bs"] + 412))
        roof["hbm"] = {"achieved": alg_bytes / (k_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": alg_bytes / (k_ms * 1e-3) * 1e-9 / hbm_peak, "peak_source": hbm_src,
                       "note": "algorithmic bytes of the whole pass: theta in, chi-squared out, element and event records written and read once"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "walkers_per_gpu": n, "walkers_total": n * world, "eclipses": wl.n_ecl,
                       "n_phase": wl.n_ph, "ndim": wl.ndim, "grid": eng.config,
                       "l2": "flushed between timed iterations (256 MB fill)",
                       "parallelism": "walkers sharded, %d rank(s)" % world,
                       "collective": "none" if world == 1 else (
                           ("ONE exchange per step of the packed [walkers, ndim + 1] rows (positions + ln_prob): the engine's "
                            "own kernel stores them into every peer's HBM over NVLink and the ranks meet on device-side "
                            "flags (csrc/peer.cuh); gathered == computed checked") if peer_used else
                           ("ONE nccl all_gather per step of a packed [walkers, ndim + 1] buffer "
                            "(positions + ln_prob); gathered == computed checked"))},
            "ensemble_passes_per_s": args.steps / (total_ms * 1e-3),
            "gather_verified": gather_ok,
            "gp_likelihood": gp,
            "emcee_steps_per_s": emcee["sharded_steps_per_s"],
            "emcee": emcee,
            "e2e": {"value": evals_per_step * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * wl.ndim * 8), "d2h_bytes_per_step": int(n * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "reference_tree_overhead": REFERENCE_TREE_OVERHEAD,
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
    ap.add_argument("--walkers", type=int, default=0, help="override the config's walkers")
    ap.add_argument("--scaling", default="", choices=["", "weak", "strong"], help="default: weak for C1-C3, strong for C4 / C5")
    ap.add_argument("--n-ph", type=int, default=0, help="override phase points per eclipse")
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: walkers per step")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="reference arm: budget of the whole run")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gp", action="store_true", help="skip the Gaussian-process likelihood leg")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
