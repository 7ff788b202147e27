"""GPU, two or more devices: walkers sharded over NCCL ranks give bit for bit what one GPU gives.

Replaces the reference's multiprocessing.Pool fan-out (/root/reference/mcmcfit.py:273-288).  Skipped on a
one-GPU box (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from lfit_python_b200 import _cabi, workloads
    from lfit_python_b200.parallel import ShardedLogProb, ShardedDeviceSampler
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = _cabi.Engine(rank)
    wl = workloads.config(1, n_ph=500)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(1001, scatter=0.05, seed=5)      # ragged shards; some walkers outside the priors
    sharded = ShardedLogProb(lambda rows: eng.log_prob(rows), device=torch.device("cuda", rank))
    res = sharded(theta)
    np.save(os.path.join(out_dir, "lnp_%d.npy" % rank), res)
    # the device-resident sharded stretch move: every rank ends with the same ensemble
    theta0 = wl.walkers(256, scatter=0.02, seed=6, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR))
    smp = ShardedDeviceSampler(eng, 256, seed=11)
    smp.set_state(theta0)
    smp.run(6)
    pos, lnp = smp.get_state()
    np.save(os.path.join(out_dir, "pos_%d.npy" % rank), pos)
    np.save(os.path.join(out_dir, "slnp_%d.npy" % rank), lnp)
    np.save(os.path.join(out_dir, "acc_%d.npy" % rank), smp.naccepted)
    with open(os.path.join(out_dir, "exchange_%d.txt" % rank), "w") as f:
        f.write(smp.exchange)
    smp.close()
    # the same chain with the library collective: the exchange cannot change a bit
    smp = ShardedDeviceSampler(eng, 256, seed=11, exchange="nccl")
    smp.set_state(theta0)
    smp.run(6)
    pos2, lnp2 = smp.get_state()
    assert smp.exchange == "nccl" and np.array_equal(pos, pos2) and np.array_equal(lnp, lnp2)
    smp.close()
    # the exchange itself: rows [a | b] of every rank, many rounds back to back (the two buffers of a window
    # alternate; a rank may run one exchange ahead of its peers), against NCCL's all-gather of the same rows
    from lfit_python_b200.parallel import PeerExchange
    rows, ca, cb = 37, 5, 2
    px = PeerExchange(eng, rows * (ca + cb) * 8)
    if px.available:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for it in range(40):
                a = torch.full((rows, ca), float(rank * 1000 + it), dtype=torch.float64, device="cuda") + torch.arange(ca, device="cuda")
                b = -torch.arange(rows * cb, dtype=torch.float64, device="cuda").reshape(rows, cb) - it
                ptr = px.allgather(a.data_ptr(), ca, b.data_ptr(), cb, rows, st.cuda_stream)
                got = px.view(ptr, rows, ca + cb).clone()
                want = torch.empty(world, rows, ca + cb, dtype=torch.float64, device="cuda")
                dist.all_gather_into_tensor(want, torch.cat([a, b], dim=1).contiguous())
                assert torch.equal(got, want), "peer exchange differs from NCCL at round %d" % it
                if rank == 0 and it % 7 == 0:
                    torch.cuda._sleep(20_000_000)     # a slow rank: its peers run ahead as far as the protocol lets them
        st.synchronize()
        assert not px.timed_out()
        px.close()
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def test_two_rank_nccl_equals_one_gpu(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from lfit_python_b200 import _cabi, workloads
    from lfit_python_b200.mcmc_utils import DeviceSampler
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    eng = _cabi.Engine(0)
    wl = workloads.config(1, n_ph=500)
    wl.make_data(lambda p, x, w: eng.calc_flux(p, x, w))
    wl.apply(eng)
    theta = wl.walkers(1001, scatter=0.05, seed=5)
    one = eng.log_prob(theta)
    assert (~np.isfinite(one)).any() and np.isfinite(one).any()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("lnp_%d.npy" % r)), one)
    # the sampler: N ranks == 1 GPU, positions, log-probabilities and acceptance counts
    theta0 = wl.walkers(256, scatter=0.02, seed=6, ln_prior_fn=lambda t: eng.log_prob(t, what=_cabi.LN_PRIOR))
    smp = DeviceSampler(eng, 256, seed=11)
    smp.set_state(theta0)
    smp.run(6)
    pos, lnp = smp.get_state()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("pos_%d.npy" % r)), pos)
        assert np.array_equal(np.load(tmp_path / ("slnp_%d.npy" % r)), lnp)
        assert np.array_equal(np.load(tmp_path / ("acc_%d.npy" % r)), smp.naccepted)
    smp.close()
    eng.close()
    # (informative) which exchange the sharded sampler used on this box
    print("exchange:", [open(tmp_path / ("exchange_%d.txt" % r)).read() for r in range(2)])
