"""Walker sharding over ranks (world_size 2, gloo on the CPU): the all-gathered result equals the
single-process one bit for bit, for even and ragged shard sizes."""
import os

import numpy as np
import pytest

from lfit_python_b200.parallel import ShardedLogProb, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _fn(theta):
    return np.sin(theta).sum(axis=1) - 0.5 * (theta ** 2).sum(axis=1)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def local(rows):
        calls.append(rows.shape[0])
        return _fn(rows)

    sharded = ShardedLogProb(local)
    rng = np.random.default_rng(4)  # the same stream on every rank
    res = {}
    for n in (8, 9, 1):
        theta = rng.standard_normal((n, 5))
        res[n] = sharded(theta)
        assert np.array_equal(res[n], _fn(theta))
    # a short sampler run driven by the sharded log-probability: identical chains on all ranks
    from lfit_python_b200 import mcmc_utils as utils
    s = utils.EnsembleSampler(12, 5, sharded, vectorize=True, rng=np.random.default_rng(7))
    pos, lnp, _ = s.run_mcmc(np.random.default_rng(8).standard_normal((12, 5)), 10)
    np.save(os.path.join(out_dir, "pos_%d.npy" % rank), pos)
    np.save(os.path.join(out_dir, "calls_%d.npy" % rank), np.asarray(calls))
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "pos_0.npy"), np.load(tmp_path / "pos_1.npy")
    assert np.array_equal(p0, p1)
    from lfit_python_b200 import mcmc_utils as utils
    ref = utils.EnsembleSampler(12, 5, _fn, vectorize=True, rng=np.random.default_rng(7))
    pos, _, _ = ref.run_mcmc(np.random.default_rng(8).standard_normal((12, 5)), 10)
    assert np.array_equal(pos, p0)  # N ranks == 1 process, bit for bit
    c0 = np.load(tmp_path / "calls_0.npy")
    assert c0[0] == 4 and c0[1] == 5 and c0[2] == 1   # rank 0's share of 8, 9 and 1 rows


# ---- the device-resident sharded sampler's orchestration, with the numpy stand-in doing each rank's work ----
def _gauss(theta):
    return -0.5 * np.sum(np.atleast_2d(theta) ** 2 / np.array([1.0, 4.0, 0.25]), axis=1)


def _sharded_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from lfit_python_b200.parallel import ShardedDeviceSampler
    from oracle.stretch import StretchOracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def fn(theta):
        calls.append(len(theta))
        return _gauss(theta)

    for nw in (14, 16):          # ragged (7 rows of a half over 2 ranks) and even shards
        ops = StretchOracle(fn, nw, 3, seed=21)
        smp = ShardedDeviceSampler(None, nw, ops=ops)
        smp.set_state(np.random.default_rng(3).standard_normal((nw, 3)))
        calls.clear()
        smp.run(25)
        pos, lnp = smp.get_state()
        np.save(os.path.join(out_dir, "spos_%d_%d.npy" % (nw, rank)), pos)
        np.save(os.path.join(out_dir, "sacc_%d_%d.npy" % (nw, rank)), smp.naccepted)
        lo, hi = smp.lo, smp.hi
        assert calls == [hi - lo] * 50        # this rank evaluated only its slice of each half-step
    dist.destroy_process_group()


def test_sharded_device_sampler_orchestration_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    from oracle.stretch import StretchOracle
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for nw in (14, 16):
        one = StretchOracle(_gauss, nw, 3, seed=21)
        one.set_state(np.random.default_rng(3).standard_normal((nw, 3)))
        one.run(25)
        for r in range(2):
            assert np.array_equal(np.load(tmp_path / ("spos_%d_%d.npy" % (nw, r))), one.pos)   # N ranks == 1, bit for bit
            assert np.array_equal(np.load(tmp_path / ("sacc_%d_%d.npy" % (nw, r))), one.naccepted)
        assert one.naccepted.sum() > 0
