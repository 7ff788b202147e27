"""Walkers sharded over the GPUs of one box: one process per GPU.

The reference spreads walkers over a multiprocessing.Pool (mcmcfit.py:273-288).  Here every
rank holds the whole ensemble and the same random stream, evaluates a contiguous slice of each
half-step's proposals on its own GPU, and an all-gather hands every rank all log-probabilities: the
engine's own exchange through NVLink peer memory (PeerExchange; device resident), NCCL on device
tensors where the peers' windows cannot be mapped, gloo on host tensors in the CPU tests.  Walkers are
independent, so an N-GPU run returns bit-for-bit what one GPU returns.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n rows for `rank` of `world`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedLogProb:
    """log_prob_fn(theta[n, ndim]) -> (n,) evaluated cooperatively by all ranks of a
    torch.distributed group.  `local_fn(theta_rows)` evaluates rows on this rank."""

    def __init__(self, local_fn, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.local_fn, self.group = local_fn, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device  # None: host tensors (gloo); else a cuda device (nccl)

    def __call__(self, theta):
        torch, dist = self.torch, self.dist
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n = theta.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        local = np.asarray(self.local_fn(theta[lo:hi]), dtype=np.float64) if hi > lo else np.empty(0)
        # equal-sized slots so that one all_gather_into_tensor serves ragged shards
        slot = -(-n // self.world)
        buf = torch.full((slot,), float("nan"), dtype=torch.float64)
        buf[: hi - lo] = torch.from_numpy(local)
        if self.device is not None:
            buf = buf.to(self.device)
        out = torch.empty(self.world * slot, dtype=torch.float64, device=buf.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.cpu().numpy().reshape(self.world, slot)
        res = np.empty(n)
        for r in range(self.world):
            a, b = shard_bounds(n, r, self.world)
            res[a:b] = out[r, : b - a]
        return res


class _DevicePointer:
    """A raw device pointer dressed as a CUDA array (for torch.as_tensor)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """All-gather of float64 rows between the ranks of one box through NVLink peer memory (C ABI
    lfb_peer_*, csrc/peer.cuh): every rank's kernel stores its rows straight into every peer's
    window and they meet on device-side flags -- no library collective on the data path.  The 64-byte
    CUDA IPC handles of the windows travel once, at set-up, through torch.distributed.

    `available` is False (and `why` says so) where the windows cannot be mapped, e.g. without CUDA
    IPC between the ranks' processes; callers then keep the NCCL all-gather."""

    def __init__(self, engine, slot_bytes, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.engine, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slot_bytes = (int(slot_bytes) + 7) & ~7
        self.available, self.why = False, ""
        lib, h = engine._lib, engine._h
        buf = C.create_string_buffer(64)
        rc = lib.lfb_peer_create(h, self.rank, self.world, self.slot_bytes, buf)
        mine = bytes(buf.raw) if rc == 0 else None
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        ok = all(hd is not None for hd in handles)
        if ok:
            ok = lib.lfb_peer_connect(h, b"".join(handles)) == 0
        if not ok:
            self.why = lib.lfb_last_error(h).decode()
        # everybody or nobody: a rank that could not map its peers must not leave the others waiting
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=group)
        self.available = all(flags)
        if not self.available:
            lib.lfb_peer_destroy(h)
            if not self.why:
                self.why = "a peer could not map the windows"
        dist.barrier(group=group)

    def allgather(self, a_ptr, ca, b_ptr, cb, rows, stream):
        """Pack [a row | b row] and exchange; returns the device pointer of gathered[world][slot_bytes]."""
        import ctypes as C
        out = C.c_void_p()
        self.engine._check(self.engine._lib.lfb_peer_allgather(self.engine._h, a_ptr, int(ca), b_ptr, int(cb), int(rows),
                                                               C.byref(out), stream), "lfb_peer_allgather")
        return out.value

    def view(self, ptr, rows_per_slot, cols):
        """gathered as a torch tensor [world, rows_per_slot, cols] over the window (no copy)."""
        import torch
        assert rows_per_slot * cols * 8 == self.slot_bytes
        return torch.as_tensor(_DevicePointer(ptr, (self.world, rows_per_slot, cols)), device=torch.device("cuda", self.engine.device))

    def timed_out(self):
        import ctypes as C
        v = C.c_int(0)
        self.engine._check(self.engine._lib.lfb_peer_status(self.engine._h, C.byref(v)), "lfb_peer_status")
        return bool(v.value)

    def close(self):
        if self.available:
            self.engine._lib.lfb_peer_destroy(self.engine._h)
            self.available = False


class ShardedDeviceSampler:
    """The stretch move over an ensemble sharded across the GPUs of one box, device resident
    (replaces the multiprocessing.Pool of /root/reference/mcmcfit.py:273-288).

    Every rank holds the whole ensemble in its HBM.  Per half-step a rank proposes, evaluates and
    accepts its contiguous slice of the half (lfb_sampler_half_begin) into a packed
    [rows, ndim + 2] buffer (new position, ln_prob, accepted); ONE all-gather of the packed rows
    (PeerExchange: the engine's own kernel over NVLink peer memory; NCCL on the device buffers as the
    fallback; gloo on host tensors in the CPU tests) hands every rank
    the whole half, and lfb_sampler_half_end writes it into the replicated ensemble.  Nothing touches
    the host.  The draws are counter based (Philox keyed by seed, counted by step / half / walker), so
    N ranks follow the 1-GPU chain bit for bit.

    `ops` does the per-rank work: by default the CUDA sampler of `engine`; tests pass a host stand-in
    with the same four methods (set_state, half_begin, half_end, get_state)."""

    def __init__(self, engine, nwalkers, a=2.0, seed=0, what=2, group=None, ops=None, exchange="peer"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nwalkers, self.half = int(nwalkers), int(nwalkers) // 2
        self.slot = -(-self.half // self.world)      # equal-sized slots serve ragged shards
        self.lo, self.hi = shard_bounds(self.half, self.rank, self.world)
        self.iterations = 0
        if ops is not None:
            self.ops, self.cuda = ops, False
            self.ndim = ops.ndim
            self._packed = torch.zeros((self.slot, self.ndim + 2), dtype=torch.float64)
            self._gathered = torch.zeros((self.world, self.slot, self.ndim + 2), dtype=torch.float64)
            return
        from .mcmc_utils import DeviceSampler
        self.ops, self.cuda = DeviceSampler(engine, nwalkers, a=a, seed=seed, what=what), True
        self.ndim = self.ops.ndim
        self.device = torch.device("cuda", self.ops.engine.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._packed = torch.zeros((self.slot, self.ndim + 2), dtype=torch.float64, device=self.device)
        self._gathered = torch.zeros((self.world, self.slot, self.ndim + 2), dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)     # the fills ran on torch's stream, the sampler has its own
        # the exchange: peer stores over NVLink (csrc/peer.cuh) where the windows can be mapped, else NCCL
        self.peer = None
        if exchange == "peer" and self.world > 1:
            px = PeerExchange(engine, self.slot * (self.ndim + 2) * 8, group=group)
            if px.available and px.slot_bytes == self.slot * (self.ndim + 2) * 8:
                self.peer = px
            else:
                px.close()
        self.exchange = "peer" if self.peer is not None else "nccl"

    def close(self):
        if self.cuda:
            if self.peer is not None:
                self.stream.synchronize()
                self.peer.close()
            self.ops.close()

    def set_state(self, pos, lnp=None):
        self.ops.set_state(pos, lnp)

    def get_state(self):
        if self.cuda:
            self.stream.synchronize()
            if self.peer is not None and self.peer.timed_out():
                raise RuntimeError("a rank never arrived at an exchange: the ensembles of the ranks have diverged")
        return self.ops.get_state()

    @property
    def naccepted(self):
        if self.cuda:
            self.stream.synchronize()
        return self.ops.naccepted

    def set_chain(self, steps):
        self.ops.set_chain(steps)

    def read_chain(self):
        if self.cuda:
            self.stream.synchronize()
        return self.ops.read_chain()

    def _half_step(self, half):
        dist = self.dist
        if self.cuda:
            lib, s, st = self.ops._lib, self.ops._s, self.stream.cuda_stream
            check = self.ops.engine._check
            check(lib.lfb_sampler_half_begin(s, half, self.lo, self.hi, self._packed.data_ptr(), st), "lfb_sampler_half_begin")
            if self.peer is not None:
                gathered = self.peer.allgather(self._packed.data_ptr(), self.ndim + 2, None, 0, self.slot, st)
            else:
                dist.all_gather_into_tensor(self._gathered, self._packed, group=self.group)
                gathered = self._gathered.data_ptr()
            check(lib.lfb_sampler_half_end(s, half, gathered, self.world, self.slot, st), "lfb_sampler_half_end")
        else:
            rows = self.ops.half_begin(half, self.lo, self.hi)
            self._packed[: self.hi - self.lo] = self.torch.from_numpy(rows)
            dist.all_gather_into_tensor(self._gathered.view(-1), self._packed.view(-1), group=self.group)
            self.ops.half_end(half, self._gathered.numpy(), self.world, self.slot)

    def run(self, nsteps):
        """nsteps full steps; asynchronous on the sampler's stream (get_state synchronises)."""
        if self.cuda:
            with self.torch.cuda.stream(self.stream):
                for _ in range(nsteps):
                    self._half_step(0)
                    self._half_step(1)
        else:
            for _ in range(nsteps):
                self._half_step(0)
                self._half_step(1)
        self.iterations += nsteps
