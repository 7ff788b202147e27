"""numpy restatement of the stretch-move sampler of lfit_python_b200/csrc/sampler.cuh.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): the checker of the CUDA sampler kernels and the
CPU stand-in that lets the sharded-ensemble orchestration run under gloo without a GPU.

What it restates: emcee's stretch move (emcee is third-party, not vendored; call sites
/root/reference/mcmcfit.py:283-288, mcmc_utils.py:114-183) --
    zz = ((a - 1) * u + 1) ** 2 / a;  q = c[rint] - (c[rint] - s) * zz[:, None]
    accept where (ndim - 1) * log(zz) + lnp(q) - lnp(s) > log(u')
with the first and the second half of the ensemble updated in turn (emcee 2.x), and the sampler's random
stream: Philox4x32-10 (Salmon et al. 2011) keyed by the seed, counter (walker row, step lo, step hi, half).
PARITY: emcee's own random stream (numpy MT19937 / PCG64) is not reproduced -- the move's algebra is.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32 on uint32 arrays (counter words c0..c3, key words k0, k1)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & _LO).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & _LO).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(hi, lo):
    """[0, 1) from 53 random bits."""
    x = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def draws(seed, step, half, half_n, a, rows, temp=0):
    """z, partner row (in the other half) and ln u' of the given rows of one half at one step."""
    rows = np.asarray(rows, dtype=np.uint32)
    seed, step = int(seed), int(step)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    tag = (int(half) | (int(temp) << 1)) & 0xFFFFFFFF
    r = philox4x32_10(rows, np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF), np.uint32(tag), k0, k1)
    q = philox4x32_10(rows, np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF),
                      np.uint32(tag | 0x80000000), k0, k1)
    t = (a - 1.0) * u01(r[0], r[1]) + 1.0
    z = t * t / a
    partner = ((r[2].astype(np.uint64) * np.uint64(half_n)) >> np.uint64(32)).astype(np.int64)
    with np.errstate(divide="ignore"):
        lnu = np.log(u01(q[0], q[1]))
    return z, partner, lnu


def shard_bounds(n, rank, world):
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class StretchOracle:
    """The sampler of csrc/sampler.cuh on the host: same draws, same arithmetic, same call shape
    (set_state / run / half_begin / half_end / get_state)."""

    def __init__(self, log_prob_fn, nwalkers, ndim, a=2.0, seed=0):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("need an even number of walkers, at least twice the number of dimensions")
        self.fn, self.n, self.half, self.ndim, self.a, self.seed = log_prob_fn, nwalkers, nwalkers // 2, ndim, float(a), seed
        self.step = 0
        self.pos = self.lnp = None
        self.naccepted = np.zeros(nwalkers, dtype=np.int64)
        self.chain = []

    def set_state(self, pos, lnp=None):
        self.pos = np.array(pos, dtype=np.float64, copy=True)
        self.lnp = np.asarray(self.fn(self.pos), dtype=np.float64).copy() if lnp is None else np.array(lnp, dtype=np.float64)

    def half_begin(self, half, lo, hi):
        """packed[hi - lo][ndim + 2] = (position, ln_prob, accepted) of rows [lo, hi) of one half."""
        rows = np.arange(lo, hi)
        z, partner, lnu = draws(self.seed, self.step, half, self.half, self.a, rows)
        k = half * self.half + rows
        j = (1 - half) * self.half + partner
        s, c = self.pos[k], self.pos[j]
        prop = c - (c - s) * z[:, None]
        new = np.asarray(self.fn(prop), dtype=np.float64) if hi > lo else np.empty(0)
        with np.errstate(invalid="ignore"):
            acc = ((self.ndim - 1.0) * np.log(z) + new - self.lnp[k]) > lnu
        packed = np.empty((hi - lo, self.ndim + 2))
        packed[:, :self.ndim] = np.where(acc[:, None], prop, s)
        packed[:, self.ndim] = np.where(acc, new, self.lnp[k])
        packed[:, self.ndim + 1] = acc
        return packed

    def half_end(self, half, gathered, world, slot):
        gathered = np.asarray(gathered).reshape(world, slot, self.ndim + 2)
        for r in range(world):
            lo, hi = shard_bounds(self.half, r, world)
            rows = gathered[r, : hi - lo]
            k = half * self.half + np.arange(lo, hi)
            self.pos[k] = rows[:, :self.ndim]
            self.lnp[k] = rows[:, self.ndim]
            self.naccepted[k] += rows[:, self.ndim + 1].astype(np.int64)
        if half == 1:
            self.step += 1

    def run(self, nsteps, record=False):
        for _ in range(nsteps):
            for half in (0, 1):
                packed = self.half_begin(half, 0, self.half)
                self.half_end(half, packed[None], 1, self.half)
            if record:
                self.chain.append(np.concatenate([self.pos, self.lnp[:, None]], axis=1))

    def get_state(self):
        return self.pos.copy(), self.lnp.copy()
