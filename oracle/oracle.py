"""ctypes wrapper around the CPU oracle (oracle/lfit_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
CPU legs of bench.py -- never by lfit_python_b200/.  PARITY UNPINNED: the
reference's arithmetic (`lfit`, `trm.roche`; /root/reference/CVModel.py:13,15)
is not vendored, so this oracle is the repo's own restatement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblfit_oracle.so")

NPAR = 18
FLAG_INCL, SKIP_WD, SKIP_DISC, SKIP_BS, SKIP_DONOR = 1, 2, 4, 8, 16
SOLVER_ROBUST, SOLVER_NEWTON = 0, 1
PRIOR_TYPES = {"gauss": 0, "gaussPos": 1, "uniform": 2, "log_uniform": 3, "mod_jeff": 4}


class Config(C.Structure):
    _fields_ = [
        ("n_wd_rings", C.c_int), ("n_disc_r", C.c_int), ("n_disc_th", C.c_int), ("n_bs", C.c_int),
        ("n_donor_th", C.c_int), ("n_quad", C.c_int), ("donor_ulimb", C.c_double),
        ("donor_gdexp", C.c_double), ("solver", C.c_int),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("n_ecl", C.c_int), ("npars", C.c_int),
        ("gather", C.POINTER(C.c_int)), ("consts", C.POINTER(C.c_double)),
        ("n_prior", C.c_int), ("prior_src", C.POINTER(C.c_int)), ("prior_type", C.POINTER(C.c_int)),
        ("prior_p1", C.POINTER(C.c_double)), ("prior_p2", C.POINTER(C.c_double)),
        ("prior_norm", C.POINTER(C.c_double)), ("prior_isvar", C.POINTER(C.c_int)),
        ("lc_off", C.POINTER(C.c_longlong)), ("lc_phase", C.POINTER(C.c_double)),
        ("lc_width", C.POINTER(C.c_double)), ("lc_y", C.POINTER(C.c_double)),
        ("lc_ye", C.POINTER(C.c_double)),
    ]


def build(force=False):
    """Compile the oracle with the Makefile next to this file."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("lfit_oracle.c", "lfit_oracle.h", "roche_core.h")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_NATIVE_PATH = os.path.join(_HERE, "_native", "liblfit_oracle_native.so")
_use_native = False


def use_native():
    """bench.py's CPU legs: compile the oracle on THIS machine with -O3 -march=native (no fused multiply-add
    contraction, so the numbers do not change) and use that build from now on.  Returns False, keeping the
    portable build, if the compiler is missing.  Must be called before the first lib()."""
    global _use_native
    if _lib is not None:
        return _use_native
    try:
        os.makedirs(os.path.dirname(_NATIVE_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-fopenmp", "-std=c11", "-ffp-contract=off",
                               "-shared", "-o", _NATIVE_PATH, os.path.join(_HERE, "lfit_oracle.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _use_native = True
    except Exception:
        _use_native = False
    return _use_native


def lib():
    global _lib
    if _lib is None:
        if not _use_native:
            build()
        _lib = C.CDLL(_NATIVE_PATH if _use_native else _LIB_PATH)
        dp = C.POINTER(C.c_double)
        _lib.lfo_default_config.argtypes = [C.POINTER(Config)]
        _lib.lfo_roche_xl1.argtypes = [C.c_double, dp]
        _lib.lfo_roche_findphi.argtypes = [C.c_double, C.c_double, dp]
        _lib.lfo_roche_findi.argtypes = [C.c_double, C.c_double, dp]
        _lib.lfo_roche_bspot.argtypes = [C.c_double, C.c_double, dp]
        _lib.lfo_roche_ingress_egress.argtypes = [C.c_double, C.c_double, dp, C.c_double, C.c_double,
                                                  C.c_int, dp, dp]
        _lib.lfo_calc_flux.argtypes = [C.POINTER(Config), dp, C.c_int, C.c_int, C.c_int, dp, dp,
                                       dp, dp, dp, dp, dp]
        _lib.lfo_chisq.argtypes = [C.POINTER(Config), dp, C.c_int, C.c_int, dp, dp, dp, dp]
        _lib.lfo_chisq.restype = C.c_double
        _lib.lfo_prior_ln_prob.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        _lib.lfo_prior_ln_prob.restype = C.c_double
        _lib.lfo_log_prob.argtypes = [C.POINTER(Config), C.POINTER(Layout), C.c_int, C.c_longlong, dp,
                                      dp, dp, C.c_int]
        _lib.lfo_max_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def config(**kw):
    cfg = Config()
    lib().lfo_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


class RocheError(Exception):
    pass


def xl1(q):
    out = C.c_double()
    if lib().lfo_roche_xl1(q, C.byref(out)):
        raise RocheError("xl1: q must be > 0")
    return out.value


def findphi(q, incl_deg):
    out = C.c_double()
    if lib().lfo_roche_findphi(q, incl_deg, C.byref(out)):
        raise RocheError("findphi failed")
    return out.value


def findi(q, dphi):
    out = C.c_double()
    if lib().lfo_roche_findi(q, dphi, C.byref(out)):
        raise RocheError("findi: no inclination gives that eclipse width")
    return out.value


def bspot(q, rad):
    out = (C.c_double * 4)()
    if lib().lfo_roche_bspot(q, rad, out):
        raise RocheError("bspot: stream does not reach that radius")
    return tuple(out)


def ingress_egress(q, incl_deg, p0, xi=0.0, eta=0.0, solver=SOLVER_NEWTON):
    a, b = C.c_double(), C.c_double()
    p = (C.c_double * 3)(*p0)
    r = lib().lfo_roche_ingress_egress(q, incl_deg, p, xi, eta, solver, C.byref(a), C.byref(b))
    if r < 0:
        raise RocheError("bad q")
    return (a.value, b.value) if r else None


def calc_flux(pars, phase, width=None, cfg=None, flags=0, components=False):
    """lfit.CV(pars).calcFlux(pars, phase, width) on the CPU oracle."""
    cfg = cfg or config()
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    phase = np.ascontiguousarray(phase, dtype=np.float64)
    n = phase.shape[0]
    if width is None:
        width = np.zeros(n)
    width = np.ascontiguousarray(np.broadcast_to(np.asarray(width, dtype=np.float64), (n,)))
    tot = np.empty(n)
    comps = [np.empty(n) for _ in range(4)] if components else [None] * 4
    st = lib().lfo_calc_flux(C.byref(cfg), _dp(pars), pars.shape[0], flags, n, _dp(phase), _dp(width),
                             _dp(tot), *[_dp(c) for c in comps])
    if components:
        return st, tot, comps
    return st, tot


def chisq(pars, phase, width, y, ye, cfg=None):
    cfg = cfg or config()
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (phase, width, y, ye)]
    return lib().lfo_chisq(C.byref(cfg), _dp(pars), pars.shape[0], arrs[0].shape[0], *[_dp(a) for a in arrs])


def prior_ln_prob(ptype, p1, p2, norm, val):
    return lib().lfo_prior_ln_prob(PRIOR_TYPES[ptype] if isinstance(ptype, str) else ptype, p1, p2, norm, val)


class FlatLayout:
    """Owns the numpy arrays behind an lfo_layout (same fields as the C-ABI's set_* calls)."""

    def __init__(self, ndim, npars, gather, consts, prior_src, prior_type, prior_p1, prior_p2,
                 prior_norm, prior_isvar, lc_off, lc_phase, lc_width, lc_y, lc_ye):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.gather = i32(gather).reshape(-1, NPAR)
        self.consts = f64(consts) if len(consts) else np.zeros(1)
        self.prior_src, self.prior_type, self.prior_isvar = i32(prior_src), i32(prior_type), i32(prior_isvar)
        self.prior_p1, self.prior_p2, self.prior_norm = f64(prior_p1), f64(prior_p2), f64(prior_norm)
        self.lc_off = np.ascontiguousarray(lc_off, dtype=np.int64)
        self.lc_phase, self.lc_width, self.lc_y, self.lc_ye = f64(lc_phase), f64(lc_width), f64(lc_y), f64(lc_ye)
        self.ndim, self.npars, self.n_ecl = int(ndim), int(npars), self.gather.shape[0]
        L = Layout()
        L.ndim, L.n_ecl, L.npars = self.ndim, self.n_ecl, self.npars
        L.gather, L.consts = _ip(self.gather), _dp(self.consts)
        L.n_prior = self.prior_src.shape[0]
        L.prior_src, L.prior_type, L.prior_isvar = _ip(self.prior_src), _ip(self.prior_type), _ip(self.prior_isvar)
        L.prior_p1, L.prior_p2, L.prior_norm = _dp(self.prior_p1), _dp(self.prior_p2), _dp(self.prior_norm)
        L.lc_off = self.lc_off.ctypes.data_as(C.POINTER(C.c_longlong))
        L.lc_phase, L.lc_width, L.lc_y, L.lc_ye = _dp(self.lc_phase), _dp(self.lc_width), _dp(self.lc_y), _dp(self.lc_ye)
        self.c = L


def log_prob(layout, theta, what=2, cfg=None, nthreads=0, return_chisq=False):
    cfg = cfg or config()
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    n = theta.shape[0]
    assert theta.shape[1] == layout.ndim
    out = np.empty(n)
    chis = np.empty((n, layout.n_ecl)) if return_chisq else None
    lib().lfo_log_prob(C.byref(cfg), C.byref(layout.c), what, n, _dp(theta), _dp(out), _dp(chis), nthreads)
    return (out, chis) if return_chisq else out


def max_threads():
    return lib().lfo_max_threads()


# ---------------------------------------------------------------------------------------------
# Gaussian-process likelihood (/root/reference/CVModel.py:527-696).
# george is not installed and its HODLR solver only approximates K^-1 (default tol 0.1), so the
# oracle is the exact dense statement of the same kernel: build K, Cholesky, ln L.  PARITY
# UNPINNED against george for the same reason as the rest of this file.

GP_WHITE_NOISE = 1.25e-12  # george.GP default white_noise = ln(TINY), TINY = 1.25e-12


def wdphases(q, incl_deg, r1, ntheta=10):
    """Third and fourth contact phases of the white dwarf (trm.roche.wdphases; call site
    CVModel.py:562, which passes rwd as r1): the earliest and the latest egress phase over
    ntheta points on the limb of a disc of radius r1 (units of a) facing the observer."""
    eg = []
    for k in range(int(ntheta)):
        a = 2.0 * np.pi * k / ntheta
        res = ingress_egress(q, incl_deg, (0.0, 0.0, 0.0), xi=r1 * np.cos(a), eta=r1 * np.sin(a), solver=SOLVER_ROBUST)
        if res is None:
            raise RocheError("wdphases: a limb point is never eclipsed")
        eg.append(res[1])
    return min(eg), max(eg)


def gp_dist_cp(q, dphi, rwd, ntheta=10):
    """Distance of the GP change points from mid-eclipse (CVModel.py:558-568)."""
    inc = findi(q, dphi)
    phi3, phi4 = wdphases(q, inc, rwd, ntheta)
    return (dphi + (phi4 - phi3)) / 2.0


def gp_changepoints(x, dist_cp, phi0):
    """[[egress of the previous eclipse, ingress of this one], ...] (CVModel.py:580-599)."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = int(np.floor(x.min())), int(np.ceil(x.max()))
    return [[(e - 1) + dist_cp + phi0, e - dist_cp + phi0] for e in range(lo, hi + 1)
            if e > x.min() and e < 1 + x.max()]


def gp_kernel_matrix(x, ye, ampin, ampout, tau, gaps):
    """ampin * Matern32(tau) + sum over gaps of ampout * Matern32(tau, block=gap), plus the
    diagonal george adds in compute(x, yerr) (CVModel.py:636-645,687)."""
    x = np.asarray(x, dtype=np.float64)
    r = np.sqrt(3.0 * (x[:, None] - x[None, :]) ** 2 / tau)
    m32 = (1.0 + r) * np.exp(-r)
    K = ampin * m32
    for a, b in gaps:
        ins = (x >= a) & (x <= b)
        K = K + ampout * m32 * (ins[:, None] & ins[None, :])
    return K + np.diag(np.asarray(ye, dtype=np.float64) ** 2 + GP_WHITE_NOISE)


def gp_log_like(x, ye, resid, ampin, ampout, tau, gaps):
    """george.GP.log_likelihood(resid, quiet=True) for that kernel, evaluated exactly."""
    resid = np.asarray(resid, dtype=np.float64)
    if not np.all(np.isfinite(resid)):
        return -np.inf
    if not (ampin > 0 and ampout > 0 and tau > 0 and np.isfinite(ampin + ampout + tau)):
        return -np.inf
    K = gp_kernel_matrix(x, ye, ampin, ampout, tau, gaps)
    try:
        L = np.linalg.cholesky(K)
    except np.linalg.LinAlgError:
        return -np.inf
    z = np.linalg.solve(L, resid)
    out = -0.5 * (z @ z) - np.log(np.diag(L)).sum() - 0.5 * len(resid) * np.log(2.0 * np.pi)
    return float(out) if np.isfinite(out) else -np.inf
