"""ctypes binding of include/lfit_b200.h -- the only way Python reaches the CUDA engine.

There is no CPU fallback: if the shared library is missing or no CUDA device is
present every entry point raises.  The reference binds the same path through
Cython (`import lfit`, /root/reference/CVModel.py:13) and the `trm.roche`
C extension (CVModel.py:15).
"""
import ctypes as C
import os

import numpy as np

from . import _build

NPAR = 18
LN_PRIOR, LN_LIKE, LN_PROB = 0, 1, 2
FLAG_INCL, FLAG_SKIP_WD, FLAG_SKIP_DISC, FLAG_SKIP_BS, FLAG_SKIP_DONOR = 1, 2, 4, 8, 16
ROCHE_XL1, ROCHE_FINDPHI, ROCHE_FINDI, ROCHE_BSPOT, ROCHE_ANGLE = 0, 1, 2, 3, 4
PRIOR_CODES = {"gauss": 0, "gaussPos": 1, "uniform": 2, "log_uniform": 3, "mod_jeff": 4}

EXPORTS = (
    "lfb_create", "lfb_destroy", "lfb_last_error", "lfb_get_config", "lfb_set_layout",
    "lfb_set_priors", "lfb_set_lightcurves", "lfb_log_prob", "lfb_calc_flux", "lfb_roche",
    "lfb_launch_count", "lfb_last_kernel_ms", "lfb_last_stage_ms", "lfb_measure_fp64_peak",
    "lfb_set_trace", "lfb_last_trace_ms", "lfb_set_gp", "lfb_gp_loglike", "lfb_wdphases", "lfb_ingress_egress",
    "lfb_sampler_create", "lfb_sampler_destroy", "lfb_sampler_set_state", "lfb_sampler_run", "lfb_sampler_half_begin",
    "lfb_sampler_half_end", "lfb_sampler_packed", "lfb_sampler_positions", "lfb_sampler_log_prob",
    "lfb_sampler_get_state", "lfb_sampler_set_chain", "lfb_sampler_read_chain", "lfb_stretch_draws",
    "lfb_chain_format", "lfb_chain_append", "lfb_robust_calls",
    "lfb_peer_create", "lfb_peer_connect", "lfb_peer_allgather", "lfb_peer_status", "lfb_peer_destroy",
)
TRACE_KERNELS = ("walker_kernel", "jobcheck_kernel", "elements_kernel<1> disc", "elements_kernel<0> white dwarf",
                 "elements_kernel<3> donor (side stream)", "donor_table_kernel (side stream)", "prep_kernel", "positions_kernel", "elements_kernel<2> strip",
                 "prep_strip_kernel + positions_kernel<1>",
                 "flux_kernel", "gp_kernel", "finish_kernel", "stream_kernel (side stream)")


class EngineError(RuntimeError):
    """The CUDA engine reported an error (or cannot run at all)."""


class Config(C.Structure):
    _fields_ = [
        ("n_wd_rings", C.c_int), ("n_disc_r", C.c_int), ("n_disc_th", C.c_int), ("n_bs", C.c_int),
        ("n_donor_th", C.c_int), ("n_quad", C.c_int), ("donor_ulimb", C.c_double), ("donor_gdexp", C.c_double),
    ]


_lib = None


def library_path():
    return os.environ.get("LFB_LIB") or _build.LIB     # (LFB_LIB: a differently tuned build, for experiments)


def load():
    """dlopen the in-tree library, declaring every prototype of include/lfit_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(
            "CUDA library %s is not built (run `python -m lfit_python_b200._build`); "
            "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    lib.lfb_create.argtypes = [C.c_int, C.POINTER(Config), C.POINTER(vp)]
    lib.lfb_destroy.argtypes = [vp]
    lib.lfb_destroy.restype = None
    lib.lfb_last_error.argtypes = [vp]
    lib.lfb_last_error.restype = C.c_char_p
    lib.lfb_get_config.argtypes = [vp, C.POINTER(Config)]
    lib.lfb_set_layout.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, C.c_int, dp]
    lib.lfb_set_priors.argtypes = [vp, C.c_int, ip, ip, dp, dp, dp, ip]
    lib.lfb_set_lightcurves.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong), dp, dp, dp, dp]
    lib.lfb_log_prob.argtypes = [vp, C.c_int, C.c_longlong, vp, vp, vp, vp]
    lib.lfb_calc_flux.argtypes = [vp, C.c_longlong, vp, C.c_int, C.c_int, C.c_int, dp, dp, vp, vp, vp]
    lib.lfb_roche.argtypes = [vp, C.c_int, C.c_longlong, dp, dp, dp, ip]
    lib.lfb_launch_count.argtypes = [vp]
    lib.lfb_launch_count.restype = C.c_longlong
    lib.lfb_robust_calls.argtypes = [vp]
    lib.lfb_robust_calls.restype = C.c_longlong
    lib.lfb_last_kernel_ms.argtypes = [vp]
    lib.lfb_last_kernel_ms.restype = C.c_float
    lib.lfb_measure_fp64_peak.argtypes = [vp, C.c_int, dp]
    lib.lfb_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.lfb_set_trace.argtypes = [vp, C.c_int]
    lib.lfb_last_trace_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.lfb_set_gp.argtypes = [vp, C.c_int, ip, dp]
    lib.lfb_gp_loglike.argtypes = [vp, C.c_longlong, C.c_int, dp, dp, dp, dp, C.c_int, dp, dp]
    lib.lfb_wdphases.argtypes = [vp, C.c_longlong, dp, dp, dp, C.c_int, dp, ip]
    lib.lfb_ingress_egress.argtypes = [vp, C.c_longlong, dp, dp, dp, dp, ip]
    ll, lp = C.c_longlong, C.POINTER(C.c_longlong)
    lib.lfb_sampler_create.argtypes = [vp, ll, C.c_double, C.c_ulonglong, C.c_int, C.POINTER(vp)]
    lib.lfb_sampler_destroy.argtypes = [vp]
    lib.lfb_sampler_destroy.restype = None
    lib.lfb_sampler_set_state.argtypes = [vp, vp, vp, vp]
    lib.lfb_sampler_run.argtypes = [vp, ll, vp]
    lib.lfb_sampler_half_begin.argtypes = [vp, C.c_int, ll, ll, vp, vp]
    lib.lfb_sampler_half_end.argtypes = [vp, C.c_int, vp, C.c_int, ll, vp]
    for name in ("lfb_sampler_packed", "lfb_sampler_positions", "lfb_sampler_log_prob"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = vp
    lib.lfb_sampler_get_state.argtypes = [vp, vp, vp, vp, lp]
    lib.lfb_sampler_set_chain.argtypes = [vp, ll]
    lib.lfb_sampler_read_chain.argtypes = [vp, vp, lp]
    lib.lfb_stretch_draws.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_int, ll, C.c_double, ll, dp]
    lib.lfb_chain_format.argtypes = [ll, ll, C.c_int, dp, C.c_char_p, ll]
    lib.lfb_chain_format.restype = ll
    lib.lfb_chain_append.argtypes = [C.c_char_p, ll, ll, C.c_int, dp]
    lib.lfb_peer_create.argtypes = [vp, C.c_int, C.c_int, ll, C.c_char_p]
    lib.lfb_peer_connect.argtypes = [vp, C.c_char_p]
    lib.lfb_peer_allgather.argtypes = [vp, vp, C.c_int, vp, C.c_int, ll, C.POINTER(vp), vp]
    lib.lfb_peer_status.argtypes = [vp, ip]
    lib.lfb_peer_destroy.argtypes = [vp]
    lib.lfb_peer_destroy.restype = None
    _lib = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Engine:
    """One lfb_handle: an engine bound to one CUDA device."""

    def __init__(self, device=0, **grid):
        self._lib = load()
        cfg = Config()
        for k, v in grid.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown grid option %r" % k)
            setattr(cfg, k, v)
        h = C.c_void_p()
        rc = self._lib.lfb_create(int(device), C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EngineError("lfb_create failed (%d): %s" % (rc, self._lib.lfb_last_error(None).decode()))
        self._h = h
        self.device = int(device)
        self.ndim = self.n_ecl = self.npars = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (what, rc, self._lib.lfb_last_error(self._h).decode()))

    @property
    def config(self):
        cfg = Config()
        self._check(self._lib.lfb_get_config(self._h, C.byref(cfg)), "lfb_get_config")
        return {k: getattr(cfg, k) for k, _ in Config._fields_}

    @property
    def launch_count(self):
        return int(self._lib.lfb_launch_count(self._h))

    @property
    def robust_calls(self):
        """Element solves that fell through to the last-resort solver (device-wide, since load)."""
        return int(self._lib.lfb_robust_calls(self._h))

    def last_kernel_ms(self):
        return float(self._lib.lfb_last_kernel_ms(self._h))

    def last_stage_ms(self):
        """Device ms of {walker, stream, elements, flux, finish, total} of the last log_prob."""
        out = (C.c_float * 6)()
        self._check(self._lib.lfb_last_stage_ms(self._h, out), "lfb_last_stage_ms")
        return dict(zip(("walker", "stream", "elements", "flux", "finish", "total"), [float(x) for x in out]))

    def set_trace(self, on=True):
        """Bracket every kernel of the last batch with CUDA events (see last_trace_ms)."""
        self._check(self._lib.lfb_set_trace(self._h, int(bool(on))), "lfb_set_trace")

    def last_trace_ms(self):
        """Device ms of every kernel of the last log_prob batch on lane 0 (needs set_trace)."""
        out = (C.c_float * len(TRACE_KERNELS))()
        self._check(self._lib.lfb_last_trace_ms(self._h, out), "lfb_last_trace_ms")
        return {k: float(x) for k, x in zip(TRACE_KERNELS, out) if x >= 0.0}

    def set_gp(self, gp_src=None, dist_cp=None):
        """Switch lfb_log_prob's likelihood to the Gaussian process of CVModel.py:603-696
        (gp_src: sources of ln_ampin_gp, ln_ampout_gp, ln_tau_gp; dist_cp per eclipse), or back
        to chi-squared with no arguments."""
        if gp_src is None:
            self._check(self._lib.lfb_set_gp(self._h, 0, None, None), "lfb_set_gp")
            return
        src = np.ascontiguousarray(gp_src, dtype=np.int32)
        dist = _f64(dist_cp).ravel()
        if src.shape != (3,) or dist.shape[0] != self.n_ecl:
            raise ValueError("set_gp: need three sources and one dist_cp per eclipse")
        self._check(self._lib.lfb_set_gp(self._h, 1, src.ctypes.data_as(C.POINTER(C.c_int)), _dp(dist)), "lfb_set_gp")

    def gp_loglike(self, x, ye, resid, hyper, gaps):
        """george-style GP log-likelihood of residual rows: x, ye of length n (x ascending),
        resid (n_sets, n), hyper (n_sets, 3) = (ampin, ampout, tau), gaps (n_sets, n_gaps, 2)."""
        x, ye = _f64(x).ravel(), _f64(ye).ravel()
        resid = np.atleast_2d(_f64(resid))
        hyper = np.atleast_2d(_f64(hyper))
        n_sets, n = resid.shape
        gaps = _f64(gaps).reshape(n_sets, -1, 2)
        if x.shape[0] != n or ye.shape[0] != n or hyper.shape != (n_sets, 3):
            raise ValueError("gp_loglike: inconsistent shapes")
        out = np.empty(n_sets)
        self._check(self._lib.lfb_gp_loglike(self._h, n_sets, n, _dp(x), _dp(ye), _dp(resid), _dp(hyper),
                                             gaps.shape[1], _dp(gaps), _dp(out)), "lfb_gp_loglike")
        return out

    def wdphases(self, q, incl_deg, r1, ntheta=10):
        """(phi3, phi4) rows and an ok mask (trm.roche.wdphases)."""
        q, incl_deg, r1 = np.broadcast_arrays(_f64(q), _f64(incl_deg), _f64(r1))
        q, incl_deg, r1 = (np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (q, incl_deg, r1))
        n = q.shape[0]
        out = np.empty((n, 2))
        ok = np.empty(n, dtype=np.int32)
        self._check(self._lib.lfb_wdphases(self._h, n, _dp(q), _dp(incl_deg), _dp(r1), int(ntheta), _dp(out),
                                           ok.ctypes.data_as(C.POINTER(C.c_int))), "lfb_wdphases")
        return out, ok.astype(bool)

    def ingress_egress(self, q, incl_deg, pts):
        """Ingress / egress phases of elements pts (n, 5) = (x, y, z, xi, eta); (n, 2) and an ok mask."""
        pts = np.atleast_2d(_f64(pts))
        n = pts.shape[0]
        q = np.ascontiguousarray(np.broadcast_to(_f64(q), (n,)))
        incl_deg = np.ascontiguousarray(np.broadcast_to(_f64(incl_deg), (n,)))
        if pts.shape[1] != 5:
            raise ValueError("ingress_egress: pts must be (n, 5)")
        out = np.empty((n, 2))
        ok = np.empty(n, dtype=np.int32)
        self._check(self._lib.lfb_ingress_egress(self._h, n, _dp(q), _dp(incl_deg), _dp(pts), _dp(out),
                                                 ok.ctypes.data_as(C.POINTER(C.c_int))), "lfb_ingress_egress")
        return out, ok.astype(bool)

    def measure_fp64_peak(self, iters=20000):
        """Sustained DFMA rate of this device in TFLOP/s (roofline denominator)."""
        out = C.c_double()
        self._check(self._lib.lfb_measure_fp64_peak(self._h, int(iters), C.byref(out)), "lfb_measure_fp64_peak")
        return out.value

    # -- flattened tree -----------------------------------------------------------------
    def set_layout(self, ndim, npars, gather, consts):
        gather = _i32(gather).reshape(-1, NPAR)
        consts = _f64(consts).ravel()
        self._check(self._lib.lfb_set_layout(self._h, int(ndim), gather.shape[0], int(npars), _ip(gather),
                                             consts.shape[0], _dp(consts) if consts.size else None),
                    "lfb_set_layout")
        self.ndim, self.n_ecl, self.npars = int(ndim), gather.shape[0], int(npars)

    def set_priors(self, src, ptype, p1, p2, norm, isvar):
        src, ptype, isvar = _i32(src), _i32(ptype), _i32(isvar)
        p1, p2, norm = _f64(p1), _f64(p2), _f64(norm)
        self._check(self._lib.lfb_set_priors(self._h, src.shape[0], _ip(src), _ip(ptype), _dp(p1), _dp(p2),
                                             _dp(norm), _ip(isvar)), "lfb_set_priors")

    def set_lightcurves(self, off, phase, width, y, ye):
        off = np.ascontiguousarray(off, dtype=np.int64)
        phase, width, y, ye = _f64(phase), _f64(width), _f64(y), _f64(ye)
        self._check(self._lib.lfb_set_lightcurves(self._h, off.shape[0] - 1,
                                                  off.ctypes.data_as(C.POINTER(C.c_longlong)),
                                                  _dp(phase), _dp(width), _dp(y), _dp(ye)), "lfb_set_lightcurves")

    # -- hot path -----------------------------------------------------------------------
    def log_prob(self, theta, what=LN_PROB, return_chisq=False, out=None):
        """theta: (n, ndim) host array -> ln-values (n,) [and chisq (n, n_ecl)].  Page-locked arrays (e.g. the
        .numpy() view of a pinned torch tensor) for theta and `out` are copied from / to as they stand; pageable
        ones are staged through the engine's own pinned buffers."""
        theta = _f64(theta)
        if theta.ndim != 2 or theta.shape[1] != self.ndim:
            raise ValueError("Wrong vector length - Expected {}, got {}".format(self.ndim, theta.shape[-1]))
        n = theta.shape[0]
        if out is None:
            out = np.empty(n)
        elif out.dtype != np.float64 or out.shape != (n,) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of shape (n,)")
        chis = np.empty((n, self.n_ecl)) if return_chisq else None
        self._check(self._lib.lfb_log_prob(self._h, int(what), n, theta.ctypes.data, out.ctypes.data,
                                           chis.ctypes.data if return_chisq else None, None), "lfb_log_prob")
        return (out, chis) if return_chisq else out

    def log_prob_device(self, theta_ptr, n, out_ptr, what=LN_PROB, chisq_ptr=None, stream=None):
        """Raw-pointer variant (device or host pointers, e.g. torch tensors' data_ptr())."""
        self._check(self._lib.lfb_log_prob(self._h, int(what), int(n), theta_ptr, out_ptr, chisq_ptr, stream),
                    "lfb_log_prob")

    def calc_flux(self, pars, phase, width=None, flags=0, components=False):
        pars = _f64(pars)
        single = pars.ndim == 1
        pars = np.atleast_2d(pars)
        if pars.shape[1] not in (14, 18):
            raise ValueError("CV takes 14 (simple BS) or 18 (complex BS) parameters, got %d" % pars.shape[1])
        phase = _f64(phase).ravel()
        n_ph = phase.shape[0]
        if width is None:
            width = np.zeros(n_ph)
        width = _f64(np.broadcast_to(np.asarray(width, dtype=np.float64), (n_ph,)))
        n = pars.shape[0]
        tot = np.empty((n, n_ph))
        comp = np.empty((4, n, n_ph)) if components else None
        self._check(self._lib.lfb_calc_flux(self._h, n, pars.ctypes.data, pars.shape[1], int(flags), n_ph, _dp(phase),
                                            _dp(width), tot.ctypes.data, comp.ctypes.data if components else None,
                                            None), "lfb_calc_flux")
        if single:
            tot = tot[0]
            comp = comp[:, 0] if components else None
        return (tot, comp) if components else tot

    def roche(self, which, a, b=None):
        a = _f64(np.atleast_1d(a))
        b = _f64(np.broadcast_to(np.atleast_1d(0.0 if b is None else b), a.shape))
        out = np.empty((a.shape[0], 4))
        ok = np.empty(a.shape[0], dtype=np.int32)
        self._check(self._lib.lfb_roche(self._h, int(which), a.shape[0], _dp(a), _dp(b), _dp(out), _ip(ok)),
                    "lfb_roche")
        return out, ok.astype(bool)


def chain_text(rows):
    """rows (n_steps, n, ndim + 1) -> the reference's chain lines (mcmc_utils.py:163-164) as bytes."""
    rows = _f64(rows)
    if rows.ndim != 3:
        raise ValueError("chain_text: rows must be (n_steps, n_walkers, ndim + 1)")
    lib = load()
    steps, n, w = rows.shape
    size = lib.lfb_chain_format(steps, n, w - 1, _dp(rows), None, 0)
    if size < 0:
        raise EngineError("lfb_chain_format failed")
    buf = C.create_string_buffer(int(size) + 1)
    lib.lfb_chain_format(steps, n, w - 1, _dp(rows), buf, size)
    return buf.raw[:size]


def chain_append(path, rows):
    """Append rows (n_steps, n, ndim + 1) to the chain file with one write."""
    rows = _f64(rows)
    if rows.ndim != 3:
        raise ValueError("chain_append: rows must be (n_steps, n_walkers, ndim + 1)")
    steps, n, w = rows.shape
    rc = load().lfb_chain_append(os.fsencode(path), steps, n, w - 1, _dp(rows))
    if rc != 0:
        raise EngineError("lfb_chain_append(%r) failed (%d)" % (path, rc))


def stretch_draws(seed, step, half, half_n, a, cnt):
    """(z, partner row, ln u') of rows [0, cnt) of one half at one step -- the sampler kernels' random stream."""
    out = np.empty((int(cnt), 3))
    rc = load().lfb_stretch_draws(int(seed), int(step), int(half), int(half_n), float(a), int(cnt), _dp(out))
    if rc != 0:
        raise EngineError("lfb_stretch_draws failed (%d)" % rc)
    return out


_default_engines = {}


def default_engine(device=0):
    """Process-wide engine with the default surface grid (created on first use)."""
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]
