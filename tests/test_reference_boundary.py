"""The drop-in boundary seen from the reference's side: the UNMODIFIED /root/reference/model.py and CVModel.py,
imported through lfit_python_b200.compat.install(), build the shipped example tree and evaluate it with this
package's `lfit` / `trm.roche` replacements (/root/reference/CVModel.py:13,15,128,138,222,288,460).

Only what is irrelevant to the path is stubbed (george, matplotlib, networkx: imported by the reference, unused for a
useGP = 0 tree).  The reference is not on the GPU box, so there are two variants:
  * CPU (runs here): the engine behind the shims is a stand-in answering the same calls with the CPU oracle -- what is
    proven is that the reference's own tree code runs on the shims' API and agrees with this package's mirror tree;
  * `-m gpu` (needs the reference AND a GPU; skipped where either is missing): the same with the CUDA engine, plus the
    batched VectorModel against the reference tree walked walker by walker.
The mirror tree against the CUDA engine and the oracle is what the other GPU tests establish.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "CVModel.py")),
                                reason="the reference tree is not on this machine")


class OracleEngine:
    """Stand-in for _cabi.Engine in a container without a GPU: the calls lfit.py / roche.py make, answered by the
    CPU oracle.  Test infrastructure only."""

    def calc_flux(self, pars, phase, width=None, flags=0, components=False):
        st, tot, comp = O.calc_flux(pars, phase, width, flags=flags, components=True)
        comp = np.asarray(comp)
        return (tot, comp) if components else tot

    def roche(self, which, a, b=None):
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        b = np.broadcast_to(np.atleast_1d(0.0 if b is None else b), a.shape)
        out, ok = np.full((a.shape[0], 4), np.nan), np.zeros(a.shape[0], dtype=bool)
        fn = {0: lambda q, _: (O.xl1(q),), 1: lambda q, i: (O.findphi(q, i),), 2: lambda q, d: (O.findi(q, d),),
              3: lambda q, r: O.bspot(q, r)}[which]
        for k in range(a.shape[0]):
            try:
                v = fn(float(a[k]), float(b[k]))
                out[k, : len(v)] = v
                ok[k] = True
            except O.RocheError:
                pass
        return out, ok


def _load_reference(monkeypatch):
    """import model, CVModel from /root/reference with the shims in place."""
    from lfit_python_b200 import compat
    for name in ("george", "networkx", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for name in ("lfit", "trm", "trm.roche", "configobj", "model", "CVModel"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    compat.install()
    mods = {}
    for name in ("model", "CVModel"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        monkeypatch.setitem(sys.modules, name, mod)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["model"], mods["CVModel"]


def _input_without_gp(tmp_path):
    """The shipped example (test_data/mcmc_input.dat) with useGP = 0 and absolute light-curve paths."""
    txt = open(os.path.join(REF, "test_data", "mcmc_input.dat")).read()
    assert "useGP = 1" in txt
    txt = txt.replace("useGP = 1", "useGP = 0").replace("lightcurves/test_data_", os.path.join(REF, "test_data", "lightcurves", "test_data_"))
    path = tmp_path / "mcmc_input.dat"
    path.write_text(txt)
    return str(path)


def _compare_trees(ref_model, mirror, rng):
    assert list(ref_model.dynasty_par_names) == list(mirror.dynasty_par_names)
    assert np.array_equal(ref_model.dynasty_par_vals, mirror.dynasty_par_vals)
    assert len(ref_model.dynasty_par_vals) == 84                     # 6 complex eclipses in 3 bands
    p0 = np.array(mirror.dynasty_par_vals, dtype=np.float64)
    vectors = [p0, p0 * (1.0 + 0.01 * rng.standard_normal(p0.shape))]
    bad = p0.copy()
    bad[list(mirror.dynasty_par_names).index("dphi_core")] = 0.09     # wider than the edge-on eclipse: CVModel.py:452-473
    vectors.append(bad)
    for v in vectors:
        ref_model.dynasty_par_vals = v
        mirror.dynasty_par_vals = v
        a, b = ref_model.ln_prior(), mirror.ln_prior()
        assert (a == b) or np.isclose(a, b, rtol=1e-12)
        if np.isfinite(b):
            assert np.isclose(ref_model.chisq(), mirror.chisq(), rtol=1e-12)
            assert np.isclose(ref_model.ln_like(), mirror.ln_like(), rtol=1e-12)
        a, b = ref_model.ln_prob(), mirror.ln_prob()
        assert (a == b) or np.isclose(a, b, rtol=1e-12)
    assert ref_model.ln_prob() == -np.inf                            # the last vector violates LCModel.ln_prior
    # leaves: parameter lists in CV order (yaw / tilt swap, CVModel.py:379,387) and the component curves
    ref_model.dynasty_par_vals = p0
    mirror.dynasty_par_vals = p0
    rl, ml = ref_model.search_node_type("Eclipse"), mirror.search_node_type("Eclipse")
    rl, ml = sorted(rl, key=lambda e: e.name), sorted(ml, key=lambda e: e.name)
    assert [e.name for e in rl] == [e.name for e in ml] and len(rl) == 6
    for r, m in zip(rl, ml):
        assert list(r.cv_parnames) == list(m.cv_parnames)
        assert np.array_equal(np.asarray(r.cv_parlist, dtype=float), np.asarray(m.cv_parlist, dtype=float))
        assert r.lc.n_data == m.lc.n_data and np.array_equal(r.lc.x, m.lc.x) and np.array_equal(r.lc.w, m.lc.w)
        for ca, cb in zip(r.calcComponents(), m.calcComponents()):
            assert np.allclose(ca, cb, rtol=1e-12, atol=0)
        assert np.isclose(r.chisq(), m.chisq(), rtol=1e-12)
    return p0


def test_unmodified_reference_tree_runs_on_the_shims(tmp_path, monkeypatch):
    from lfit_python_b200 import _cabi
    from lfit_python_b200 import CVModel as mirror_cv
    monkeypatch.setattr(_cabi, "default_engine", lambda device=0: OracleEngine())
    model_mod, cv_mod = _load_reference(monkeypatch)
    assert cv_mod.lfit.__name__ == "lfit_python_b200.lfit" and cv_mod.roche.__name__ == "lfit_python_b200.roche"
    path = _input_without_gp(tmp_path)
    ref_model = cv_mod.construct_model(path)
    mirror = mirror_cv.construct_model(path)
    assert type(ref_model).__name__ == "LCModel" and type(ref_model).__module__ == "CVModel"
    _compare_trees(ref_model, mirror, np.random.default_rng(0))
    # the weak contact with real lfit: parameters fitted with lfit give a sane chi-squared
    dof = sum(e.lc.n_data for e in ref_model.search_node_type("Eclipse")) - 84 - 1
    assert 1.0 < ref_model.chisq() / dof < 10.0


@pytest.mark.gpu
def test_unmodified_reference_tree_on_the_cuda_engine(tmp_path, monkeypatch):
    from lfit_python_b200 import CVModel as mirror_cv
    model_mod, cv_mod = _load_reference(monkeypatch)
    path = _input_without_gp(tmp_path)
    ref_model = cv_mod.construct_model(path)
    mirror = mirror_cv.construct_model(path)
    rng = np.random.default_rng(0)
    p0 = _compare_trees(ref_model, mirror, rng)
    # the batched path against the reference tree walked walker by walker (mcmcfit.py:37-41)
    vec = mirror.vectorised()
    theta = p0 * (1.0 + 0.02 * rng.standard_normal((12, p0.shape[0])))
    got = vec.ln_prob(theta)
    for k in range(theta.shape[0]):
        ref_model.dynasty_par_vals = theta[k]
        want = ref_model.ln_prob()
        assert (got[k] == want) or np.isclose(got[k], want, rtol=1e-9)
