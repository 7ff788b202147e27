"""Conformance of the host-side tree mirror with the behaviours of the reference's
model.py / CVModel.py / mcmcfit.py listed in SURVEY.md appendix A (no GPU needed)."""
import os

import numpy as np
import pytest

from lfit_python_b200 import mcmcfit
from lfit_python_b200.CVModel import (Band, ComplexEclipse, LCModel, Lightcurve, SimpleEclipse, construct_model)
from lfit_python_b200.configobj import ConfigObj
from lfit_python_b200.flatten import FlatLayout
from lfit_python_b200.model import Node, Param, Prior, extract_par_and_key

INPUT = """
fit = 0   # do not fit
nburn = 10
nprod = 10
nwalkers = 40
nthread = 2
usePT = 0
ntemps = 1
first_scatter = 0.10
double_burnin = 0
second_scatter = 0.05
comp_scat = 1
phi_start = -0.2
phi_end = 0.3
complex = {complex}
useGP = 0
{extra}
q = 0.1037 uniform 0.03 0.5 1
dphi = 0.0392 uniform 0.01 0.1 1
rwd = 0.0187 uniform 0.001 0.1 1
wdFlux_g = 0.0528 uniform 0.001 0.2 1
rsFlux_g = 0.0131 uniform 0.001 0.2 1
ulimb_g = 0.284 gauss 0.284 0.001 {ulimb_var}
wdFlux_r = 0.0324 uniform 0.001 0.2 1
rsFlux_r = 0.0262 uniform 0.001 0.2 1
ulimb_r = 0.284 gauss 0.284 0.001 1
wdFlux_unused = 0.0324 uniform 0.001 0.2 1
rsFlux_unused = 0.0262 uniform 0.001 0.2 1
ulimb_unused = 0.284 gauss 0.284 0.001 1
file_0 = lc0.calib
band_0 = g
dFlux_0 = 0.0707 uniform 0.001 0.2 1
sFlux_0 = 0.0613 uniform 0.001 0.2 1
rdisc_0 = 0.2953 uniform 0.2 0.7 1
scale_0 = 0.043 log_uniform 0.001 0.2 1
az_0 = 120.0 uniform 50.0 175.0 1
fis_0 = 0.048 uniform 0.001 1.0 1
dexp_0 = 0.5 log_uniform 0.001 2.0 1
phi0_0 = 0.001 uniform -0.2 0.2 1
exp1_0 = 1.1342 uniform 0.001 5.0 1
exp2_0 = 4.5971 uniform 0.5 5.0 1
yaw_0 = 5.4 uniform -90.0 90.0 1
tilt_0 = 72.0006 uniform 0.001 180.0 1
file_b = lc1.calib
band_b = r
dFlux_b = 0.1238 uniform 0.001 0.2 1
sFlux_b = 0.1518 uniform 0.001 0.2 1
rdisc_b = 0.5214 uniform 0.2 0.7 1
scale_b = 0.0497 log_uniform 0.001 0.2 1
az_b = 122.0724 uniform 50.0 175.0 1
fis_b = 0.1684 uniform 0.001 1.0 1
dexp_b = 1.9539 log_uniform 0.001 2.0 1
phi0_b = -0.0013 uniform -0.2 0.2 1
exp1_b = 3.4876 uniform 0.001 5.0 1
exp2_b = 1.4429 uniform 0.5 5.0 1
yaw_b = 15.6635 uniform -90.0 90.0 1
tilt_b = 52.472 uniform 0.001 180.0 1
"""


def write_input(tmp_path, complex=1, ulimb_var=1, extra=""):
    rng = np.random.default_rng(1)
    for k, (n, sep) in enumerate([(260, " "), (180, ",")]):
        x = np.linspace(-0.25, 0.35, n)
        y = 0.2 + 0.01 * rng.standard_normal(n)
        y[5] = np.nan
        rows = ["# phase flux err"] + [sep.join("%.10f" % v for v in (a, b, 0.004)) for a, b in zip(x, y)]
        (tmp_path / ("lc%d.calib" % k)).write_text("\n".join(rows) + "\n")
    path = tmp_path / "mcmc_input.dat"
    path.write_text(INPUT.format(complex=complex, ulimb_var=ulimb_var, extra=extra))
    return str(path)


def test_extract_par_and_key():
    assert extract_par_and_key("wdFlux_long_label") == ("wdFlux", "long_label")
    assert extract_par_and_key("ln_ampin_gp_core") == ("ln_ampin_gp", "core")
    assert extract_par_and_key("q") == ("q", "")


def test_param_from_string_and_priors():
    p = Param.fromString("q", "0.1037 uniform 0.03 0.5 1")
    assert (p.startVal, p.currVal, p.isVar, p.prior.type, p.prior.p1, p.prior.p2) == (0.1037, 0.1037, True, "uniform", 0.03, 0.5)
    assert Param.fromString("q", "0.1 uniform 0 1").isVar is True
    assert Param.fromString("q", "0.1 uniform 0 1 0").isVar is False
    assert p.isValid
    p.currVal = 0.6
    assert not p.isValid
    assert abs(Prior("log_uniform", 0.001, 0.2).normalise - 0.513979827) < 1e-8
    assert abs(Prior("mod_jeff", 0.01, 1.0).normalise - np.log(101.0)) < 1e-14
    assert Prior("gauss", 0.0, 1.0).ln_prob(40.0) == -np.inf
    with pytest.raises(AssertionError):
        Prior("lorentz", 0, 1)


def test_config_shim(tmp_path):
    path = write_input(tmp_path)
    cfg = ConfigObj(path)
    assert cfg["fit"] == "0" and cfg["q"] == "0.1037 uniform 0.03 0.5 1" and cfg["file_0"] == "lc0.calib"
    assert "neclipses" not in cfg


def test_lightcurve_reading_and_trim(tmp_path):
    write_input(tmp_path)
    for name in ("lc0.calib", "lc1.calib"):  # space and comma separated
        lc = Lightcurve.from_calib(str(tmp_path / name))
        n = lc.n_data
        assert not np.isnan(lc.y).any() and lc.name == name
        w0 = lc.w[0]
        assert np.allclose(lc.w, np.mean(np.diff(lc.x)) / 2)
        lc.trim(-0.2, 0.3)
        assert lc.n_data < n and lc.x.min() > -0.2 and lc.x.max() < 0.3
        assert np.all(lc.w == w0)  # width comes from the untrimmed curve (CVModel.py:64 vs :894)


def test_construct_model_structure(tmp_path):
    m = construct_model(write_input(tmp_path))
    assert isinstance(m, LCModel) and m.name == "LCModel_core" and m.is_root
    assert [b.name for b in m.children] == ["Band_g", "Band_r"]  # band without eclipses pruned
    assert [e.name for b in m.children for e in b.children] == ["ComplexEclipse_0", "ComplexEclipse_b"]
    names = m.dynasty_par_names
    assert names[:6] == ["q_core", "dphi_core", "rwd_core", "wdFlux_g", "rsFlux_g", "ulimb_g"]
    assert names[6:18] == ["dFlux_0", "sFlux_0", "rdisc_0", "scale_0", "az_0", "fis_0", "dexp_0", "phi0_0",
                           "exp1_0", "exp2_0", "yaw_0", "tilt_0"]
    assert len(names) == 3 + 2 * 3 + 2 * 12
    assert m.search_Node("Band", "r").label == "r"
    assert {n.name for n in m.search_node_type("Eclipse")} == {"ComplexEclipse_0", "ComplexEclipse_b"}
    assert m["rdisc_b"].currVal == 0.5214
    m["rdisc_b"] = 0.5
    assert m.search_par("b", "rdisc").currVal == 0.5
    ecl = m.search_Node("ComplexEclipse", "0")
    assert ecl.cv_parlist[:6] == [0.0528, 0.0707, 0.0613, 0.0131, 0.1037, 0.0392]
    assert ecl.cv_parlist[-4:] == [1.1342, 4.5971, 72.0006, 5.4]  # CV order is (tilt, yaw), node order (yaw, tilt)
    assert m.structure["id"] == "LCModel_core" and len(m.structure["children"]) == 2
    assert set(m.create_tree().nodes) == {"LCModel_core", "Band_g", "Band_r", "ComplexEclipse_0", "ComplexEclipse_b"}


def test_simple_model_and_neclipses(tmp_path):
    m = construct_model(write_input(tmp_path, complex=0, extra="neclipses = 1"))
    ecls = [e for b in m.children for e in b.children]
    assert len(ecls) == 1 and isinstance(ecls[0], SimpleEclipse) and not isinstance(ecls[0], ComplexEclipse)
    assert len(m.dynasty_par_names) == 14 and len(ecls[0].cv_parlist) == 14


def test_nodata_dummy(tmp_path):
    m = construct_model(write_input(tmp_path), nodata=True)
    lc = m.children[0].children[0].lc
    assert lc.n_data == 1000 and lc.x[0] == -0.5 and np.all(lc.y == 0) and np.all(lc.ye == 1)


def test_gp_input_needs_the_hyper_parameters(tmp_path):
    path = write_input(tmp_path)
    txt = open(path).read().replace("useGP = 0", "useGP = 1")
    open(path, "w").write(txt)
    with pytest.raises(KeyError):        # ln_ampin_gp & co. are missing from the file (tests/test_gp.py has them)
        construct_model(path)


def test_vector_set_get_and_errors(tmp_path):
    m = construct_model(write_input(tmp_path))
    v = np.arange(len(m.dynasty_par_vals), dtype=float)
    m.dynasty_par_vals = v
    assert m.dynasty_par_vals == list(v)
    assert m.q.currVal == 0 and m.children[1].children[0].tilt.currVal == v[-1]
    assert m.dynasty_par_dict["dphi_core"] == 1.0
    with pytest.raises(ValueError):
        m.dynasty_par_vals = v[:-1]
    with pytest.raises(TypeError):
        Band("x", [Param.fromString("wdFlux", "0.1 uniform 0 1")])
    with pytest.raises(TypeError):
        Node(3, [])
    with pytest.raises(NotImplementedError):
        Node("leaf", []).chisq()


def test_fixed_parameter_leaves_the_vector(tmp_path):
    m = construct_model(write_input(tmp_path, ulimb_var=0))
    assert "ulimb_g" not in m.dynasty_par_names and len(m.dynasty_par_names) == 32
    L = FlatLayout(m)
    assert L.ndim == 32 and L.consts.tolist() == [0.284]
    assert L.gather[0][7] == -1            # ulimb of eclipse 0 comes from the constant slot
    assert L.gather[1][7] >= 0             # band r keeps its own variable ulimb
    assert L.prior_isvar.sum() == 32 and len(L.prior_isvar) == 33  # fixed Params still face their prior


def test_flat_layout_matches_the_tree(tmp_path):
    m = construct_model(write_input(tmp_path))
    L = FlatLayout(m)
    assert L.names == m.dynasty_par_names and np.allclose(L.p0, m.dynasty_par_vals)
    th = np.arange(L.ndim)
    assert th[L.gather[0]].tolist() == [3, 6, 7, 4, 0, 1, 8, 5, 2, 9, 10, 11, 12, 13, 14, 15, 17, 16]
    rng = np.random.default_rng(0)
    v = rng.uniform(0.01, 1.0, L.ndim)
    m.dynasty_par_vals = v
    for k, ecl in enumerate(L.eclipses):
        assert np.allclose(v[L.gather[k]], ecl.cv_parlist)
    assert L.lc_off.tolist() == [0, L.eclipses[0].lc.n_data, L.eclipses[0].lc.n_data + L.eclipses[1].lc.n_data]
    assert np.array_equal(L.lc_phase[: L.lc_off[1]], L.eclipses[0].lc.x)
    # same tables as the hand-built synthetic workload of the same shape
    from lfit_python_b200 import workloads
    wl = workloads.Workload("x", 2, 1, 10)
    assert wl.gather[0].tolist() == L.gather[0].tolist()
    assert np.allclose(wl.prior_norm[:6], L.prior_norm[:6])


def test_scatter_vector(tmp_path):
    m = construct_model(write_input(tmp_path))
    s = mcmcfit.scatter_vector(m, 0.1, comp_scat=True)
    names = m.dynasty_par_names
    assert s[names.index("dphi_core")] == pytest.approx(0.02)
    assert s[names.index("ulimb_g")] == pytest.approx(1e-7)
    assert s[names.index("q_core")] == pytest.approx(0.1)
    assert np.all(mcmcfit.scatter_vector(m, 0.1, comp_scat=False) == 0.1)


def test_dof_formula(tmp_path):
    m = construct_model(write_input(tmp_path))
    n = sum(e.lc.n_data for e in m.search_node_type("Eclipse"))
    assert int(n - len(m.dynasty_par_names) - 1) == n - 34  # mcmcfit.py:143-148
