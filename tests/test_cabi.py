"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol the header declares,
and refuses to run without a GPU (no compute calls are made here)."""
import os
import re

import pytest

from lfit_python_b200 import _build, _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_the_header():
    lib_path = _build.build()
    assert os.path.exists(lib_path)
    header = open(os.path.join(ROOT, "include", "lfit_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lfb_[a-z0-9_]+)\s*\(", header)))
    assert "lfb_log_prob" in declared and "lfb_calc_flux" in declared and len(declared) >= 12
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert sorted(_cabi.EXPORTS) == declared


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.EngineError, match="no CUDA device"):
        _cabi.Engine(0)
    from lfit_python_b200 import lfit
    cv = lfit.CV([0.05, 0.07, 0.06, 0.013, 0.1, 0.04, 0.45, 0.3, 0.02, 0.04, 120.0, 0.05, 0.5, 0.0])
    with pytest.raises(_cabi.EngineError):
        cv.calcFlux(cv.pars, [0.0, 0.1], None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lfit_python_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f in ("workloads.py",), f
