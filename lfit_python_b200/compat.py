"""Make the unmodified reference scripts import this package's replacements.

    import lfit_python_b200.compat as compat; compat.install()
    import lfit                  # -> lfit_python_b200.lfit
    from trm import roche        # -> lfit_python_b200.roche
    import configobj             # -> lfit_python_b200.configobj (only if the real one is missing)
"""
import sys
import types


def install(force_configobj=False):
    from . import configobj as _configobj
    from . import lfit as _lfit
    from . import roche as _roche
    sys.modules["lfit"] = _lfit
    trm = sys.modules.get("trm") or types.ModuleType("trm")
    trm.roche = _roche
    sys.modules["trm"] = trm
    sys.modules["trm.roche"] = _roche
    if force_configobj or "configobj" not in sys.modules:
        try:
            if force_configobj:
                raise ImportError
            import configobj  # noqa: F401
        except ImportError:
            sys.modules["configobj"] = _configobj
