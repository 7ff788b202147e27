from . import _cabi  # noqa: F401
