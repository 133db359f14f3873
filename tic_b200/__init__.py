"""Import alias for the package directory `socialmedia-textimage-classification-auxlosses_b200/` (whose name is not a
valid Python identifier): `import tic_b200` resolves every submodule (capi, plan, mm_late, utils, ...) from there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "socialmedia-textimage-classification-auxlosses_b200")
if not _os.path.isdir(_PKG_DIR):
    raise ImportError("package directory not found: %s" % _PKG_DIR)
__path__.insert(0, _PKG_DIR)  # submodules are looked up in the real package directory first

from . import capi  # noqa: E402,F401
from .capi import TicError, load  # noqa: E402,F401
