"""B200-native late-fusion head + image-text auxiliary losses (drop-in for the reference's models/mm_late.py path).

The directory name follows the repository naming contract and is not a valid Python identifier; import it through the
alias package `tic_b200` at the repository root (`import tic_b200`), which points its __path__ here.
"""
from . import capi  # noqa: F401
from .capi import TicError, load  # noqa: F401

__all__ = ["capi", "TicError", "load"]
