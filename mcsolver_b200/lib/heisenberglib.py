"""Drop-in for the reference's `heisenberglib` extension module (heisenbergLib.c:478, :887-905).

`mcMain.py:233-237` does `from heisenberglib import MCMainFunction` (fallback
`mcsolver.lib.heisenberglib`); placing this directory on sys.path - or copying the three shim files
into the reference's mcsolver/lib/ - routes the O(3) engine to the B200 library.  Same 23 positional
arguments, same 29-item result tuple.  Raises (instead of segfaulting) on malformed input, and
raises if the CUDA library or a GPU is unavailable: there is no CPU fallback.
Environment: MCSOLVER_B200_SEED (Philox seed, default 1), MCSOLVER_B200_PRECISION (64 | 32).
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from mcsolver_b200.engine import run_on_args as _run  # noqa: E402


def MCMainFunction(*args):
    """the only function in our c lib (heisenbergLib.c:888)"""
    return _run(3, args)
