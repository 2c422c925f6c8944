"""Drop-in for the reference's `isinglib` extension module (isingLib.c:259, :453-471).

`mcMain.py:110-113` does `from isinglib import MCMainFunction` (fallback `mcsolver.lib.isinglib`).
Same 16 positional arguments, same 11-item result tuple; slot 4/5 (<e>, <e^2>) hold the ABSOLUTE
energy per site for both algorithms (the reference's Metropolis energy is relative to an arbitrary
zero, isingLib.c:348,359 - SURVEY 8 quirks).  No CPU fallback.
Environment: MCSOLVER_B200_SEED, MCSOLVER_B200_PRECISION (64 | 32).
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from mcsolver_b200.engine import run_ising_args as _run  # noqa: E402


def MCMainFunction(*args):
    """execute Monte Carlo sims. on Ising model (isingLib.c:454)"""
    return _run(args)
