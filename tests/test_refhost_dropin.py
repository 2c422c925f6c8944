"""Build-container-only check (needs /root/reference): the UNMODIFIED reference host code
(win.startSimulation -> mcMain.MC -> `from xylib import MCMainFunction`) reaches our shim modules
when mcsolver_b200/lib is on sys.path, hands them the positional tuples they expect, and - there
being no GPU in the container - gets the loud MCG_ERR_CUDA instead of any CPU fallback.
On the GPU box /root/reference is absent and the test is skipped."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.refhost
def test_reference_host_code_imports_our_shims(tmp_path, monkeypatch):
    from oracle import refharness as rh
    if not rh.have_reference_host():
        pytest.skip("/root/reference not present")
    from mcsolver_b200 import _ffi
    import ctypes
    n = ctypes.c_int(0)
    if _ffi.lib().mcg_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present: covered by the gpu tests")
    libdir = os.path.join(ROOT, "mcsolver_b200", "lib")
    monkeypatch.syspath_prepend(libdir)
    for name in ("isinglib", "xylib", "heisenberglib"):
        sys.modules.pop(name, None)
    Lattice, mcMain, win, fileio = rh.load_reference_host()
    import numpy as np
    from mcsolver_b200.engine import McgError
    monkeypatch.chdir(tmp_path)
    bond = Lattice.Bond(0, 0, np.array([1, 0, 0]), -1.0, -1.0, -1.0, 0, 0, 0, 0, 0, 0, True)
    mc = mcMain.MC(0, np.eye(3), pos=np.zeros((1, 3)), S=[1.0], D=[[0, 0, 0]], bondList=[bond], T=1.0, Lx=6, Ly=6, Lz=1,
                   On=2, orbGroupList=[[0]], groupInSC=True)
    with pytest.raises(McgError) as e:
        mc.mainLoopViaCLib_On(nsweep=10, nthermal=10, ninterval=0, algo="Metropolis", On=2)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    assert "mcsolver_b200" in sys.modules["xylib"].__file__
