"""CPU tests of the NVRTC specialisation: the colour-pass kernel source (struct_pass.cuh) plus the
generated lattice prologue compiles for sm_100a for every model family / precision - no GPU needed up
to the cubin.  (On the GPU the parity suite runs once more under MCG_JIT=1.)"""
import os

import pytest

from mcsolver_b200 import engine
from mcsolver_b200.lattice import add_dipole_stencil
from tests.specs import spec_of

CASES = [("cubic", (8, 8, 16), 3, 32, 2), ("cubic", (8, 8, 16), 3, 64, 2), ("cubic", (8, 8, 16), 1, 32, 2), ("square", (16, 16, 1), 2, 32, 2),
         ("skyrmion", (16, 16, 1), 3, 32, 2), ("cri3", (16, 16, 1), 3, 32, 8), ("aniso", (16, 16, 2), 2, 64, None)]


@pytest.mark.parametrize("name,L,model,prec,ncol", CASES, ids=lambda v: str(v))
def test_specialised_kernels_compile_for_sm100a(name, L, model, prec, ncol, tmp_path, monkeypatch):
    monkeypatch.setenv("MCG_CACHE_DIR", str(tmp_path))          # force a real compilation, keep the tree clean
    try:
        n, report = engine.jit_check(spec_of(name, L), model, prec)
    except engine.McgError as e:
        if "cannot dlopen libnvrtc" in str(e):
            pytest.skip("NVRTC not installed")
        raise
    assert "V=%d" % (4 if prec == 32 else 2) in report, report
    if ncol is not None:
        assert n == ncol, report
    topo = 1 if (name == "skyrmion" and model == 3) else 0       # + the specialised topological-charge kernel
    assert ("topological charge: module" in report) == bool(topo), report
    assert n >= 2 and len(os.listdir(tmp_path)) == n + topo     # one cached cubin per colour


def test_wide_stencils_are_left_to_the_runtime_table_kernel(tmp_path, monkeypatch):
    monkeypatch.setenv("MCG_CACHE_DIR", str(tmp_path))
    spec = add_dipole_stencil(spec_of("cubic", (16, 16, 16)), 0.2, 2.0)
    n, report = engine.jit_check(spec, 3, 32)
    assert n == 0 and "not eligible" in report                  # 32 links per site: unrolling would thrash the I-cache


def test_block_spin_descriptor_is_validated_on_the_host(tmp_path, monkeypatch):
    """The structured block-spin tables are built with the class tables (no GPU needed): supercells the descriptor cannot
    express exactly (odd sizes; sizes where the doubled bonds +2d and -2d fold onto one site) are refused by name."""
    monkeypatch.setenv("MCG_CACHE_DIR", str(tmp_path))
    try:
        n, _ = engine.jit_check(spec_of("cubic", (8, 8, 16), ), 3, 32, block_spin=True)
    except engine.McgError as e:
        if "cannot dlopen libnvrtc" in str(e):
            pytest.skip("NVRTC not installed")
        raise
    assert n == 2
    for L in [(5, 8, 16), (4, 8, 16)]:
        with pytest.raises(engine.McgError, match="block_spin"):
            engine.jit_check(spec_of("cubic", L), 3, 32, block_spin=True)
