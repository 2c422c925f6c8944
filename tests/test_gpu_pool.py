"""scan.run_points keeps created systems and recycles them (mcg_recycle): a recycled system must reproduce a fresh one bit for bit."""
import numpy as np
import pytest

from tests.specs import spec_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [("cubic", (16, 16, 16), 3, 0, 32, False), ("square", (32, 32, 1), 1, 0, 8, False), ("square", (16, 16, 1), 2, 1, 64, True),
                                  ("skyrmion", (12, 12, 1), 3, 0, 32, False), ("square", (320, 320, 1), 1, 1, 32, False)],
                         ids=["heisenberg-fp32", "ising-int8", "xy-wolff-tables-fp64", "skyrmion-fp32", "ising-wolff-hybrid-fp32"])
def test_recycled_system_reproduces_a_fresh_one(case, monkeypatch):
    from mcsolver_b200 import scan
    name, L, model, algo, prec, tables = case
    spec = spec_of(name, L)
    Tc = {1: 2.3, 2: 0.9, 3: 1.4}[model] * (0.3 if name == "skyrmion" else 1.0)
    jobs = [(np.linspace(0.8, 1.2, 4) * Tc, np.zeros(4), 3), (np.linspace(0.9, 1.5, 4) * Tc, np.full(4, 0.1 if name == "skyrmion" else 0.0), 11),
            (np.linspace(0.8, 1.2, 4) * Tc, np.zeros(4), 3)]
    nint = 5 if algo == 1 else 0

    def run(T, H, seed):
        return scan.run_points(spec, model, T, H, 20, 30, ninterval=nint, algorithm=algo, precision=prec, seed=seed, tables=tables)[1]

    monkeypatch.setenv("MCG_POOL", "0")
    scan.clear_pool()
    fresh = [run(*j) for j in jobs]
    assert len(scan._pool) == 0
    monkeypatch.setenv("MCG_POOL", "1")
    pooled = [run(*j) for j in jobs]
    assert len(scan._pool) == 1                       # one lattice: one system, created once and recycled twice
    for a, b in zip(fresh, pooled):
        assert np.array_equal(a, b)
    assert np.array_equal(pooled[0], pooled[2]) and not np.array_equal(pooled[0], pooled[1])
    scan.clear_pool()


def test_pool_is_bounded_and_keyed_by_everything_creation_depends_on():
    from mcsolver_b200 import scan
    scan.clear_pool()
    T, H = np.array([1.0, 1.5]), np.zeros(2)
    for L in (8, 10, 12, 8):
        scan.run_points(spec_of("cubic", (L, L, L)), 3, T, H, 2, 4, precision=32)
    assert len(scan._pool) == scan.POOL_MAX
    scan.run_points(spec_of("cubic", (8, 8, 8)), 3, T, H, 2, 4, precision=64)       # other precision: another system
    scan.run_points(spec_of("cubic", (8, 8, 8)), 3, np.array([1.0]), np.zeros(1), 2, 4, precision=64)   # other batch size
    assert len(scan._pool) == scan.POOL_MAX
    scan.clear_pool()
    assert len(scan._pool) == 0
