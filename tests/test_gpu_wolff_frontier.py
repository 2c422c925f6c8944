"""Wolff hybrid (kernels_wolff.cuh: k_wolff_frontier + global passes): growing the seed's cluster breadth-first from the
seed - what the reference's FIFO does (isingLib.c:165-236, xyLib.c:256-380, heisenbergLib.c:310-439) - must select exactly
the cluster the global bond-percolation passes select, because a bond's state is a function of (site pair, step) alone.

MCG_WOLFF_FRONTIER: 0 plain global sequence, 1 adaptive hybrid (default for >= 65536 sites), 2 frontier tried at every step,
3 hybrid bookkeeping with the frontier kernel declining every step.  MCG_WOLFF_FRONTIER_CAP: member-queue length (a cluster
that outgrows it falls back to the global passes of the same step)."""
import numpy as np
import pytest

from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu

CASES = [("square", (8, 8, 1), 0.9, 2, 0.0), ("cubic", (6, 6, 6), 1.4, 3, 0.0), ("aniso", (6, 6, 1), 0.7, 3, 0.3),
         ("aniso", (6, 6, 1), 0.7, 2, 0.3), ("square", (8, 16, 1), 2.3, 1, 0.05), ("cubic", (6, 6, 6), 4.4, 1, 0.0),
         ("skyrmion", (6, 6, 1), 0.3, 3, 0.2)]
MODES = [("2", None), ("3", None), ("1", "6"), ("2", "20")]
MODE_IDS = ["frontier", "global-only", "adaptive-cap6", "frontier-cap20"]


def _start(o, t, model, seed):
    if model == 1:
        rng = np.random.RandomState(seed)
        return rng.choice([-1.0, 1.0], size=t.N) * np.abs(t.S)
    return o.init_spins_philox(0.6, seed=seed)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("path", ["tables", "structured"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-m%d" % (c[0], "x".join(map(str, c[1])), c[3]))
def test_hybrid_wolff_selects_the_oracles_clusters_fp64(case, path, mode, monkeypatch):
    """40 steps against the oracle's FIFO restatement, whichever mixture of frontier growth and global passes ran them:
    same spins (1e-9) and the same attempt / accept / cluster-size counters."""
    from mcsolver_b200 import engine
    name, L, T, model, h = case
    monkeypatch.setenv("MCG_WOLFF_FRONTIER", mode[0])
    if mode[1]:
        monkeypatch.setenv("MCG_WOLFF_FRONTIER_CAP", mode[1])
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    o = util.oracle_system(t, h / T)
    if path == "tables":
        mk = lambda: engine.System.from_tables(t, precision=64, field=[h / T], seed=77)
    else:
        mk = lambda: engine.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=77)
    with mk() as s:
        start = _start(o, t, model, 77)
        s.set_spins(start)
        r = o.run(3, 40, 1, 1, seed=77, spins=start)
        s.wolff_steps(17)
        s.wolff_steps(24)      # two calls: the per-replica bookkeeping carries over
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9
        assert s.counters() == tuple(int(v) for v in r["counters"])
        nf = s.wolff_frontier_steps()
        if mode == ("2", None):
            assert nf == 41          # every cluster fits the default queue of these lattices
        if mode[0] == "3":
            assert nf == 0


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("case", [("square", (256, 256, 1), 1, [2.0, 2.269, 2.6, 3.2]), ("square", (256, 256, 1), 2, [0.7, 0.9, 1.3]),
                                  ("cubic", (32, 32, 64), 3, [1.2, 1.44, 2.0]), ("cri3", (192, 192, 1), 3, [30.0, 400.0])],
                         ids=["ising-256x256", "xy-256x256", "heisenberg-32x32x64", "cri3-192x192-residual"])
def test_adaptive_hybrid_reproduces_the_global_passes_at_size(case, prec, monkeypatch):
    """At sizes where the default is the adaptive hybrid (>= 65536 sites): several temperatures in one batch, some with
    percolating clusters (global passes), some with small ones (frontier growth).  The spins after 150 steps are the ones the
    plain global sequence produces - bit for bit where the reflection is always accepted (no residual); with a residual
    (CrI3: D and anisotropic J) the two paths sum it in different orders, so the comparison allows rounding."""
    from mcsolver_b200 import engine
    name, L, model, Ts = case
    spec = spec_of(name, L)
    R = len(Ts)
    out = {}
    for mode in ("0", None):
        if mode is None:
            monkeypatch.delenv("MCG_WOLFF_FRONTIER", raising=False)
        else:
            monkeypatch.setenv("MCG_WOLFF_FRONTIER", mode)
        with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1.0 / np.asarray(Ts), seed=9) as s:
            s.init_spins(0.0)
            s.metropolis_sweeps(30)
            s.wolff_steps(100)
            s.metropolis_sweeps(1)       # spins change behind the hybrid's back: projections must be rebuilt
            s.wolff_steps(50)
            out[mode] = ([s.get_spins(r) for r in range(R)], [s.counters(r) for r in range(R)],
                         [s.wolff_frontier_steps(r) for r in range(R)])
    a, b = out["0"], out[None]
    assert a[2] == [0] * R
    assert b[2][-1] > 100, b[2]            # the hottest replica grew (nearly) every cluster from its seed
    if name == "cri3":
        for r in range(R):
            assert np.max(np.abs(a[0][r] - b[0][r])) < (1e-9 if prec == 64 else 2e-3)
    else:
        for r in range(R):
            assert np.array_equal(a[0][r], b[0][r]), r
        assert a[1] == b[1]
