"""The resident kernel (kernels_resident.cuh: whole MCMainFunction loop in one launch, one block per replica) against the
launch-per-phase path on the same tables: identical trajectory (same Philox counters, same per-site code), reductions
equal up to summation order."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

CASES = [("skyrmion", (6, 6, 1), 0.3, 3, 0.2), ("cri3", (4, 4, 1), 35.0, 3, 0.0), ("aniso", (6, 6, 2), 0.7, 3, 0.3),
         ("aniso", (6, 6, 1), 0.7, 2, 0.3), ("square", (8, 8, 1), 0.9, 2, 0.0), ("cubic", (6, 6, 6), 1.4, 3, 0.1),
         ("square", (8, 8, 1), 2.3, 1, 0.05), ("cubic", (6, 6, 6), 4.4, 1, 0.0)]
IDS = ["%s-%s-m%d" % (c[0], "x".join(map(str, c[1])), c[3]) for c in CASES]


def _mode(monkeypatch, mode):
    """mode: 'resident' (one block per replica), 'coop' (cooperative kernel, several blocks per replica), 'phases'."""
    for k in ("MCG_NO_RESIDENT", "MCG_NO_COOP", "MCG_RESIDENT_MAXN", "MCG_COOP_SITES"):
        monkeypatch.delenv(k, raising=False)
    if mode == "phases":
        monkeypatch.setenv("MCG_NO_RESIDENT", "1")
        monkeypatch.setenv("MCG_NO_COOP", "1")
    elif mode == "coop":
        monkeypatch.setenv("MCG_RESIDENT_MAXN", "0")
        monkeypatch.setenv("MCG_COOP_SITES", "24")


def _run(eng, t, model, hT, algo, ninterval, prec, R, monkeypatch, resident):
    _mode(monkeypatch, resident if isinstance(resident, str) else ("resident" if resident else "phases"))
    beta = np.linspace(1.0, 0.8, R)
    with eng.System.from_tables(t, precision=prec, nReplica=R, beta=beta, field=np.full(R, hT), seed=31) as s:
        s.init_spins(0.3)
        l0 = s.launch_count()
        frames = s.run(algo, 4, 12, ninterval, spinFrame=3)
        launches = s.launch_count() - l0
        res = [s.results(r) for r in range(R)]
        cnt = [s.counters(r) for r in range(R)]
        spins = [s.get_spins(r) for r in range(R)]
    return dict(frames=frames, res=res, cnt=cnt, spins=spins, launches=launches)


@pytest.mark.parametrize("case", CASES, ids=IDS)
@pytest.mark.parametrize("algo", [0, 1], ids=["metropolis", "wolff"])
@pytest.mark.parametrize("prec", [64, 32])
@pytest.mark.parametrize("mode", ["resident", "coop"])
def test_resident_kernel_equals_launch_per_phase(case, algo, prec, mode, monkeypatch):
    from mcsolver_b200 import engine as eng
    name, L, T, model, h = case
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    hT = h / T
    ninterval = t.N if algo == 0 else 2
    a = _run(eng, t, model, hT, algo, ninterval, prec, 3, monkeypatch, resident=mode)
    b = _run(eng, t, model, hT, algo, ninterval, prec, 3, monkeypatch, resident=False)
    assert a["launches"] <= 2 < b["launches"]          # one launch for thermalisation, one for the measured sweeps
    tol = 1e-11 if prec == 64 else 2e-5
    for r in range(3):
        assert a["cnt"][r] == b["cnt"][r]
        if prec == 64:
            assert np.max(np.abs(a["spins"][r] - b["spins"][r])) < 1e-12
            assert np.max(np.abs(a["frames"][r] - b["frames"][r])) < 1e-12
        (oa, ga), (ob, gb) = a["res"][r], b["res"][r]
        for k in range(len(oa)):
            if np.isnan(ob[k]):
                assert np.isnan(oa[k])
                continue
            assert abs(oa[k] - ob[k]) <= tol * max(1.0, abs(ob[k])), (k, oa[k], ob[k])
        if gb is not None and np.size(gb):
            assert np.max(np.abs(ga - gb) / np.maximum(1.0, np.abs(gb))) < tol


def test_resident_partial_sweeps_and_many_chunks(monkeypatch):
    """ninterval < N (attempt probability < 1) and more measured sweeps than one launch carries."""
    from mcsolver_b200 import engine as eng
    t = util.tables_for(dict(spec="square", L=(8, 8, 1), T=0.9, model=2))
    out = {}
    for resident in (True, False):
        _mode(monkeypatch, "resident" if resident else "phases")
        with eng.System.from_tables(t, precision=64, seed=8) as s:
            s.init_spins(0.2)
            s.run(0, 3, 40, 17)
            out[resident] = (s.results()[0], s.counters(), s.get_spins())
    assert out[True][1] == out[False][1]
    assert np.max(np.abs(out[True][2] - out[False][2])) < 1e-12
    assert np.max(np.abs(out[True][0] - out[False][0]) / np.maximum(1.0, np.abs(out[False][0]))) < 1e-11
    _mode(monkeypatch, "resident")
    t = util.tables_for(dict(spec="square", L=(8, 8, 1), T=2.3, model=1))
    with eng.System.from_tables(t, precision=32, seed=8) as s:
        l0 = s.launch_count()
        s.run(1, 10, 70000, 1)          # Wolff, tau = 1: 70000 measured sweeps -> two launches of <= 65536
        assert s.launch_count() - l0 == 3
        o, _ = s.results()
        assert abs(o[8] - 1.0) < 0.5 and s.counters()[0] == 70010   # U4 finite; every cluster step counted
