"""Parity level 2 (north_star): equilibrium observables of the CUDA engine agree with the reference's
own CPU runs within 3 sigma.  tests/golden/stats.json holds K=8 independently seeded runs of the
reference's compiled C engines per (model, T, H) point (mean, sigma over seeds); here the engine
runs K=8 replicas (independent Philox streams) of the same point with the same sweep counts through
the C ABI and the two sample means are compared with the combined standard error:
    |mean_gpu - mean_ref| <= 3 * sqrt(sigma_gpu^2/K + sigma_ref^2/K) + tol_abs
autoCorr (slot 7) depends on the update dynamics (random site vs colour sweeps) and is not compared;
Ising <e> of the reference's Metropolis path is relative to an arbitrary zero and is
compared only for Wolff."""
import numpy as np
import pytest

from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu

STATS = util.load_json("stats.json")
IDS = ["%s-T%g-H%g" % (p["tag"], p["T"], p["H"]) for p in STATS]

# slots compared: O(n): <|Si|> (0..2), <SiSj> (6), e (8), e^2 (9), U4 (10), projections (20..25), Q (26)
ON_SLOTS = [0, 1, 2, 6, 8, 9, 10, 20, 22, 23, 25, 26]
ISING_SLOTS = [0, 2, 8, 9]


def _run_gpu(p, precision, tables):
    from mcsolver_b200 import engine, scan
    spec = spec_of(p["spec"], tuple(p["L"]))
    K = p["K"]
    T = np.full(K, p["T"])
    H = np.full(K, p["H"])
    idx, rows, _ = scan.run_points(spec, p["model"], T, H, p["nthermal"], p["nsweep"], ninterval=p["ninterval"],
                                   algorithm=p["algo"], precision=precision, seed=2024, tables=tables)
    return rows


@pytest.mark.parametrize("p", STATS, ids=IDS)
@pytest.mark.parametrize("path", ["structured-fp32", "tables-fp64"])
def test_equilibrium_observables_within_3_sigma_of_reference(p, path):
    if path == "structured-fp32" and p["algo"] == 1:
        pytest.skip("Wolff runs on the table path")
    rows = _run_gpu(p, 32 if path == "structured-fp32" else 64, tables=(path != "structured-fp32"))
    ref = np.array(p["rows"])
    K = p["K"]
    slots = ISING_SLOTS if p["model"] == 1 else ON_SLOTS
    if p["model"] == 1:
        slots = [0, 2, 8] + ([4, 5] if p["algo"] == 1 else [])
    bad = []
    for k in slots:
        g, r = rows[:, k], ref[:, k]
        if not np.all(np.isfinite(r)):
            continue
        se = np.sqrt(g.var(ddof=1) / K + r.var(ddof=1) / K)
        tol = 3.0 * se + 1e-9 + 2e-4 * max(abs(r.mean()), abs(g.mean()))   # the relative term covers fp32 state
        if abs(g.mean() - r.mean()) > tol:
            bad.append((k, g.mean(), r.mean(), se))
    # 12 slots x 3 sigma: allow one marginal excursion up to 4.5 sigma before calling it a failure
    hard = [b for b in bad if abs(b[1] - b[2]) > 4.5 * b[3] + 1e-9 + 2e-4 * max(abs(b[1]), abs(b[2]))]
    assert not hard and len(bad) <= 1, bad


def test_parallel_tempering_matches_independent_scan():
    """A PT run over an 8-temperature ladder (labels swap, configurations stay) gives the same
    <e>(T), <M>(T) as 8 independent runs, within statistical errors; swaps do happen."""
    from mcsolver_b200 import pt, scan
    spec = spec_of("cubic", (8, 8, 8))
    T = np.linspace(1.2, 1.9, 8)
    p = pt.ParallelTempering(spec, 3, T, precision=32, seed=5)
    rows = p.run(600, 3000, sweeps_per_swap=2)
    rates = p.swap_rates()
    p.close()
    assert np.all(rates > 0.05), rates
    # independent reference: 6 seeds per temperature -> mean and scatter
    ref = []
    for seed in range(1, 7):
        _, r, _ = scan.run_points(spec, 3, T, np.zeros(8), 600, 3000, precision=32, seed=100 + seed)
        ref.append(r)
    ref = np.array(ref)
    for k in (8, 10, 0):   # <e>, U4, <|Sx|>
        mu, sd = ref[:, :, k].mean(axis=0), ref[:, :, k].std(axis=0, ddof=1)
        tol = 4.0 * sd * np.sqrt(1 + 1 / 6.0) + 1e-3 * np.abs(mu) + 1e-6
        assert np.all(np.abs(rows[:, k] - mu) <= tol), (k, rows[:, k], mu, sd)
