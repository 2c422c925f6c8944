"""Dipole-dipole coupling.  The reference defines it (Lattice.py:286-305: all ordered
pairs, open boundary, J = alpha/r^3 (1 - 3 r^r^T); Ising alpha/r^3) but cannot run it (TypeError in the
builder, and the engines dereference lattice[-1] on the padded block-spin table), so there is no
reference run to compare with: the link TABLES are pinned against the reference's own loop
(tests/test_oracle_golden.py), the physics against the oracle restatement on those tables."""
import numpy as np
import pytest

from mcsolver_b200.lattice import add_dipole_all_pairs, add_dipole_stencil, build_tables
from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu


def _eng():
    from mcsolver_b200 import engine
    return engine


@pytest.mark.parametrize("model", [1, 3])
def test_all_pairs_dipole_on_the_table_path_matches_oracle(model):
    eng = _eng()
    spec = spec_of("cubic", (3, 3, 2)) if model == 3 else spec_of("square", (4, 4, 1))
    T = 1.5
    t = add_dipole_all_pairs(spec, build_tables(spec, T, model), 0.4 / T)
    o = util.oracle_system(t, 0.0)
    o.nR = 0                                       # block-spin tables are undefined with appended links
    with eng.System.from_tables(t, precision=64, seed=5) as s:
        assert s.num_colours() == t.N              # complete graph: every site its own colour
        if model == 1:
            start = np.random.RandomState(3).choice([-1.0, 1.0], size=t.N)
        else:
            start = o.init_spins_philox(0.9, seed=5)
        s.set_spins(start)
        assert abs(s.energy() - o.total_energy(start)) <= 1e-12 * max(1.0, abs(o.total_energy(start)))
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        r = o.run(2, 7, 1, t.N, order=order, seed=5, spins=start)
        s.metropolis_sweeps(8)
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9


@pytest.mark.parametrize("model", [1, 3])
def test_wolff_with_duplicate_pair_links_matches_oracle_and_metropolis(model):
    """Tables from add_dipole_all_pairs hold TWO link slots for nearest neighbours (exchange bond + forceAdd'ed dipole
    link, Lattice.py:298).  Each slot is an independent bond of the cluster construction (heisenbergLib.c:355-366,
    isingLib.c:183-185): its uniform is keyed by the occurrence index as well, the clusters equal the oracle's FIFO
    growth step for step, and - the property that breaks if two slots share a uniform (pair added with max(p1,p2)
    instead of 1-(1-p1)(1-p2)) - the Wolff chain samples the same distribution as the Metropolis chain."""
    eng = _eng()
    spec = spec_of("cubic", (3, 3, 2)) if model == 3 else spec_of("square", (4, 4, 1))
    T = 2.2 if model == 1 else 1.3
    alpha = -0.6 if model == 1 else 0.0            # Ising: ferromagnetic 1/r^3 tail so that BOTH slots of a pair can activate
    t = build_tables(spec, T, model)
    if model == 1:
        t = add_dipole_all_pairs(spec, t, alpha / T)
    else:                                          # O(3): duplicate every exchange link by hand (isotropic: no residual)
        import dataclasses
        assert np.all(t.nlink == t.maxL)
        rNbr = np.full((t.rNbr.shape[0], 2 * t.maxL), -1, dtype=np.int32)
        rNbr[:, :t.maxL] = t.rNbr
        t = dataclasses.replace(t, maxL=2 * t.maxL, nbr=np.concatenate([t.nbr, t.nbr], axis=1), J=np.concatenate([0.5 * t.J, 0.5 * t.J], axis=1),
                                nlink=(2 * t.nlink).astype(np.int32), rNbr=rNbr)
    o = util.oracle_system(t, 0.0)
    o.nR = 0
    nb = np.asarray(t.nbr).reshape(t.N, -1)
    assert any(len(set(row[:n])) < n for row, n in zip(nb, np.asarray(t.nlink)))       # the tables do hold duplicate pairs
    with eng.System.from_tables(t, precision=64, seed=8) as s:
        start = np.random.RandomState(3).choice([-1.0, 1.0], size=t.N) if model == 1 else o.init_spins_philox(0.9, seed=8)
        s.set_spins(start)
        r = o.run(3, 40, 1, 1, seed=8, spins=start)
        s.wolff_steps(41)
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9
        assert s.counters() == tuple(int(v) for v in r["counters"])
    # detailed balance: K independent Wolff chains against K Metropolis chains of the same tables
    K = 32
    rows = {}
    for algo in (0, 1):
        with eng.System.from_tables(t, precision=64, nReplica=K, seed=77 + algo) as s:
            s.init_spins(0.0)
            s.run(algo, 2000, 20000, t.N if algo == 0 else 2)
            rows[algo] = np.array([s.results(k)[0] for k in range(K)])
    eslot, mslot = (4, 0) if model == 1 else (8, 10)
    for k in (eslot, mslot):
        if model == 1 and k == eslot:
            continue                               # isingLib's Metropolis energy is relative to the start
        a, b = rows[0][:, k], rows[1][:, k]
        se = np.sqrt(a.var(ddof=1) / K + b.var(ddof=1) / K)
        assert abs(a.mean() - b.mean()) <= 4.0 * se + 1e-9, (k, a.mean(), b.mean(), se)


@pytest.mark.parametrize("model,L", [(3, (8, 8, 8)), (3, (12, 6, 6)), (1, (8, 8, 8))])
def test_dipole_stencil_on_the_structured_path_matches_oracle(model, L):
    """Cut-off periodic dipole stencil (32 neighbours on sc within r<=2) = ordinary bond templates for the
    structured engine: colouring period found automatically, energy and trajectory equal the oracle's."""
    eng = _eng()
    T = 1.4
    spec = add_dipole_stencil(spec_of("cubic", L), 0.3, 2.0, ising=(model == 1))
    t = build_tables(spec, T, model)
    assert t.maxL == 32
    o = util.oracle_system(t, 0.0)
    with eng.System.from_spec(spec, model, precision=64, beta=[1 / T], seed=9) as s:
        assert s.num_colours() >= 8                # sites within distance 2 must all differ in colour
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        if model == 1:
            start = np.random.RandomState(3).choice([-1.0, 1.0], size=t.N)
        else:
            start = o.init_spins_philox(0.9, seed=9)
        s.set_spins(start)
        E, Eo = s.energy(), o.total_energy(start)
        assert abs(E - Eo) <= 1e-12 * max(1.0, abs(Eo))
        r = o.run(2, 5, 1, t.N, order=order, seed=9, spins=start)
        s.metropolis_sweeps(6)
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9


def test_dipole_stencil_fp32_jit_path_runs_and_lowers_symmetry():
    """With a dipole term the Heisenberg model is no longer isotropic: thin-film geometry (Lz = 1 layer
    thick slab is not periodic here, so use anisotropic box) still gives finite, normalised results on the
    fp32 vector path (V=4) with 32 links per site."""
    eng = _eng()
    spec = add_dipole_stencil(spec_of("cubic", (16, 16, 16)), 0.2, 2.0)
    with eng.System.from_spec(spec, 3, precision=32, nReplica=2, beta=[1 / 1.0, 1 / 2.0], seed=4) as s:
        s.init_spins(0.0)
        s.run(0, 50, 100, spec.nsite)
        for r in range(2):
            out = s.results(r)[0]
            assert np.all(np.isfinite(out[:11]))
        sp = s.get_spins(0)
        assert np.allclose(np.linalg.norm(sp, axis=1), 1.0, atol=2e-5)
        assert s.results(0)[0][8] * 1.0 < s.results(1)[0][8] * 2.0      # energy (in K) rises with temperature


@pytest.mark.parametrize("L", [(16, 16, 16), (8, 12, 32)])
def test_dipole_stencil_fp32_fused_energy_equals_recomputed_energy(L):
    """The fp32 vector pass with 32 full-tensor links accumulates the bond energy from the field of the lower-colour
    neighbours (snapshot taken between the two halves of the sorted link list).  After one measured sweep that sum
    must equal the energy of the final configuration recomputed by the measurement-only kernel, and the fp64 engine
    started from the same configuration must agree on it."""
    eng = _eng()
    spec = add_dipole_stencil(spec_of("cubic", L), 0.3, 2.0)
    N = spec.nsite
    T = np.array([0.7, 1.4, 3.0])
    with eng.System.from_spec(spec, 3, precision=32, nReplica=3, beta=1 / T, seed=11) as s:
        assert s.num_colours() >= 8
        s.init_spins(0.4)
        s.metropolis_sweeps(3)
        s.reset_measurements()
        s.run(0, 0, 1, N)
        for r in range(3):
            E_fused = s.results(r)[0][8] * N
            E = s.energy(r)
            assert abs(E_fused - E) <= 3e-6 * abs(E) + 1e-3
            sp = s.get_spins(r)
            assert np.abs(np.linalg.norm(sp, axis=1) - 1.0).max() < 2e-5
            with eng.System.from_spec(spec, 3, precision=64, beta=[1 / T[r]], seed=11) as d:
                d.set_spins(sp)
                assert abs(d.energy() - E) <= 3e-6 * abs(E) + 1e-3


def test_async_link_pipeline_reproduces_the_synchronous_pass_bit_for_bit(monkeypatch):
    """k_struct_async (cp.async stages, FFMA2 field sums) and the synchronous runtime-table pass evaluate the same
    multiply-add chain over the same link order with the same Philox words: identical fp32 trajectories, accept counters
    and fused measurements."""
    eng = _eng()
    spec = add_dipole_stencil(spec_of("cubic", (16, 8, 32)), 0.25, 2.0)
    N = spec.nsite
    T = np.array([0.8, 1.5, 2.5, 4.0])

    def go():
        with eng.System.from_spec(spec, 3, precision=32, nReplica=4, beta=1 / T, field=np.array([0.0, 0.1, 0.0, 0.3]), seed=21) as s:
            s.init_spins(0.5)
            s.metropolis_sweeps(3)
            s.metropolis_sweeps(2, p_attempt=0.37)
            s.reset_measurements()
            s.run(0, 0, 4, N)
            return [s.get_spins(r).copy() for r in range(4)], [s.counters(r) for r in range(4)], [s.results(r)[0][:11].copy() for r in range(4)]

    monkeypatch.delenv("MCG_NO_ASYNC", raising=False)
    a = go()
    monkeypatch.setenv("MCG_NO_ASYNC", "1")
    b = go()
    for r in range(4):
        assert np.array_equal(a[0][r], b[0][r])
        assert a[1][r] == b[1][r]
        assert np.allclose(a[2][r], b[2][r], rtol=1e-12, atol=1e-12)
