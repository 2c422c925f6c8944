"""GPU parity tests of the structured (descriptor-driven) path: same checks as the table path,
plus equality of the two paths' observables on the same configuration."""
import numpy as np
import pytest

from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu

CASES = [("skyrmion", (6, 6, 1), 0.3, 3, 0.2), ("skyrmion", (8, 12, 1), 0.3, 3, 0.2), ("cri3", (4, 4, 1), 35.0, 3, 0.0),
         ("cri3", (6, 6, 1), 35.0, 3, 0.0), ("aniso", (6, 6, 2), 0.7, 3, 0.3), ("aniso", (6, 6, 1), 0.7, 2, 0.3),
         ("square", (8, 8, 1), 0.9, 2, 0.05), ("square", (16, 8, 1), 0.9, 2, 0.05), ("cubic", (6, 6, 6), 1.4, 3, 0.1),
         ("cubic", (4, 6, 8), 1.4, 3, 0.1), ("cubic", (8, 8, 16), 1.4, 3, 0.0), ("square", (8, 8, 1), 2.3, 1, 0.05),
         ("cubic", (6, 6, 8), 4.4, 1, 0.0), ("cubic", (3, 3, 3), 1.4, 3, 0.0), ("square", (5, 5, 1), 0.9, 2, 0.0)]
IDS = ["%s-%s-m%d" % (c[0], "x".join(map(str, c[1])), c[3]) for c in CASES]


def _eng():
    from mcsolver_b200 import engine
    return engine


def _start(o, t, model, seed):
    if model == 1:
        rng = np.random.RandomState(seed)
        return rng.choice([-1.0, 1.0], size=t.N) * np.abs(t.S)
    return o.init_spins_philox(0.8, seed=seed)


def _check_jit_ran(s, jit_mode):
    """MCG_JIT=1 must really have run the specialised kernels wherever the lattice has vector items (V > 1: the grouped
    Philox layout tells); V == 1 lattices always take the scalar runtime-table kernel."""
    if jit_mode == "offline":
        assert s.jit_launch_count() == 0
    elif s.rng_layout()[1] > 1:
        assert s.jit_launch_count() > 0, "MCG_JIT=1 but the offline kernel ran"


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_structured_energy_observables_and_trajectory_fp64(case, jit_mode):
    eng = _eng()
    name, L, T, model, h = case
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    hT = h / T
    o = util.oracle_system(t, hT)
    # structured systems take UNSCALED couplings and beta = 1/T per replica
    with eng.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=4321) as s:
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        assert sorted(order.tolist()) == list(range(t.N))
        start = _start(o, t, model, 4321)
        if model != 1:
            s.init_spins(0.8)
            assert np.max(np.abs(s.get_spins() - start)) < 1e-12
        s.set_spins(start)
        E = s.energy()
        Eo = o.total_energy(start)
        assert abs(E - Eo) <= 1e-12 * max(1.0, abs(Eo))
        s.measure()
        out, _ = s.results()
        if model != 1:
            oo, _ = o.observe(start)
            for k in util.ON_CORE_SLOTS:
                assert abs(out[k] - oo[k]) <= 1e-11 * max(1.0, abs(oo[k])), (k, out[k], oo[k])
        s.reset_measurements()
        r = o.run(2, 9, 1, t.N, order=order, seed=4321, spins=start)
        s.metropolis_sweeps(10)
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9
        att, acc, _ = s.counters()
        assert (att, acc) == (int(r["counters"][0]), int(r["counters"][1]))
        _check_jit_ran(s, jit_mode)


@pytest.mark.parametrize("case", CASES[:11], ids=IDS[:11])
def test_structured_whole_run_fused_measurement_matches_oracle(case, jit_mode):
    """mcg_run on a structured system: the measurement sums come out of the colour passes
    themselves (fused); the result tuple must equal the oracle's restatement of the same loop."""
    eng = _eng()
    name, L, T, model, h = case
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    hT = h / T
    o = util.oracle_system(t, hT)
    with eng.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=17) as s:
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.init_spins(0.3)
        fr = s.run(0, 4, 15, 2 * t.N, spinFrame=3)
        out, grp = s.results()
        _check_jit_ran(s, jit_mode)
    r = o.run(2, 4, 15, 2 * t.N, flunc=0.3, spinFrame=3, order=order, seed=17)
    for k in util.ON_CORE_SLOTS + [7]:
        assert abs(out[k] - r["out"][k]) <= 1e-9 * max(1.0, abs(r["out"][k])), (k, out[k], r["out"][k])
    if t.nG:      # slot 28: orbital-group statistics from the fused class sums
        assert grp.shape == r["group"].shape
        assert np.max(np.abs(grp - r["group"]) / np.maximum(1.0, np.abs(r["group"]))) < 1e-9
    assert np.max(np.abs(fr[0] - r["frames"])) < 1e-9


@pytest.mark.parametrize("case", [CASES[1], CASES[8], CASES[11]], ids=[IDS[1], IDS[8], IDS[11]])
def test_structured_whole_run_with_a_fractional_interval_matches_oracle(case, jit_mode):
    """ninterval = 1.5 N on the structured path: one whole sweep, then a half sweep that carries the fused measurement."""
    eng = _eng()
    name, L, T, model, h = case
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    o = util.oracle_system(t, h / T)
    nint = t.N + t.N // 2
    with eng.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=17) as s:
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.init_spins(0.3)
        s.run(0, 3, 10, nint)
        out, _ = s.results()
        _check_jit_ran(s, jit_mode)
    r = o.run(2, 3, 10, nint, flunc=0.3, order=order, seed=17)
    slots = [0, 2, 4, 5, 8] if model == 1 else util.ON_CORE_SLOTS + [7]
    for k in slots:
        assert abs(out[k] - r["out"][k]) <= 1e-9 * max(1.0, abs(r["out"][k])), (k, out[k], r["out"][k])


def test_structured_replicas_are_independent_points_of_a_scan():
    """A batch of replicas = the reference's (T,H) grid: replica r of a batch gives exactly what a
    single-replica system with replica_offset=r gives (GPU-count independent streams)."""
    eng = _eng()
    spec = spec_of("cubic", (6, 6, 8))
    Ts = np.array([1.0, 1.4, 2.0])
    hs = np.array([0.0, 0.1, 0.2])
    with eng.System.from_spec(spec, 3, precision=64, nReplica=3, beta=1 / Ts, field=hs, seed=9) as s:
        s.init_spins(0.2)
        s.run(0, 3, 6, spec.nsite)
        batch = [s.results(r)[0] for r in range(3)]
        sp = [s.get_spins(r) for r in range(3)]
    for r in range(3):
        with eng.System.from_spec(spec, 3, precision=64, nReplica=1, beta=[1 / Ts[r]], field=[hs[r]], seed=9, replica_offset=r) as s:
            s.init_spins(0.2)
            s.run(0, 3, 6, spec.nsite)
            one = s.results(0)[0]
            assert np.array_equal(s.get_spins(0), sp[r])
        assert np.allclose(one, batch[r], rtol=1e-12, atol=1e-14)


def test_structured_fp32_vector_path_statistics():
    """fp32 float4 path on a lattice large enough for V=4: energy of the ordered state is exact,
    a short run stays normalised and close to the fp64 run's energy."""
    eng = _eng()
    spec = spec_of("cubic", (8, 8, 16))
    res = {}
    for prec in (32, 64):
        with eng.System.from_spec(spec, 3, precision=prec, beta=[1 / 1.2], seed=3) as s:
            s.init_spins(0.0)
            assert abs(s.energy() - (-3.0 * spec.nsite / 1.2)) < 1e-3
            s.run(0, 200, 400, spec.nsite)
            res[prec] = s.results()[0]
            n = np.linalg.norm(s.get_spins(), axis=1)
            assert np.allclose(n, 1.0, atol=3e-6)
    # <e> at T=1.2 on 8x8x16: statistical agreement (two different chains, ~1% tolerance)
    assert abs(res[32][8] - res[64][8]) < 0.02 * abs(res[64][8])
    assert abs(res[32][10] - res[64][10]) < 0.05


WOLFF_S = [("square", (8, 8, 1), 0.9, 2, 0.0), ("cubic", (6, 6, 8), 1.4, 3, 0.0), ("aniso", (6, 6, 1), 0.7, 3, 0.3),
           ("square", (8, 16, 1), 2.3, 1, 0.05), ("cubic", (6, 6, 6), 4.4, 1, 0.0), ("skyrmion", (6, 6, 1), 0.3, 3, 0.2)]


@pytest.mark.parametrize("case", WOLFF_S, ids=lambda c: "%s-%s-m%d" % (c[0], "x".join(map(str, c[1])), c[3]))
def test_structured_wolff_trajectory_matches_oracle_fp64(case):
    """Wolff on the structured path (neighbours computed, not stored) selects the same clusters as the
    oracle's FIFO growth: the bond uniforms are keyed by site-id pairs, not by layout."""
    eng = _eng()
    name, L, T, model, h = case
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    o = util.oracle_system(t, h / T)
    with eng.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=77) as s:
        start = _start(o, t, model, 77)
        s.set_spins(start)
        r = o.run(3, 30, 1, 1, seed=77, spins=start)
        s.wolff_steps(31)
        got = s.get_spins()
        assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9
        assert s.counters() == tuple(int(v) for v in r["counters"])


def test_structured_groups_in_cell_zero_only():
    """groupInSC = False: a group holds the listed orbitals of cell (0,0,0) only (Lattice.py:211)."""
    eng = _eng()
    spec = spec_of("aniso", (6, 6, 1), groupInSC=False)
    t = util.tables_for(dict(spec="aniso", L=(6, 6, 1), T=0.7, model=3))
    from mcsolver_b200.lattice import build_tables
    t = build_tables(spec, 0.7, 3)
    assert t.maxG == 1 and t.groups.tolist() == [[0], [1]]
    o = util.oracle_system(t, 0.3 / 0.7)
    with eng.System.from_spec(spec, 3, precision=64, beta=[1 / 0.7], field=[0.3], seed=5) as s:
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.init_spins(0.4)
        s.run(0, 2, 6, t.N)
        out, grp = s.results()
    r = o.run(2, 2, 6, t.N, flunc=0.4, order=order, seed=5)
    assert np.max(np.abs(grp - r["group"]) / np.maximum(1.0, np.abs(r["group"]))) < 1e-9


RG_CASES = [c for c in CASES if all(l == 1 or (l % 2 == 0 and l >= 6) for l in c[1])]


@pytest.mark.parametrize("case", RG_CASES, ids=["%s-%s-m%d" % (c[0], "x".join(map(str, c[1])), c[3]) for c in RG_CASES])
def test_structured_block_spin_statistics_match_table_path_and_oracle(case):
    """block_spin=True: tuple slots 11-19 (Ising 6, 7) from computed neighbours equal the table path, whose tables are the
    reference's own rOrb / rOrbCluster / linkedOrb_rnorm (heisenbergLib.c:255-286, 748-803), on one configuration and
    accumulated over a short run (same trajectory on both paths: same colouring order is not required for a set configuration)."""
    eng = _eng()
    name, L, T, model, h = case
    spec = spec_of(name, L)
    t = util.tables_for(dict(spec=name, L=L, T=T, model=model))
    hT = h / T
    o = util.oracle_system(t, hT)
    start = _start(o, t, model, 77)
    slots = [6, 7] if model == 1 else util.ON_RG_SLOTS
    with eng.System.from_tables(t, precision=64, field=[hT], seed=77) as s:
        s.set_spins(start)
        s.measure()
        ref, _ = s.results()
    with eng.System.from_spec(spec, model, precision=64, beta=[1.0 / T], field=[h], seed=77, block_spin=True) as s:
        s.set_spins(start)
        s.measure()
        out, _ = s.results()
        for k in slots:
            if np.isnan(ref[k]):
                assert np.isnan(out[k]), (k, out[k])
                continue
            assert abs(out[k] - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (k, out[k], ref[k])
        if model != 1:
            oo, _ = o.observe(start)
            for k in slots:
                if not np.isnan(oo[k]):
                    assert abs(out[k] - oo[k]) <= 1e-11 * max(1.0, abs(oo[k])), (k, out[k], oo[k])
        # accumulation over a whole run against the oracle's restatement of the loop
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.set_spins(start)
        s.reset_measurements()
        s.run(0, 2, 6, t.N)
        out, _ = s.results()
    if model != 1:
        r = o.run(2, 2, 6, t.N, order=order, seed=77, spins=start)
        for k in slots:
            if not np.isnan(r["out"][k]):
                assert abs(out[k] - r["out"][k]) <= 1e-9 * max(1.0, abs(r["out"][k])), (k, out[k], r["out"][k])


def test_structured_block_spin_rejects_supercells_the_descriptor_cannot_express():
    """L = 4 folds the doubled bonds +2d and -2d onto the same site (the reference merges them, Lattice.py:60-66) and odd L
    leaves the all-even sublattice: both are table-path cases, refused loudly rather than answered differently."""
    eng = _eng()
    for L in [(4, 6, 8), (5, 5, 1)]:
        with pytest.raises(eng.McgError, match="block_spin"):
            eng.System.from_spec(spec_of("cubic" if L[2] > 1 else "square", L), 3 if L[2] > 1 else 2, precision=64, block_spin=True)


@pytest.mark.parametrize("flunc", [0.3, 3.0], ids=["smooth", "rough"])
def test_packed_fp32_topological_charge_equals_the_fp64_evaluation(flunc, monkeypatch):
    """The fp32 specialised topological-charge kernel works on two cells at a time with packed arithmetic and a hand-written
    2 atan(im/re) (topo_pass.cuh: topo_cell_pair); the fp64 engine keeps the reference's double arithmetic (calcSignedArea,
    heisenbergLib.c:114-127; <= 1e-12 against the reference's known answers).  Same configuration in both: Q must agree to the
    fp32 rounding of the inputs, for smooth textures (small solid angles: relative accuracy) and rough ones (all quadrants)."""
    eng = _eng()
    monkeypatch.setenv("MCG_JIT", "1")
    spec = spec_of("skyrmion", (96, 64, 1))
    with eng.System.from_spec(spec, 3, precision=64, beta=[1 / 0.3], field=[0.2], seed=4) as d:
        d.init_spins(flunc)
        d.metropolis_sweeps(2)
        sp = d.get_spins()
        d.measure()
        q64 = d.results()[0][26]
    with eng.System.from_spec(spec, 3, precision=32, beta=[1 / 0.3], field=[0.2], seed=4) as s:
        s.set_spins(sp)
        s.measure()
        q32 = s.results()[0][26]
        assert s.jit_launch_count() > 0
    ntri = 4 * 96 * 64
    assert abs(q32 - q64) < 2e-6 * np.sqrt(ntri) + 1e-6 * abs(q64), (q32, q64)
    assert abs(q64) > 1e-3 or flunc < 1.0
