"""Ising spins as int8 (precision 8, csrc/ising8.cuh): one byte per spin, acceptance by integer thresholds of the fp64
probabilities.  The thresholds take the same decision as exp(corr) > u for every 32-bit word (isingLib.c:238-254 with the
engine's fp64 uniform), so the checks are the deterministic ones of the fp64 path: energy, trajectory, accept counters and
whole-run result tuple against the oracle's restatement."""
import numpy as np
import pytest

from mcsolver_b200.lattice import LatticeSpec
from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu
_J = [-1.0] + [0.0] * 8


def _eng():
    from mcsolver_b200 import engine
    return engine


def _spec(name, L, S=1.0, J=-1.0):
    Jv = [J] + [0.0] * 8
    if name == "square":
        return LatticeSpec(L=L, S=[S], bonds=[(0, 0, (1, 0, 0), Jv), (0, 0, (0, 1, 0), Jv)])
    if name == "cubic":
        return LatticeSpec(L=L, S=[S], bonds=[(0, 0, (1, 0, 0), Jv), (0, 0, (0, 1, 0), Jv), (0, 0, (0, 0, 1), Jv)])
    if name == "tri":     # triangular: 6 neighbours, 3 colours, period 3 - items of 4 sites, links with cz = +-1 and 0
        return LatticeSpec(L=L, S=[S], bonds=[(0, 0, (1, 0, 0), Jv), (0, 0, (0, 1, 0), Jv), (0, 0, (1, 1, 0), Jv)])
    raise KeyError(name)


# (lattice, L, T, h, S, J, expected sites per item)
CASES = [("square", (8, 32, 1), 2.3, 0.0, 1.0, -1.0, 16), ("square", (6, 8, 1), 2.3, 0.07, 1.0, -1.0, 4),
         ("cubic", (4, 6, 32), 4.4, 0.1, 1.0, -1.0, 16), ("cubic", (6, 6, 8), 4.4, 0.0, 1.5, -0.7, 4),
         ("square", (12, 64, 1), 2.0, -0.2, 1.0, 1.0, 16), ("tri", (6, 12, 1), 3.5, 0.05, 1.0, -1.0, 4)]
IDS = ["%s-%s-S%g-J%g" % (c[0], "x".join(map(str, c[1])), c[4], c[5]) for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
@pytest.mark.parametrize("p_att", [1.0, 0.37])
def test_int8_ising_energy_trajectory_and_counters_match_oracle(case, p_att, jit_mode):
    from mcsolver_b200.lattice import build_tables
    eng = _eng()
    name, L, T, h, S, J, V = case
    spec = _spec(name, L, S, J)
    t = build_tables(spec, T, 1)
    o = util.oracle_system(t, h / T)
    with eng.System.from_spec(spec, 1, precision=8, beta=[1.0 / T], field=[h], seed=4321) as s:
        assert s.rng_layout()[1] == V
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.init_spins(0.0)
        assert np.array_equal(s.get_spins(), np.full(t.N, S))
        start = np.random.RandomState(7).choice([-1.0, 1.0], size=t.N) * S
        s.set_spins(start)
        assert np.array_equal(s.get_spins(), start)
        E, Eo = s.energy(), o.total_energy(start)
        assert abs(E - Eo) <= 1e-12 * max(1.0, abs(Eo))
        nint = t.N if p_att == 1.0 else int(round(p_att * t.N))
        r = o.run(2, 9, 1, nint, order=order, seed=4321, spins=start)
        s.metropolis_sweeps(10, p_attempt=nint / t.N)
        got = s.get_spins()
        assert np.array_equal(got, r["spins"].reshape(got.shape))              # +-S exactly: same decision at every attempt
        att, acc, _ = s.counters()
        assert (att, acc) == (int(r["counters"][0]), int(r["counters"][1]))
        assert (s.jit_launch_count() > 0) == (jit_mode == "jit")      # both builds of the pass: runtime tables / lattice as literals


@pytest.mark.parametrize("case", CASES[:4], ids=IDS[:4])
def test_int8_ising_whole_run_with_fused_measurement_matches_oracle(case, jit_mode):
    from mcsolver_b200.lattice import build_tables
    eng = _eng()
    name, L, T, h, S, J, V = case
    spec = _spec(name, L, S, J)
    t = build_tables(spec, T, 1)
    o = util.oracle_system(t, h / T)
    with eng.System.from_spec(spec, 1, precision=8, beta=[1.0 / T], field=[h], seed=17) as s:
        order = s.colour_order()
        o.rng_layout = s.rng_layout()
        s.init_spins(0.0)
        fr = s.run(0, 4, 15, 2 * t.N, spinFrame=3)
        out, _ = s.results()
    r = o.run(2, 4, 15, 2 * t.N, spinFrame=3, order=order, seed=17)
    for k in (0, 1, 2, 3, 4, 5, 8, 9):       # 6, 7: block-spin energies (not computed at precision 8)
        assert abs(out[k] - r["out"][k]) <= 1e-10 * max(1.0, abs(r["out"][k])), (k, out[k], r["out"][k])
    assert np.array_equal(fr[0], r["frames"])


def test_int8_replicas_are_independent_scan_points_with_their_own_thresholds():
    eng = _eng()
    spec = _spec("square", (8, 32, 1))
    Ts, hs = np.array([1.5, 2.269, 3.5]), np.array([0.0, 0.1, -0.3])
    with eng.System.from_spec(spec, 1, precision=8, nReplica=3, beta=1 / Ts, field=hs, seed=9) as s:
        s.init_spins(0.0)
        s.run(0, 3, 6, spec.nsite)
        batch = [s.results(r)[0] for r in range(3)]
        sp = [s.get_spins(r) for r in range(3)]
    for r in range(3):
        with eng.System.from_spec(spec, 1, precision=8, nReplica=1, beta=[1 / Ts[r]], field=[hs[r]], seed=9, replica_offset=r) as s:
            s.init_spins(0.0)
            s.run(0, 3, 6, spec.nsite)
            assert np.array_equal(s.get_spins(0), sp[r])
            assert np.allclose(s.results(0)[0], batch[r], rtol=1e-12, atol=1e-14)


def test_int8_is_refused_where_the_byte_tricks_do_not_apply():
    eng = _eng()
    J2 = [-0.5] + [0.0] * 8
    two_J = LatticeSpec(L=(8, 32, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), _J), (0, 0, (0, 1, 0), J2)])
    with pytest.raises(eng.McgError, match="precision 8"):
        eng.System.from_spec(two_J, 1, precision=8)
    with pytest.raises(eng.McgError, match="precision 8"):
        eng.System.from_spec(spec_of("square", (8, 32, 1)), 2, precision=8)               # XY model
    with pytest.raises(eng.McgError, match="precision 8"):
        eng.System.from_spec(_spec("square", (8, 6, 1)), 1, precision=8)                  # 3 coarse cells along the vector axis
    with pytest.raises(eng.McgError, match="block_spin"):
        eng.System.from_spec(_spec("square", (8, 32, 1)), 1, precision=8, block_spin=True)
    with eng.System.from_spec(_spec("square", (8, 32, 1)), 1, precision=8) as s:
        s.init_spins(0.0)
        with pytest.raises(eng.McgError, match="Wolff"):
            s.wolff_steps(1)
