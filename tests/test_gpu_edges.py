"""Edge cases of the boundary (the reference has no tests; these are the inputs its code mishandles or
handles implicitly): degenerate supercells, self images, ragged link counts, empty optional tables,
zero updates, frame capping, odd (non-bipartite) supercells."""
import numpy as np
import pytest

from mcsolver_b200.lattice import LatticeSpec, build_tables
from tests import util
from tests.specs import spec_of

pytestmark = pytest.mark.gpu
J = [-1.0, -0.7, -1.2] + [0.0] * 6


def _eng():
    from mcsolver_b200 import engine
    return engine


@pytest.mark.parametrize("L", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (1, 4, 1), (3, 3, 1), (5, 1, 1), (2, 2, 2)])
@pytest.mark.parametrize("model", [2, 3])
def test_degenerate_supercells_match_oracle_on_both_paths(L, model):
    """L=1: the bond's image is the site itself (one self link, Lattice.py:261); L=2: +d and -d reach the same
    neighbour (merged, Lattice.py:36-52); odd L: the ring is not bipartite (3 colours)."""
    eng = _eng()
    spec = LatticeSpec(L=L, S=[1.3], D=[[0.1, 0.0, -0.2]], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
    T = 0.8
    t = build_tables(spec, T, model)
    o = util.oracle_system(t, 0.2 / T)
    o.nR = 0                                          # block-spin tables are ill-defined on these supercells (reference: UB)
    start = o.init_spins_philox(1.0, seed=3)
    Eo = o.total_energy(start)
    for path in ("tables", "structured"):
        if path == "tables":
            s = eng.System.from_tables(t, precision=64, field=[0.2 / T], seed=3)
        else:
            s = eng.System.from_spec(spec, model, precision=64, beta=[1 / T], field=[0.2], seed=3)
        with s:
            s.set_spins(start)
            assert abs(s.energy() - Eo) <= 1e-12 * max(1.0, abs(Eo)), path
            order = s.colour_order()
            o.rng_layout = s.rng_layout()
            r = o.run(2, 5, 1, t.N, order=order, seed=3, spins=start)
            s.metropolis_sweeps(6)
            got = s.get_spins()
            assert np.max(np.abs(got - r["spins"].reshape(got.shape))) < 1e-9, path


def test_ragged_link_counts_and_empty_optional_tables():
    """Orbitals with different coordination (nlink ragged, -1 padded), no circuits, no groups, no block-spin tables."""
    eng = _eng()
    spec = LatticeSpec(L=(4, 4, 1), S=[1.0, 2.0, 0.5], bonds=[(0, 1, (0, 0, 0), J), (0, 1, (1, 0, 0), J), (1, 2, (0, 0, 0), J), (0, 0, (0, 1, 0), J)],
                       pair=(0, 2, (1, 1, 0)))
    t = build_tables(spec, 1.1, 3)
    assert len(set(t.nlink.tolist())) > 1 and (t.nbr == -1).any() and t.nG == 0 and t.tri.shape[0] == 0
    o = util.oracle_system(t, 0.0)
    from mcsolver_b200.engine import _TableArrays
    ta = _TableArrays(3, t.S, t.nlink, t.J, t.nbr, t.pairs, D=t.D)          # rOrb/rCl/rNbr/groups/tri all empty
    with eng.System.from_tables(ta, precision=64, seed=8) as s:
        s.init_spins(0.6)
        sp = s.get_spins()
        assert np.allclose(np.linalg.norm(sp, axis=1), np.abs(t.S))          # per-orbital spin lengths
        assert abs(s.energy() - o.total_energy(sp)) < 1e-12 * abs(o.total_energy(sp))
        s.run(0, 3, 4, t.N)
        out, g = s.results()
        assert np.all(out[11:20] == 0.0) and g.size == 2                     # no block-spin tables -> zeros; nG=0
    with eng.System.from_spec(spec, 3, precision=64, beta=[1 / 1.1], seed=8) as s2:
        s2.set_spins(sp)
        assert abs(s2.energy() - o.total_energy(sp)) < 1e-12 * abs(o.total_energy(sp))


def test_zero_updates_and_frame_capping():
    """ninterval=0 through the raw entry point = no updates at all (the reference's KAT mode); spinFrame larger
    than nsweep, or nsweep not divisible by spinFrame, never writes more than spinFrame frames (the reference
    overflows its tuple, heisenbergLib.c:645-675)."""
    eng = _eng()
    t = util.tables_for(dict(spec="square", L=(6, 6, 1), T=0.9, model=2))
    out = eng.run_on_args(2, t.on_args(0, 5, 3, 0, 0.0, 0.0, 3), seed=1, precision=64)
    fr = np.array(out[27])
    assert fr.shape == (3, 36, 3) and np.all(fr[:, :, 0] == 1.0) and np.all(fr[:, :, 1:] == 0.0)   # untouched polarised state
    assert out[8] == -2.0 / 0.9 and out[10] == 1.0
    out = eng.run_on_args(2, t.on_args(0, 2, 7, 36, 0.0, 0.0, 3), seed=1, precision=64)               # 7 // 3 = 2 -> sweeps 0,2,4 (6 capped)
    assert np.array(out[27]).shape == (3, 36, 3)
    out = eng.run_on_args(2, t.on_args(0, 2, 2, 36, 0.0, 0.0, 5), seed=1, precision=64)               # spinFrame > nsweep
    fr = np.array(out[27])
    assert fr.shape == (5, 36, 3) and np.all(fr[2:] == 0.0) and np.any(fr[1] != fr[0])


def test_invalid_inputs_are_rejected_with_messages():
    eng = _eng()
    t = util.tables_for(dict(spec="square", L=(4, 4, 1), T=1.0, model=3))
    good = list(t.on_args(0, 1, 1, 16, 0.0, 0.0, 0))
    for idx, bad, exc in [(7, tuple([99] * 16), eng.McgError),                     # nlink > maxL
                          (11, (0, 16), eng.McgError),                             # pair index out of range
                          (10, (0, 1, 400), eng.McgError),                         # circuit index out of range
                          (8, good[8][:-1], ValueError),                           # J table too short
                          (4, 0, eng.McgError),                                    # nsweep = 0
                          (3, 1.5, TypeError)]:                                    # non-integer count
        a = list(good)
        a[idx] = bad
        with pytest.raises(exc):
            eng.run_on_args(3, tuple(a))
    with pytest.raises(eng.McgError):
        eng.System.from_spec(LatticeSpec(L=(7, 7, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (1, 1, 0), J)]), 3,
                             precision=32, nReplica=0)


def test_prime_supercell_uses_full_period_or_reports():
    """L=7 (prime, >6): the colouring period must be the whole axis; L=17 has no admissible period <=16 for a
    nearest-neighbour ring, so the structured path refuses loudly and the table path (greedy colouring) works."""
    eng = _eng()
    sp7 = LatticeSpec(L=(7, 1, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J)])
    with eng.System.from_spec(sp7, 3, precision=64, seed=1) as s:
        assert s.num_colours() == 3
    sp17 = LatticeSpec(L=(17, 1, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J)])
    with pytest.raises(eng.McgError) as e:
        eng.System.from_spec(sp17, 3, precision=64, seed=1)
    assert "table path" in str(e.value)
    with eng.System.from_tables(build_tables(sp17, 1.0, 3), precision=64, seed=1) as s:
        assert s.num_colours() == 3
