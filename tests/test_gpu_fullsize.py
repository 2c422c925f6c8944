"""Full-size checks at BASELINE.json's configurations through size-independent properties (the oracle
cannot run 1.7e7 sites in seconds): exact ground-state energies, colouring structure, |s| conservation,
replica-permutation invariance, bit-exact dyadic Ising energy, determinism."""
import numpy as np
import pytest

from mcsolver_b200.lattice import LatticeSpec

pytestmark = pytest.mark.gpu

J = [-1.0, -1.0, -1.0] + [0.0] * 6


def _eng():
    from mcsolver_b200 import engine
    return engine


def test_c5_heisenberg_cubic_256_full_size():
    eng = _eng()
    spec = LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
    N = spec.nsite
    T = np.array([0.5, 1.443, 5.0])
    with eng.System.from_spec(spec, 3, precision=32, nReplica=3, beta=1 / T, seed=7) as s:
        assert s.num_colours() == 2
        s.init_spins(0.0)
        for r in range(3):     # polarised state: E = -3N*beta exactly (every partial sum is an integer multiple of beta)
            assert abs(s.energy(r) - (-3.0 * N / T[r])) <= 3e-6 * N / T[r]      # fp32 products, per-thread fp32 partial sums
        s.run(0, 10, 10, N)
        rows = [s.results(r)[0] for r in range(3)]
        a0, c0, _ = s.counters(0)
        assert a0 == 20 * N and 0 < c0 < a0
        # ordered / critical / disordered ordering of energy and magnetisation
        e = np.array([rw[8] * T[i] for i, rw in enumerate(rows)])
        m = np.array([np.linalg.norm(rw[0:3]) for rw in rows])
        assert e[0] < e[1] < e[2] < 0 and m[0] > m[1] > m[2] and m[0] > 0.8
        # energy accumulated by the fused colour passes == energy of the final configuration, recomputed
        E_fused_last = None
        s.reset_measurements()
        s.run(0, 0, 1, N)
        for r in range(3):
            E_fused_last = s.results(r)[0][8] * N
            assert abs(E_fused_last - s.energy(r)) <= 5e-8 * abs(s.energy(r)) + 1e-3
        sp = s.get_spins(2)
        nrm = np.linalg.norm(sp, axis=1)
        assert abs(nrm.mean() - 1.0) < 1e-6 and np.abs(nrm - 1.0).max() < 2e-5


def test_c5_with_dipole_stencil_256_full_size():
    """BASELINE config 5 as named: sc 256^3 with the dipole term (cut-off stencil r <= 2, 32 full-tensor links, 16
    colours).  Polarised state (S,0,0): every site has the same exactly known energy (lattice sum of the stencil); after a
    measured sweep the fused energy equals the recomputed one and |s| is conserved."""
    eng = _eng()
    from mcsolver_b200.lattice import add_dipole_stencil, dipole_tensor
    alpha = 0.1
    spec = add_dipole_stencil(LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)]), alpha, 2.0)
    N = spec.nsite
    T = np.array([0.9, 2.0])
    # energy per site of the polarised (S,0,0) state: sum over bond templates of J_xx (each unordered pair once)
    e_site = 0.0
    for b in spec.bonds:
        J9 = b[3]
        e_site += float(J9[0])
    with eng.System.from_spec(spec, 3, precision=32, nReplica=2, beta=1 / T, seed=5) as s:
        assert s.num_colours() == 16
        s.init_spins(0.0)
        for r in range(2):
            assert abs(s.energy(r) - e_site * N / T[r]) <= 5e-6 * abs(e_site) * N / T[r]
        s.metropolis_sweeps(2)
        s.reset_measurements()
        s.run(0, 0, 1, N)
        for r in range(2):
            E = s.energy(r)
            assert abs(s.results(r)[0][8] * N - E) <= 3e-6 * abs(E) + 1e-3
        a0, c0, _ = s.counters(0)
        assert a0 == N and 0 < c0 < a0      # counters restart with the measured run
        sp = s.get_spins(1)
        assert np.abs(np.linalg.norm(sp, axis=1) - 1.0).max() < 2e-5


def test_c2_ising_square_4096_bit_exact_dyadic_energy_and_determinism():
    eng = _eng()
    spec = LatticeSpec(L=(4096, 4096, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J)])
    N = spec.nsite
    # beta|J| = 0.5 is dyadic: every bond term is +-0.5 and every partial sum is exact in fp64, so the
    # energy equals (number of unsatisfied - satisfied bonds)/2 bit for bit, in any summation order
    finals = []
    for prec in (64, 64, 8, 8):      # fp64 planes; int8 planes (1 byte per spin, integer acceptance thresholds)
        with eng.System.from_spec(spec, 1, precision=prec, nReplica=1, beta=[0.5], seed=11) as s:
            s.init_spins(0.0)
            assert s.energy(0) == -2.0 * N * 0.5
            s.metropolis_sweeps(3)
            E = s.energy(0)
            sp = s.get_spins(0)
            x = sp.reshape(4096, 4096)
            bonds = (x * np.roll(x, -1, 0)).sum() + (x * np.roll(x, -1, 1)).sum()
            assert E == -0.5 * bonds                    # bit-exact
            finals.append(sp)
            a0, c0, _ = s.counters(0)
            assert a0 == 3 * N and 0 < c0 < a0
    assert np.array_equal(finals[0], finals[1])         # same seed, same trajectory
    assert np.array_equal(finals[2], finals[3])


def test_c4_skyrmion_hex_1024_topological_charge_of_uniform_and_tilted_states():
    eng = _eng()
    from tests.specs import spec_of
    spec = spec_of("skyrmion", (1024, 1024, 1))
    with eng.System.from_spec(spec, 3, precision=32, nReplica=1, beta=[1 / 0.3], field=[0.3], seed=2) as s:
        assert s.num_colours() == 2
        s.init_spins(0.0)
        s.measure()
        q0 = s.results(0)[0][26]
        assert abs(q0) < 1e-9                           # a uniform state covers no solid angle
        s.reset_measurements()
        s.run(0, 20, 5, spec.nsite)
        out = s.results(0)[0]
        assert np.isfinite(out[26]) and np.isfinite(out[8]) and out[25] > 0   # field along z polarises <Sz> > 0


def test_c3_cri3_512_colouring_and_energy():
    eng = _eng()
    from tests.specs import spec_of
    spec = spec_of("cri3", (512, 512, 1))
    with eng.System.from_spec(spec, 3, precision=32, nReplica=1, beta=[1 / 40.0], seed=2) as s:
        assert s.num_colours() == 8                     # 1NN+2NN+3NN honeycomb with period 2x2: all 8 classes adjacent
        s.init_spins(0.0)
        N = spec.nsite
        S = 1.5
        # polarised along x: per site 3*J1xx + 6*J2xx + 3*J3xx bonds (each counted once per site pair /2 *2) + Dx*S^2
        J1, J2, J3, D = -19.49182553875, -5.7387258225, 4.57737083, -3.12276251875
        e_site = (0.5 * (3 * J1 + 6 * J2 + 3 * J3) + D) * S * S / 40.0
        assert abs(s.energy(0) / N - e_site) < 1e-6 * abs(e_site)
        s.run(0, 5, 5, N)
        assert np.isfinite(s.results(0)[0][8])
