"""Wolff throughput on the C2 configuration: 2D Ising 4096^2 across Tc (cluster updates/s, flipped spins/s)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
L = int(os.environ.get("L2", "4096"))
spec = LatticeSpec(L=(L, L, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J)])
T = np.array([2.0, 2.2, 2.269, 2.35, 2.6])
R = len(T)
with engine.System.from_spec(spec, 1, precision=32, nReplica=R, beta=1 / T, seed=1) as s:
    s.init_spins(0.0)
    s.metropolis_sweeps(20)
    s.wolff_steps(20)
    s.reset_measurements()
    n = 200
    t0 = time.time()
    s.wolff_steps(n)
    dt = time.time() - t0
    for r in range(R):
        att, acc, cl = s.counters(r)
        print("T=%.3f: %d cluster updates, mean cluster size %.3e (%.2f%% of N)" % (T[r], att, cl / max(1, acc), 100 * cl / max(1, acc) / spec.nsite))
    tot = sum(s.counters(r)[2] for r in range(R))
    print("L=%d R=%d: %.1f cluster updates/s per replica-batch step (%.3f ms per step for %d replicas), %.3e flipped spins/s" % (
        L, R, n * R / dt, dt / n * 1e3, R, tot / dt))
    s.run(1, 0, 20, 5)
    print("U4:", [round(float(s.results(r)[0][8]), 4) for r in range(R)])
