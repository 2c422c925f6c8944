"""Quick throughput probe of the structured Metropolis kernels (not the bench): prints attempts/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec

def probe(name, spec, model, prec, R, nsw=10, meas=False):
    T = np.linspace(1.0, 2.0, R)
    with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1 / T, seed=1) as s:
        s.init_spins(0.0)
        s.timed_sweeps(3, with_measure=meas)
        ms = s.timed_sweeps(nsw, with_measure=meas)
        att = R * spec.nsite * nsw
        w = (4 if prec == 32 else 8) * model
        print("%-28s prec=%d R=%d colours=%d meas=%d : %.3f ms/sweep  %.3e attempts/s  (%.0f GB/s at 3w=%dB)" % (
            name, prec, R, s.num_colours(), meas, ms / nsw, att / ms * 1e3, att / ms * 1e3 * 3 * w / 1e9, 3 * w), flush=True)

J = [-1, -1, -1] + [0] * 6
cubic = lambda L: LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
square = lambda L: LatticeSpec(L=(L, L, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J)])
L3 = int(os.environ.get("L3", "256"))
for meas in (False, True):
    probe("heis sc %d^3" % L3, cubic(L3), 3, 32, 8, meas=meas)
probe("heis sc %d^3" % L3, cubic(L3), 3, 32, 1)
probe("heis sc %d^3 fp64" % L3, cubic(L3), 3, 64, 2)
probe("xy sq 4096^2", square(4096), 2, 32, 4)
probe("ising sq 4096^2", square(4096), 1, 32, 4)
