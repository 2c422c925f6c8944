"""Cost of the opt-in block-spin statistics on the structured path (C5 lattice, 8 replicas)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
R = 8
for bs in (False, True):
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(1.2, 1.9, R), seed=1, block_spin=bs) as s:
        s.init_spins(0.0)
        s.timed_sweeps(3, with_measure=True)
        ms = s.timed_sweeps(10, with_measure=True)
        print("block_spin=%d: %.3f ms/sweep %.3e attempts/s" % (bs, ms / 10, R * spec.nsite * 10 / ms * 1e3), flush=True)
        out, _ = s.results(0)
        print("  slots 11-19:", np.array2string(out[11:20], precision=5))
