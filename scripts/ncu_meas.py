"""Workload for ncu captures of the measurement kernels: C4 topological charge and C5 block-spin statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
from tests.specs import spec_of
which = sys.argv[1]
if which == "c4":
    spec = spec_of("skyrmion", (1024, 1024, 1))
    R = 16
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=np.full(R, 1 / 0.3), field=np.linspace(0, 0.7, R), seed=1) as s:
        s.init_spins(0.5)
        s.timed_sweeps(3, with_measure=True)
else:
    J = [-1, -1, -1] + [0] * 6
    spec = LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
    R = 8
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(1.2, 1.9, R), seed=1, block_spin=True) as s:
        s.init_spins(0.5)
        s.timed_sweeps(3, with_measure=True)
