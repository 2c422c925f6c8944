import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec, add_dipole_stencil
J = [-1, -1, -1] + [0] * 6
spec = add_dipole_stencil(LatticeSpec(L=(128, 128, 128), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)]), 0.1, 2.0)
R = 8
with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(1.2, 1.9, R), seed=1) as s:
    s.init_spins(0.3)
    s.timed_sweeps(2, with_measure=True)
