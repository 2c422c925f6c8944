import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(4096, 4096, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J)])
T = np.array([2.0, 2.2, 2.269, 2.35, 2.6])
with engine.System.from_spec(spec, 1, precision=32, nReplica=5, beta=1 / T, seed=1) as s:
    s.init_spins(0.0)
    s.metropolis_sweeps(20)
    s.wolff_steps(6)
