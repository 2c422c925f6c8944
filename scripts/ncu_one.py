import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
L = int(os.environ.get("L3", "256"))
spec = LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
R = int(os.environ.get("R", "8"))
T = np.linspace(1.0, 2.0, R)
meas = bool(int(os.environ.get("MEAS", "0")))
with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / T, seed=1) as s:
    s.init_spins(0.0)
    s.timed_sweeps(3, with_measure=meas)
