import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
R = 16
spec = bench.square_spec(4096)
T = np.linspace(2.0, 2.6, R)
with engine.System.from_spec(spec, 1, precision=8, nReplica=R, beta=1 / T, seed=1) as s:
    s.init_spins(0.0)
    s.timed_sweeps(6, with_measure=bool(int(os.environ.get("MEAS", "1"))))
