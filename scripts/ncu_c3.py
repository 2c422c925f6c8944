import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from tests.specs import spec_of
spec = spec_of("cri3", (512, 512, 1))
R = 21
with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(30, 50, R), seed=1) as s:
    s.init_spins(0.3)
    s.timed_sweeps(3, with_measure=True)
