import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from tests.specs import spec_of
spec = spec_of("skyrmion", (1024, 1024, 1))
R = 16
for circ in (True, False):
    sp = spec if circ else spec_of("skyrmion", (1024, 1024, 1), circuits=())
    with engine.System.from_spec(sp, 3, precision=32, nReplica=R, beta=np.full(R, 1 / 0.3), field=np.linspace(0, 0.7, R), seed=1) as s:
        s.init_spins(0.0)
        for meas in (False, True):
            s.timed_sweeps(3, with_measure=meas)
            ms = s.timed_sweeps(10, with_measure=meas)
            print("circuits=%d meas=%d: %.3f ms/sweep %.3e attempts/s" % (circ, meas, ms / 10, R * sp.nsite * 10 / ms * 1e3), flush=True)
