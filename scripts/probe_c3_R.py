"""CrI3 512^2 x 2: throughput against the replica count (state size vs the 126 MB L2)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from tests.specs import spec_of
spec = spec_of("cri3", (512, 512, 1))
for R in (42, 48, 64, 96):
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(30, 50, R), seed=1) as s:
        s.init_spins(0.0)
        s.timed_sweeps(5, with_measure=True)
        ms = s.timed_sweeps(40, with_measure=True)
    att = R * spec.nsite * 40 / (ms * 1e-3)
    print("R=%2d state %5.1f MB  %.3e attempts/s  %.1f us per pass" % (R, R * spec.nsite * 12 / 1e6, att, ms / 40 / 8 * 1e3), flush=True)
