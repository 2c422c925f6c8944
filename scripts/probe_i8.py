"""int8 Ising pass at C2 size: device-timed sweeps, update-only and with the fused measurement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
R = int(os.environ.get("R", "16"))
L = int(os.environ.get("L2", "4096"))
spec = bench.square_spec(L)
T = np.linspace(2.0, 2.6, R)
for prec in (8, 32):
    with engine.System.from_spec(spec, 1, precision=prec, nReplica=R, beta=1 / T, seed=1) as s:
        s.init_spins(0.0)
        s.timed_sweeps(5, with_measure=True)
        for meas in (False, True):
            ms = s.timed_sweeps(40, with_measure=meas)
            att = R * spec.nsite * 40 / (ms * 1e-3)
            w = 1 if prec == 8 else 4
            print("prec %2d measure=%d: %.3e attempts/s  %.3f ms/sweep  %.0f GB/s at %d B (%.1f%% of 6547.5)" % (prec, meas, att, ms / 40, att * 3 * w / 1e9, 3 * w, att * 3 * w / 65.475e9), flush=True)
