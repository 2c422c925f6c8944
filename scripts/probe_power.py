"""Does the sustained colour-pass rate sit on the board power cap?  Clocks/power vs run length."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from bench import cubic_spec, ClockSampler
spec = cubic_spec(256)
R = 8
T = np.linspace(1.0, 2.0, R)
with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / T, seed=1) as s:
    s.init_spins(0.0)
    s.timed_sweeps(3, with_measure=True)
    for nsw in (16, 100, 400, 1600, 1600):
        c = ClockSampler(0); c.start()
        ms = s.timed_sweeps(nsw, with_measure=True)
        clk = c.stop()
        print("sweeps=%4d  %.3f ms/sweep  %.3e attempts/s  clocks=%s" % (nsw, ms / nsw, R * spec.nsite * nsw / ms * 1e3, clk), flush=True)
