"""Throughput of the sc 256^3 colour pass against the replica count (= memory footprint): fp32 and fp64 state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
spec = bench.cubic_spec(int(os.environ.get("L3", 256)))
for prec, w in ((32, 36), (64, 72)):
    for R in (1, 2, 4, 8, 16):
        with engine.System.from_spec(spec, 3, precision=prec, nReplica=R, beta=1 / bench.ladder(R), seed=1) as s:
            s.init_spins(0.0)
            s.timed_sweeps(3, with_measure=True)
            ms = s.timed_sweeps(10, with_measure=True)
        att = R * spec.nsite * 10 / (ms * 1e-3)
        print("fp%d R=%2d  state %5.2f GB  %.3e attempts/s  %.1f%% of 6548 GB/s" % (prec, R, R * spec.nsite * 3 * prec / 8 / 1e9, att, 100 * att * w / 6548.2e9), flush=True)
