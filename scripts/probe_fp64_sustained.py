"""sustained throughput of the fp64-state headline workload (sc 256^3 x 8 replicas): 3 x 300 measured sweeps after warm-up"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
spec = bench.cubic_spec(256)
R = 8
with engine.System.from_spec(spec, 3, precision=64, nReplica=R, beta=1 / bench.ladder(R), seed=1) as s:
    s.init_spins(0.0)
    s.timed_sweeps(300, with_measure=True)
    out = []
    for _ in range(3):
        ms = s.timed_sweeps(300, with_measure=True)
        out.append(R * spec.nsite * 300 / (ms * 1e-3))
    print("fp64 sustained attempts/s:", " ".join("%.3e" % v for v in out), " e/kT", s.results(0)[0][8], flush=True)
