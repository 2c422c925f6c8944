"""sustained throughput of the dipole-stencil pass (sc 256^3 x 8 replicas, fp32): 3 x 80 measured sweeps after warm-up"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import add_dipole_stencil
import bench
spec = add_dipole_stencil(bench.cubic_spec(256), 0.1, 2.0)
R = 8
with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / bench.ladder(R), seed=1) as s:
    s.init_spins(0.0)
    s.timed_sweeps(40, with_measure=True)
    out = []
    for _ in range(3):
        ms = s.timed_sweeps(80, with_measure=True)
        out.append(R * spec.nsite * 80 / (ms * 1e-3))
print("NO_PAIR=%s  sustained attempts/s: %s" % (os.environ.get("MCG_ASYNC_NO_PAIR", ""), " ".join("%.3e" % v for v in out)), flush=True)
