"""fused M,E of the colour passes vs the recomputed energy/magnetisation of the final configuration (fp32 state, 256^3)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
spec = bench.cubic_spec(256); N = spec.nsite
T = np.array([1.0, 1.44, 2.5])
with engine.System.from_spec(spec, 3, precision=32, nReplica=3, beta=1 / T, seed=3) as s:
    s.init_spins(0.0)
    s.run(0, 10, 10, N)
    s.reset_measurements()
    s.run(0, 0, 1, N)
    for r in range(3):
        row = s.results(r)[0]
        E = s.energy(r)
        sp = s.get_spins(r).astype(np.float64)
        M = np.abs(sp.reshape(-1, 3).sum(0)) / N if sp.shape[-1] == 3 else None
        print("r=%d  E_fused/E_recomputed - 1 = %.3e   |<Sx>| fused %.9f  host fp64 sum %s" % (r, row[8] * N / E - 1, row[0], M))
