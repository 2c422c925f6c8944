import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
R = 8
T = np.linspace(1.0, 2.0, R)
for jit, minb in [("0", "4"), ("1", "4"), ("1", "3"), ("1", "2"), ("1", "5")]:
    os.environ["MCG_JIT"] = jit
    os.environ["MCG_JIT_MINB"] = minb
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / T, seed=1) as s:
        s.init_spins(0.0)
        for meas in (False, True):
            s.timed_sweeps(3, with_measure=meas)
            ms = s.timed_sweeps(16, with_measure=meas)
            att = R * spec.nsite * 16
            print("jit=%s minb=%s meas=%d : %.3f ms/sweep %.3e attempts/s (%.1f%% of 6547 GB/s at 36 B)" % (
                jit, minb, meas, ms / 16, att / ms * 1e3, att / ms * 1e3 * 36 / 6547.5e9 * 100), flush=True)
        s.run(0, 0, 1, spec.nsite)
        print("   <e>/kT replica0 = %.6f  U4 = %.6f" % (s.results(0)[0][8], s.results(0)[0][10]))
