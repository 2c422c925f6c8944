import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(256, 256, 256), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
R = 8
T = np.linspace(1.0, 2.0, R)
for var in ["", "B", "C", "D", "E", "F", "NOFAST"]:
    if var == "NOFAST":
        os.environ["MCG_NO_FAST"] = "1"
    else:
        os.environ.pop("MCG_NO_FAST", None)
        os.environ["MCG_FAST_VARIANT"] = var
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / T, seed=1) as s:
        s.init_spins(0.0)
        for meas in (False, True):
            s.timed_sweeps(3, with_measure=meas)
            ms = s.timed_sweeps(10, with_measure=meas)
            att = R * spec.nsite * 10
            print("variant %-6s meas=%d : %.3f ms/sweep %.3e attempts/s (%.1f%% of 6547 GB/s at 36 B)" % (
                var or "A", meas, ms / 10, att / ms * 1e3, att / ms * 1e3 * 36 / 6547.5e9 * 100), flush=True)
