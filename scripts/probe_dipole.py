import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec, add_dipole_stencil
J = [-1, -1, -1] + [0] * 6
cu = lambda L: LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
spec = add_dipole_stencil(cu(128), 0.1, 2.0)
R = 8
cases = [("0", "4"), ("1", "4"), ("1", "2"), ("1", "1")] if os.environ.get("DIP_ALL") else [("0", "4")]
LL = int(os.environ.get("DIP_L", "128"))
MEAS = bool(int(os.environ.get("DIP_MEAS", "1")))
NSW = int(os.environ.get("DIP_SWEEPS", "8"))
spec = add_dipole_stencil(cu(LL), 0.1, 2.0)
for jit, minb in cases:
    os.environ["MCG_JIT"] = jit
    os.environ["MCG_JIT_MINB"] = minb
    with engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / np.linspace(1.2, 1.9, R), seed=1) as s:
        s.init_spins(0.0)
        s.timed_sweeps(2, with_measure=MEAS)
        ms = s.timed_sweeps(NSW, with_measure=MEAS) * 4 / NSW
        print("dipole r<=2 %d^3 meas=%d jit" % (LL, MEAS) + " sweeps=%d" % NSW + "=%s minb=%s colours=%d: %.3f ms/sweep %.3e attempts/s" % (jit, minb, s.num_colours(), ms / 4, R * spec.nsite * 4 / ms * 1e3), flush=True)
