"""Per-sweep cost of the resident kernel on the reference's sample sizes (time of s.run() alone, two sweep counts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import build_tables
from tests.specs import spec_of

only = sys.argv[1] if len(sys.argv) > 1 else None
cases = [("xy16-metro", "square", (16, 16, 1), 2, 0, 0.9), ("xy16-wolff", "square", (16, 16, 1), 2, 1, 0.9),
         ("skyr16-metro", "skyrmion", (16, 16, 1), 3, 0, 0.3), ("cri3-32-metro", "cri3", (32, 32, 1), 3, 0, 35.0),
         ("ising64-wolff", "square", (64, 64, 1), 1, 1, 2.3), ("cubic16-metro", "cubic", (16, 16, 16), 3, 0, 1.4),
         ("cubic20-metro", "cubic", (20, 20, 20), 3, 0, 1.4), ("xy90-wolff", "square", (90, 90, 1), 2, 1, 0.9),
         ("cri3-64-metro", "cri3", (64, 64, 1), 3, 0, 35.0), ("cubic32-metro", "cubic", (32, 32, 32), 3, 0, 1.4),
         ("xy256-wolff", "square", (256, 256, 1), 2, 1, 0.9), ("skyr128-metro", "skyrmion", (128, 128, 1), 3, 0, 0.3)]
R = 8
for tag, name, L, model, algo, T in cases:
    if only and only != tag:
        continue
    spec = spec_of(name, L)
    t = build_tables(spec, T, model)
    nint = t.N if algo == 0 else 1
    for resident in ([True] if only else [True, False]):
        if resident:
            os.environ.pop("MCG_NO_RESIDENT", None); os.environ.pop("MCG_NO_COOP", None)
        else:
            os.environ["MCG_NO_RESIDENT"] = "1"; os.environ["MCG_NO_COOP"] = "1"
        with engine.System.from_tables(t, precision=32, nReplica=R, beta=np.linspace(1.0, 0.8, R), seed=1) as s:
            s.init_spins(0.0)
            s.run(algo, 10, 200, nint)
            ts = []
            for n in (2000, 6000):
                t0 = time.time()
                s.run(algo, 0, n, nint)
                ts.append(time.time() - t0)
            per = (ts[1] - ts[0]) / 4000
            print("%-14s N=%-5d resident=%d  %.2f us per measured sweep (%d replicas)  attempts/s=%.3e" % (
                tag, t.N, resident, per * 1e6, R, (R * t.N / per) if algo == 0 else 0), flush=True)
