import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from bench import cubic_spec, ladder
spec = cubic_spec(256); R = 8
T = ladder(R)
for it in range(6):
    t0 = time.time()
    s = engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1 / T, field=np.zeros(R), seed=1)
    t1 = time.time(); s.init_spins(0.0); s.energy(0)
    t2 = time.time(); s.run(0, 0, 100, spec.nsite)
    t3 = time.time(); rows = [s.results(r)[0] for r in range(R)]
    t4 = time.time(); s.close()
    t5 = time.time()
    print("create %.1f ms  init+sync %.1f ms  run %.1f ms  results %.1f ms  destroy %.1f ms  total %.1f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)), flush=True)
