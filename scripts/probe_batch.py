"""throughput of a bench config against the replica count per GPU:  CFG=<prefix of the config name> python scripts/probe_batch.py 8 16 32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
cfg = os.environ["CFG"]
name, mk, model, prec, R0, Tf, Hf, z, nsw = [r for r in bench.config_table() if r[0].startswith(cfg)][0]
spec = mk()
for R in [int(x) for x in sys.argv[1:]]:
    with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1 / np.asarray(Tf(R), float), field=Hf(R), seed=1) as s:
        s.init_spins(0.0)
        s.timed_sweeps(3, with_measure=True)
        ms = s.timed_sweeps(nsw, with_measure=True)
    print("%s  R=%3d  %.3e attempts/s" % (name[:40], R, R * spec.nsite * nsw / (ms * 1e-3)), flush=True)
