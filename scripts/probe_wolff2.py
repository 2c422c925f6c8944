"""Wolff on 2D Ising 4096^2: cluster updates/s per temperature, plain global sequence vs adaptive hybrid (R replicas of the SAME T)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from bench import square_spec

L = int(os.environ.get("L", 4096)); R = int(os.environ.get("R", 8)); steps = int(os.environ.get("STEPS", 200))
spec = square_spec(L)
for T in [2.0, 2.269, 2.35, 2.6, 3.0]:
    row = []
    for mode in ("0", "1"):
        os.environ["MCG_WOLFF_FRONTIER"] = mode
        with engine.System.from_spec(spec, 1, precision=32, nReplica=R, beta=np.full(R, 1 / T), seed=1) as s:
            s.init_spins(0.0)
            s.metropolis_sweeps(30)
            s.wolff_steps(30)
            c0 = [s.counters(r) for r in range(R)]
            t0 = time.time()
            s.wolff_steps(steps)
            dt = time.time() - t0
            c1 = [s.counters(r) for r in range(R)]
            fl = sum(b[2] - a[2] for a, b in zip(c0, c1))
            nf = sum(s.wolff_frontier_steps(r) for r in range(R))
            row.append((steps * R / dt, fl / (steps * R), nf))
    print("T=%.3f  global %.0f upd/s   hybrid %.0f upd/s (x%.1f)  mean cluster %.0f sites  frontier steps %d/%d" % (
        T, row[0][0], row[1][0], row[1][0] / row[0][0], row[1][1], row[1][2], (steps + 30) * R), flush=True)
