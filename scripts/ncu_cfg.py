"""A few measured sweeps of one bench config, for ncu captures:  CFG=<prefix of the config name in bench.config_table(), or 'headline'>
    ncu --set full --clock-control none --import-source on -k regex:<kernel> --launch-skip 8 -c 2 -o gpurun_out/x python scripts/ncu_cfg.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench

cfg = os.environ.get("CFG", "headline")
nsw = int(os.environ.get("SWEEPS", "4"))
if cfg == "headline":
    rows = [("headline", lambda: bench.cubic_spec(256), 3, 32, 8, bench.ladder, lambda n: np.zeros(n), 6, nsw)]
else:
    rows = [r for r in bench.config_table() if r[0].startswith(cfg)]
name, mk, model, prec, R, Tf, Hf, z, _ = rows[0]
spec = mk()
with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1 / np.asarray(Tf(R), float), field=Hf(R), seed=1) as s:
    s.init_spins(0.0)
    ms = s.timed_sweeps(nsw, with_measure=True)
    print(name, "key", s.jit_module_key(1) if s.jit_launch_count() else None, "ms/sweep", ms / nsw, "attempts/launch", R * spec.nsite / s.num_colours())
