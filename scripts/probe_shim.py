"""Drop-in boundary at the reference's own job sizes: one MCMainFunction call (Python tuples in, tuple out) through
mcsolver_b200/lib/*lib.py vs the reference's compiled C engine (oracle/_ref), same arguments, sweep counts of the
samples scaled by 1/100 (the reference needs ~20 min per point at full length)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200.lattice import build_tables
from mcsolver_b200 import engine
from oracle import refharness as rh
from tests.specs import spec_of

jobs = [("samples/CrI3With2NNCoupling 32x32x2, T=40, 800+6400 sweeps (1/100)", "cri3", (32, 32, 1), 3, 0, 40.0, 800, 6400, 0),
        ("samples/SkyrmionOnHexLattice 16x16x2, T=0.3, 400+800 sweeps (1/100)", "skyrmion", (16, 16, 1), 3, 0, 0.3, 400, 800, 0),
        ("samples/Square_XY_isotropic 16x16, T=0.9, Wolff 4000+8000 steps (1/10)", "square", (16, 16, 1), 2, 1, 0.9, 4000, 8000, 1)]
for name, spec_name, L, model, algo, T, nth, nsw, tau in jobs:
    t = build_tables(spec_of(spec_name, L), T, model)
    nint = t.N if tau == 0 else tau
    args = t.on_args(algo, nth, nsw, nint, 0.0, 0.0, 0)
    engine.run_on_args(model, args, seed=1, precision=32)          # context, pool, module load
    for prec in (32, 64):
        t0 = time.time(); out = engine.run_on_args(model, args, seed=1, precision=prec); dt = time.time() - t0
        print("%-72s gpu fp%d  %.3f s   <e>=%.5f U4=%.4f" % (name, prec, dt, out[8], out[10]), flush=True)
    if rh.have_ref_engine():
        t0 = time.time(); ref = rh.run_ref_engine(model, args, seed=1); dr = time.time() - t0
        print("%-72s reference (1 core) %.3f s   <e>=%.5f U4=%.4f   -> %.0fx" % (name, dr, ref[8], ref[10], dr / dt), flush=True)
