"""a few global-pass Wolff steps of 8 x 4096^2 Ising at T = 2.2 (percolating clusters), for ncu captures of k_wolff_bonds / k_wolff_flip"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
import bench
os.environ["MCG_WOLFF_FRONTIER"] = os.environ.get("MCG_WOLFF_FRONTIER", "3")
R = 8
with engine.System.from_spec(bench.square_spec(4096), 1, precision=32, nReplica=R, beta=np.full(R, 1 / 2.2), seed=1) as s:
    s.init_spins(0.0)
    s.metropolis_sweeps(20)
    s.wolff_steps(6)
    print("sites per launch", R * 4096 * 4096)
