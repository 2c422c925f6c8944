"""Repeat one statistical fixture point with several seeds (is a 3-sigma excursion a fluctuation or a bias?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
from tests.specs import spec_of
from mcsolver_b200 import scan
tag, T, H = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
tables = len(sys.argv) > 4 and sys.argv[4] == "tables"
p = [q for q in util.load_json("stats.json") if q["tag"] == tag and abs(q["T"] - T) < 1e-9 and abs(q["H"] - H) < 1e-9][0]
ref = np.array(p["rows"])
spec = spec_of(p["spec"], tuple(p["L"]))
K = 8
slots = [0, 1, 2, 6, 8, 9, 10, 26] if p["model"] != 1 else [0, 1, 2, 4, 5, 8]
print("ref mean", np.array2string(ref[:, slots].mean(0), precision=5), "\nref se  ", np.array2string(ref[:, slots].std(0, ddof=1) / np.sqrt(len(ref)), precision=5))
allrows = []
for seed in range(100, 112):
    idx, rows, _ = scan.run_points(spec, p["model"], np.full(K, p["T"]), np.full(K, p["H"]), p["nthermal"], p["nsweep"], ninterval=p["ninterval"],
                                   algorithm=p["algo"], precision=64 if tables else 32, seed=seed, tables=tables)
    allrows.append(rows)
    print("seed", seed, np.array2string(rows[:, slots].mean(0), precision=5))
a = np.concatenate(allrows)
print("gpu mean", np.array2string(a[:, slots].mean(0), precision=5), "\ngpu se  ", np.array2string(a[:, slots].std(0, ddof=1) / np.sqrt(len(a)), precision=5))
