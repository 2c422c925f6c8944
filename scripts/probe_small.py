"""Wall time of the reference's own sample workloads through loadMC(): resident kernel vs launch-per-phase.
  C1 sample: samples/Square_XY_isotropic  (16x16 XY, Wolff, 40000 + 80000 sweeps, tau 1, 8 temperatures)
  C3 sample size: CrI3-like honeycomb 32x32x2 orbitals is not in tests/paramfiles; the skyrmion sample (16x16x2, Metropolis) stands in."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcsolver_b200
from tests import paramfiles

jobs = [
    ("C1 sample XY 16x16 Wolff 40000+80000 x 8 T", paramfiles.XY_SQUARE.format(L=16, T0=0.9, T1=1.2, nT=8, nthermal=40000, nsweep=80000, tau=1, model="XY", algo="Wolff")),
    ("XY 16x16 Metropolis 4000+8000 x 8 T", paramfiles.XY_SQUARE.format(L=16, T0=0.9, T1=1.2, nT=8, nthermal=4000, nsweep=8000, tau=0, model="XY", algo="Metropolis")),
    ("Skyrmion 16x16x2 Metropolis 4000+16000 x 8 H", paramfiles.SKYRMION_HEX.format(L=16, H0=0.0, H1=0.7, nH=8, frames=0, nthermal=4000, nsweep=16000)),
]
for name, text in jobs:
    for resident in (True, False):
        if resident:
            os.environ.pop("MCG_NO_RESIDENT", None)
        else:
            os.environ["MCG_NO_RESIDENT"] = "1"
        with tempfile.TemporaryDirectory() as d:
            f = os.path.join(d, "p")
            open(f, "w").write(text)
            mcsolver_b200.loadMC(f, workdir=d, precision=32, quiet=True)      # warm-up (context, pool)
            t0 = time.time()
            res = mcsolver_b200.loadMC(f, workdir=d, precision=32, quiet=True)
            dt = time.time() - t0
        print("%-50s resident=%d  %.3f s   E[0]=%.5f U4[0]=%.5f" % (name, resident, dt, res["Energy"][0], res["U4"][0]), flush=True)
