"""Throughput of every BASELINE.json config at its named size (informational; bench.py is the contract).
B_alg = (2 + min(C-1, z)) * w bytes per attempt; w = 4*NC bytes (fp32 planes)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcsolver_b200 import engine
from mcsolver_b200.lattice import LatticeSpec, add_dipole_stencil
from tests.specs import spec_of

J = [-1, -1, -1] + [0] * 6
PEAK = 6547.5e9
rows = []


ONLY = os.environ.get("ONLY", "")
PREC = int(os.environ.get("PREC", "32"))   # 64: fp64 state (w doubles)


def run(name, spec, model, R, T, H=None, z=None, nsw=10, meas=True):
    if ONLY and not any(name.startswith(o) for o in ONLY.split(",")):
        return
    spec = spec() if callable(spec) else spec
    H = np.zeros(R) if H is None else H
    with engine.System.from_spec(spec, model, precision=PREC, nReplica=R, beta=1 / np.asarray(T, float), field=H, seed=1) as s:
        C = s.num_colours()
        s.init_spins(0.0)
        s.timed_sweeps(3, with_measure=meas)
        ms = s.timed_sweeps(nsw, with_measure=meas)
        att = R * spec.nsite * nsw / (ms * 1e-3)
        w = (PREC // 8) * model
        balg = (2 + min(C - 1, z)) * w
        rows.append(dict(config=name, N=spec.nsite, replicas=R, colours=C, z=z, attempts_per_s=att, B_alg=balg, GBps=att * balg / 1e9,
                         roofline_frac=att * balg / PEAK, ms_per_sweep=ms / nsw))
        print("%-44s N=%9d R=%2d C=%d z=%2d : %.3e attempts/s  %6.0f GB/s (%4.1f%% of 6547.5) at %d B/attempt" % (
            name, spec.nsite, R, C, z, att, att * balg / 1e9, 100 * att * balg / PEAK, balg), flush=True)


sq = lambda L: LatticeSpec(L=(L, L, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J)])
cu = lambda L: LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
run("C1 XY square 4096^2 T-scan", lambda: sq(4096), 2, 8, np.linspace(0.9, 1.2, 8), z=4)
run("C2 Ising square 4096^2 T-scan", lambda: sq(4096), 1, 16, np.linspace(2.0, 2.6, 16), z=4)
run("C3 CrI3 honeycomb 512^2 (1NN+2NN+3NN, D)", lambda: spec_of("cri3", (512, 512, 1)), 3, int(os.environ.get("C3_R", 64)), np.linspace(30, 50, int(os.environ.get("C3_R", 64))), z=12)
run("C4 skyrmion hex 1024^2 (DMI, D, h, Q)", lambda: spec_of("skyrmion", (1024, 1024, 1)), 3, 16, np.full(16, 0.3), H=np.linspace(0, 0.7, 16), z=3)
run("C5 Heisenberg sc 256^3 T-scan", lambda: cu(256), 3, 8, 0.8 * 1.443 * (1.3 / 0.8) ** (np.arange(8) / 7), z=6)
run("C5 + dipole stencil r<=2 (32 links), 128^3", lambda: add_dipole_stencil(cu(128), 0.1, 2.0), 3, 8, np.linspace(1.2, 1.9, 8), z=32, nsw=4)
run("C5 + dipole stencil r<=2 (32 links), 256^3", lambda: add_dipole_stencil(cu(256), 0.1, 2.0), 3, 8, 0.8 * 1.443 * (1.3 / 0.8) ** (np.arange(8) / 7), z=32, nsw=8)
if not ONLY and PREC == 32:
    json.dump(rows, open("gpurun_out/configs_r1.json", "w"), indent=1)
