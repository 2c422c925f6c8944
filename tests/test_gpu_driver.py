"""GPU tests of the headless entry: mcsolver_b200.loadMC(parameterfile) writes the reference's files."""
import numpy as np
import pytest

from tests import paramfiles, util

pytestmark = pytest.mark.gpu


def test_loadmc_xy_wolff_sample_layout_and_values(tmp_path):
    import mcsolver_b200
    f = tmp_path / "Square_XY"
    f.write_text(paramfiles.XY_SQUARE.format(L=16, T0=0.9, T1=0.9, nT=1, nthermal=4000, nsweep=16000, tau=1, model="XY", algo="Wolff"))
    res = mcsolver_b200.loadMC(str(f), workdir=str(tmp_path), precision=64, seed=3, quiet=True)
    lines = (tmp_path / "result.txt").read_text().splitlines()
    assert lines[0].startswith("#Temp          #Field         #<Si>") and len(lines) == 2 and len(lines[1]) == 150
    row = np.array([float(lines[1][15 * i:15 * (i + 1)]) for i in range(10)])
    assert row[0] == 0.9 and row[1] == 0.0
    out = (tmp_path / "out").read_text().splitlines()
    assert out[0] == "#T #H" and out[1].startswith("T= 9.000000E-01 h= 0.000000E+00 <Siz>=") and "<Q>=" in out[1]
    assert len((tmp_path / "spinDotSpin.txt").read_text().splitlines()) == 2
    # same physics as the reference's seeded runs of this point (stats fixture C1_xy_wolff, T=0.9)
    ref = [p for p in util.load_json("stats.json") if p["tag"] == "C1_xy_wolff"][0]
    rr = np.array(ref["rows"])
    e_ref, e_sd = (rr[:, 8] * 0.9).mean(), (rr[:, 8] * 0.9).std(ddof=1)
    assert abs(res["Energy"][0] - e_ref) < 5 * e_sd + 1e-3
    assert abs(res["U4"][0] - rr[:, 10].mean()) < 5 * rr[:, 10].std(ddof=1) + 1e-3


def test_loadmc_skyrmion_field_scan_with_frames(tmp_path):
    import mcsolver_b200
    f = tmp_path / "Skyrmion"
    f.write_text(paramfiles.SKYRMION_HEX.format(L=12, H0=0.0, H1=0.6, nH=3, frames=1, nthermal=500, nsweep=1000))
    res = mcsolver_b200.loadMC(str(f), workdir=str(tmp_path), precision=32, seed=1, quiet=True)
    assert np.allclose(res["H"], [0.0, 0.3, 0.6]) and np.all(res["T"] == 0.3)
    assert np.all(np.isfinite(res["TopoQ"])) and res["Energy"][2] < res["Energy"][0]      # field lowers the energy
    frames = sorted(p.name for p in tmp_path.glob("OnSpinDistribution.*"))
    assert frames == ["OnSpinDistribution.T0.300.H0.000.0.txt", "OnSpinDistribution.T0.300.H0.300.0.txt", "OnSpinDistribution.T0.300.H0.600.0.txt"]
    fr = np.loadtxt(tmp_path / frames[0])
    assert fr.shape == (288, 6) and np.allclose(np.linalg.norm(fr[:, 3:], axis=1), 1.0, atol=1e-5)


def test_shims_inside_a_forked_process_pool_like_win_py(tmp_path):
    """win.py:90-91,131-132 farms grid points over multiprocessing.Pool (fork).  The parent never touches
    the library, each forked worker creates its own CUDA context on first call: run exactly that pattern
    in a fresh interpreter (this pytest process may already hold a context, which must not be forked)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "pool.py"
    script.write_text('''
import sys, json
sys.path.insert(0, %r)
sys.path.insert(0, %r)
from multiprocessing import Pool
from mcsolver_b200.lattice import build_tables
from tests.specs import spec_of

def work(T):
    from heisenberglib import MCMainFunction          # the shim, imported inside the worker as mcMain.py does
    t = build_tables(spec_of("cubic", (6, 6, 6)), T, 3)
    out = MCMainFunction(*t.on_args(0, 50, 100, t.N, 0.0, 0.0, 0))
    return T, out[8] * T, out[10]

if __name__ == "__main__":
    with Pool(processes=3) as pool:
        res = sorted(pool.imap_unordered(work, [0.8, 1.4, 3.0]))
    print(json.dumps(res))
''' % (root, os.path.join(root, "mcsolver_b200", "lib")))
    out = subprocess.run([sys.executable, str(script)], check=True, capture_output=True, text=True, timeout=300).stdout
    import json
    res = json.loads(out.strip().splitlines()[-1])
    assert [r[0] for r in res] == [0.8, 1.4, 3.0]
    assert res[0][1] < res[1][1] < res[2][1] < 0          # energy rises with temperature
    assert res[0][2] > 0.99 and res[2][2] < 0.9            # U4: ordered vs disordered


def test_field_sweep_carries_the_configuration_and_shows_hysteresis():
    """Opt-in extension: ramping H up and down with state carry-over gives a hysteresis loop for an easy-axis
    ferromagnet at low T, whereas independent runs from the polarised (x) state (the reference's semantics) do not."""
    from mcsolver_b200 import scan
    from mcsolver_b200.lattice import LatticeSpec
    Jf = [-1.0, -1.0, -1.0] + [0.0] * 6
    spec = LatticeSpec(L=(12, 12, 1), S=[1.0], D=[[0.0, 0.0, -0.5]], bonds=[(0, 0, (1, 0, 0), Jf), (0, 0, (0, 1, 0), Jf)])
    up = np.linspace(-1.5, 1.5, 13)
    path = np.concatenate([up, up[::-1]])
    rows = scan.run_field_sweep(spec, 3, [0.2], path, 100, 100, precision=32, seed=2)
    mz = rows[:, 0, 25]                      # <Sz_tot>/nLat along the field axis (signed)
    m_up, m_down = mz[:13], mz[13:][::-1]    # same H grid, ascending and descending branch
    assert m_up[-1] > 0.9 and m_down[0] < -0.9          # saturated at both ends of the ramp
    # loop opens around H = 0: the descending branch stays magnetised where the ascending one is not yet
    k0 = 6
    assert m_down[k0] - m_up[k0] > 0.5


def test_loadmc_sharded_over_ranks_writes_the_one_gpu_files(tmp_path):
    """One process per rank (RANK / WORLD_SIZE / LOCAL_RANK as a launcher sets them; the GPU index wraps, so a one-GPU box runs both
    ranks on its GPU): each rank runs its share of the field scan, rank 0 collects the rows through files and writes result.txt,
    out, spinDotSpin.txt and the frames - byte for byte what a single process writes (streams follow the grid point, not the rank)."""
    import os
    import subprocess
    import sys
    import mcsolver_b200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    f = tmp_path / "Skyrmion"
    f.write_text(paramfiles.SKYRMION_HEX.format(L=12, H0=0.0, H1=0.6, nH=5, frames=1, nthermal=200, nsweep=400))
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    mcsolver_b200.loadMC(str(f), workdir=str(one), precision=32, seed=1, quiet=True)
    code = ("import sys; sys.path.insert(0, %r); import mcsolver_b200; mcsolver_b200.loadMC(%r, workdir=%r, precision=32, seed=1, quiet=True)"
            % (root, str(f), str(two)))
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o[-2000:]
    names = sorted(p.name for p in one.iterdir())
    assert names == sorted(p.name for p in two.iterdir()) and "result.txt" in names and len(names) == 3 + 5
    for n in names:
        assert (one / n).read_bytes() == (two / n).read_bytes(), n
