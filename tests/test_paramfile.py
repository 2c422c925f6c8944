"""CPU tests of the parameter-file reader (mirror of fileio.py:169-292)."""
import os

import numpy as np
import pytest

from mcsolver_b200 import paramfile
from tests import paramfiles
from tests.specs import SPECS


def test_parse_authored_files(tmp_path):
    f = tmp_path / "xy"
    f.write_text(paramfiles.XY_SQUARE.format(L=16, T0=0.9, T1=1.2, nT=8, nthermal=40000, nsweep=80000, tau=1, model="XY", algo="Wolff"))
    p = paramfile.parse(str(f))
    assert (p.modelType, p.algorithm, p.LPack, p.nT, p.nH, p.ninterval, p.spinFrame) == ("XY", "Wolff", [16, 16, 1], 8, 1, 1, 0)
    assert p.groupInSC and p.orbGroupList == [[0]] and p.GcOrb == [[0, 0], [0, 0, 0]]
    T, H = p.grid()
    assert np.allclose(T, np.linspace(0.9, 1.2, 8)) and np.all(H == 0.0)
    s = p.spec()
    assert s.nsite == 256 and s.bonds[1][2] == (0, 1, 0) and s.bonds[0][3][:3] == [-1.0, -1.0, -1.0]
    f = tmp_path / "sk"
    f.write_text(paramfiles.SKYRMION_HEX.format(L=16, H0=0, H1=0.7, nH=16, frames=1, nthermal=40000, nsweep=80000))
    p = paramfile.parse(str(f))
    s = p.spec()
    ref = SPECS["skyrmion"]
    assert [list(b[3]) for b in s.bonds] == [list(map(float, b[3])) for b in ref["bonds"]]
    assert [b[2] for b in s.bonds] == [b[2] for b in ref["bonds"]]
    assert s.D == [[0.0, 0.0, -0.1]] * 2 and len(s.circuits) == 4 and s.circuits[2] == ((1, (0, 0, 0)), (0, (1, 0, 0)), (1, (1, 1, 0)))
    T, H = p.grid()
    assert len(T) == 16 and np.allclose(H, np.linspace(0, 0.7, 16)) and p.spinFrame == 1 and p.xAxisType == "H"


def test_missing_section_and_bad_version(tmp_path):
    txt = paramfiles.XY_SQUARE.format(L=4, T0=1, T1=1, nT=1, nthermal=1, nsweep=1, tau=0, model="XY", algo="Metropolis")
    f = tmp_path / "bad"
    f.write_text(txt.replace("Ncores:\n4\n", ""))
    with pytest.raises(ValueError):
        paramfile.parse(str(f))
    f.write_text(txt.replace("version: 3.0", "version: 2.0"))
    with pytest.raises(ValueError):
        paramfile.parse(str(f))


@pytest.mark.refhost
def test_same_values_as_reference_fileio_on_the_reference_samples():
    from oracle import refharness as rh
    if not rh.have_reference_host():
        pytest.skip("/root/reference not present")
    _, _, _, fileio = rh.load_reference_host()
    sdir = "/root/reference/samples"
    for name in ("Square_XY_isotropic", "CrI3With2NNCoupling", "SkyrmionOnHexLattice", "SkyrmionOnSqaureLattice"):
        path = os.path.join(sdir, name)
        assert fileio.loadParam(updateGUI=False, rpath=path)
        p = paramfile.parse(path)
        for k in ("LMatrix", "LPack", "pos", "S", "DList", "T0", "T1", "nT", "H0", "H1", "nH", "dipoleAlpha", "nthermal", "nsweep",
                  "ninterval", "xAxisType", "modelType", "algorithm", "GcOrb", "ncores", "spinFrame", "orbGroupList", "groupInSC"):
            assert getattr(p, k) == getattr(fileio, k), (name, k)
        assert [list(b) for b in p.bondList] == [list(b) for b in fileio.bondList], name
        assert [[(o, tuple(d)) for o, d in c] for c in p.localCircuitList] == [[(o, tuple(d)) for o, d in c] for c in fileio.localCircuitList]
