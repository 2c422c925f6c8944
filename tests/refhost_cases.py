"""The reference's own sample parameter files at test-sized sweep counts, shared by tests/golden/make_golden.py (which runs
the UNMODIFIED reference host + reference engines on them) and tests/test_gpu_refhost.py (which runs the same host against
our shims).  The samples are read from the reference tree or from its staged copy (oracle/_ref/host/samples); only the
numeric lines named here are replaced, in a temporary file."""
import re

# name -> edits (the line FOLLOWING the header that starts with the key is replaced / patched)
CASES = {
    "Square_XY_isotropic": dict(sweeps="2000 8000 1", ncores=4),                                    # XY 16^2, Wolff, 8 temperatures
    "SkyrmionOnHexLattice": dict(sweeps="500 2000 0", ncores=4, nH=4, frames=0),                    # Heisenberg + DMI, 16^2 x 2, H scan, Q
    "CrI3With2NNCoupling": dict(sweeps="800 3200 0", ncores=4, L="8 8 1", nT=5),                    # Heisenberg 1NN+2NN+3NN, D
}
COLUMNS = ["Temp", "Field", "Si", "Sj", "Susc", "Energy", "Capacity", "TopoQ", "U4", "AutoCorr"]


def edited(text, sweeps=None, ncores=None, L=None, nT=None, nH=None, frames=None):
    lines = text.split("\n")
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        key = ln.strip().split(" ")[0] if ln.strip() else ""
        if key.startswith("Sweeps") and sweeps is not None:
            out += [ln, sweeps]; i += 2; continue
        if key.startswith("Ncores") and ncores is not None:
            out += [ln, str(ncores)]; i += 2; continue
        if key.startswith("Supercell") and L is not None:
            out += [ln, L]; i += 2; continue
        if key.startswith("Tmin") and nT is not None:
            ln = re.sub(r"nT\s+\d+", "nT %d" % nT, ln)
        if key.startswith("Hmin") and nH is not None:
            ln = re.sub(r"nH\s+\d+", "nH %d" % nH, ln)
        if key.startswith("Distribution") and frames is not None:
            ln = re.sub(r"frame:\s*\d+", "frame: %d" % frames, ln)
        out.append(ln)
        i += 1
    return "\n".join(out)


def parse_result(text):
    """result.txt -> header line, rows sorted by (Field, Temp) (win.py writes them in task-completion order)."""
    lines = [ln for ln in text.split("\n") if ln.strip()]
    rows = sorted(([float(ln[15 * k:15 * (k + 1)]) for k in range(10)] for ln in lines[1:]), key=lambda r: (r[1], r[0]))
    return lines[0], rows
