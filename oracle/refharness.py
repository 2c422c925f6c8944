"""TEST INFRASTRUCTURE ONLY - drives the UNMODIFIED reference (golddoushi/mcsolver) in this container.

Used by tests/golden/make_golden.py to generate the committed fixtures and by the (container-only)
pin tests.  It needs /root/reference (host Python: Lattice.py, mcMain.py) and oracle/_ref (the
reference C engines compiled by oracle/Makefile).  /root/reference does not exist on the GPU box:
nothing in `-m gpu` tests, smoke() or bench.py imports this module's reference-host half; the
`load_ref_engine` half (oracle/_ref only) is what `bench.py --impl reference` uses.

The product path (mcsolver_b200/) never imports anything from oracle/.
"""
import ctypes
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference/mcsolver"
HOST_DIR = os.path.join(REF_DIR, "host")          # sourceless bytecode of the reference's host modules (oracle/Makefile: host)

_libc = None


def srand(k):
    """glibc srand(k): the reference never seeds rand(); we do it from outside."""
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL("libc.so.6")
    _libc.srand(ctypes.c_uint(k))


def have_ref_engine():
    return os.path.isdir(REF_DIR) and any(f.startswith("heisenberglib") for f in os.listdir(REF_DIR))


def have_reference_host():
    """The reference's host Python is importable: from /root/reference (build container) or from the compiled copy."""
    return os.path.isfile(os.path.join(REFERENCE_SRC, "Lattice.py")) or os.path.isfile(os.path.join(HOST_DIR, "Lattice.mcb"))


def reference_host_dir():
    return REFERENCE_SRC if os.path.isfile(os.path.join(REFERENCE_SRC, "Lattice.py")) else HOST_DIR


class _HostFinder:
    """Imports the staged host modules (oracle/_ref/host/<name>.mcb = sourceless bytecode) under their script-mode names."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        f = os.path.join(HOST_DIR, name + ".mcb")
        if "." in name or not os.path.isfile(f):
            return None
        return importlib.util.spec_from_loader(name, importlib.machinery.SourcelessFileLoader(name, f))


def sample_file(name):
    """Path of one of the reference's sample parameter files (samples/<name>)."""
    for d in (os.path.join(os.path.dirname(REFERENCE_SRC), "samples"), os.path.join(HOST_DIR, "samples")):
        if os.path.isfile(os.path.join(d, name)):
            return os.path.join(d, name)
    raise FileNotFoundError(name)


def load_ref_engine(name):
    """Import oracle/_ref/<name> (isinglib | xylib | heisenberglib): the reference's compiled C."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # the product shims carry the same module names; make sure we get the compiled one
    mod = sys.modules.get(name)
    if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
        del sys.modules[name]
    import importlib.machinery
    import importlib.util
    for f in os.listdir(REF_DIR):
        if f.startswith(name + ".") and f.endswith(".so"):
            loader = importlib.machinery.ExtensionFileLoader(name, os.path.join(REF_DIR, f))
            spec = importlib.util.spec_from_loader(name, loader)
            m = importlib.util.module_from_spec(spec)
            loader.exec_module(m)
            return m
    raise ImportError("oracle/_ref/%s*.so not built (run make -f oracle/Makefile)" % name)


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


def _stub(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda n: _Dummy  # permissive
    sys.modules[name] = m
    return m


def load_reference_host():
    """Import the reference's Lattice / mcMain / win / fileio with tkinter+matplotlib stubbed
    (both absent in this image).  Returns (Lattice, mcMain, win, fileio)."""
    for n in ("tkinter", "tkinter.filedialog", "tkinter.ttk", "matplotlib", "matplotlib.pyplot",
              "matplotlib.figure", "matplotlib.backends", "matplotlib.backends.backend_tkagg",
              "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d"):
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:
                _stub(n)
    src = reference_host_dir()
    if src == HOST_DIR:
        if not any(f is _HostFinder for f in sys.meta_path):
            sys.meta_path.append(_HostFinder)      # after the path finders: real modules of the same name win
    elif src not in sys.path:
        sys.path.insert(0, src)
    import Lattice
    import mcMain
    import win
    import fileio
    return Lattice, mcMain, win, fileio


class CapturedArgs(Exception):
    pass


def reference_tables(LMatrix, pos, S, D, bonds, T, L, ki=(0, 0, (0, 0, 0)), orbGroupList=(),
                     groupInSC=False, h=0.0, On=3, spinFrame=0, circuits=(), algo="Metropolis",
                     nsweep=1, nthermal=0, ninterval=0, flunc=0.0, dipoleAlpha=0.0):
    """Build the positional argument tuple the reference's mcMain.py would hand to MCMainFunction,
    by running the reference's own MC.__init__ + mainLoopViaCLib[_On] with the engine import
    intercepted.  bonds: list of (src, tgt, (n1,n2,n3), J9) with J9 in the reference order
    xx,yy,zz,xy,xz,yz,yx,zx,zy (Ising: J9[0] used)."""
    import numpy as np
    Lattice, mcMain, win, fileio = load_reference_host()
    bondList = [Lattice.Bond(b[0], b[1], np.array([int(x) for x in b[2]]), *[float(v) for v in b[3]],
                             True if On != 1 else False) for b in bonds]
    mc = mcMain.MC(0, np.array(LMatrix, dtype=float), pos=np.array(pos, dtype=float), S=list(S),
                   D=[list(d) for d in D], bondList=bondList, T=T, Lx=L[0], Ly=L[1], Lz=L[2],
                   ki_s=ki[0], ki_t=ki[1], ki_overLat=list(ki[2]), orbGroupList=list(orbGroupList),
                   groupInSC=groupInSC, h=h, dipoleAlpha=dipoleAlpha, On=On, spinFrame=spinFrame,
                   localCircuitList=list(circuits))
    captured = {}

    def fake(*args):
        captured["args"] = args
        raise CapturedArgs()

    name = {1: "isinglib", 2: "xylib", 3: "heisenberglib"}[On]
    fake_mod = types.ModuleType(name)
    fake_mod.MCMainFunction = fake
    saved = sys.modules.get(name)
    sys.modules[name] = fake_mod
    try:
        if On == 1:
            mc.mainLoopViaCLib(nsweep=nsweep, nthermal=nthermal, ninterval=ninterval, algo=algo)
        else:
            mc.mainLoopViaCLib_On(nsweep=nsweep, nthermal=nthermal, ninterval=ninterval, algo=algo,
                                  On=On, flunc=flunc)
    except CapturedArgs:
        pass
    finally:
        if saved is not None:
            sys.modules[name] = saved
        else:
            del sys.modules[name]
    return captured["args"]


def patch_reference_dipole():
    """Runtime shim that makes the reference's OWN dipole loop (Lattice.py:286-308) runnable for Ising:
    it calls addLinking(target, J, forceAdd=True) without the required `distance` argument (TypeError at
    Lattice.py:298 vs :34).  We wrap the method to default that argument; no reference source is modified."""
    Lattice, _, _, _ = load_reference_host()
    if getattr(Lattice.Orbital, "_mcg_patched", False):
        return
    orig = Lattice.Orbital.addLinking

    def addLinking(self, targetOrb, strength, distance=None, quiet=False, forceAdd=False):
        return orig(self, targetOrb, strength, distance, quiet=quiet, forceAdd=forceAdd)
    Lattice.Orbital.addLinking = addLinking
    Lattice.Orbital._mcg_patched = True


def run_ref_engine(On, args, seed=None):
    """Call the reference's compiled MCMainFunction on a positional tuple; srand(seed) first."""
    name = {1: "isinglib", 2: "xylib", 3: "heisenberglib"}[On]
    mod = load_ref_engine(name)
    if seed is not None:
        srand(seed)
    # silence the engine's printf chatter
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        out = mod.MCMainFunction(*args)
    finally:
        ctypes.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return out
