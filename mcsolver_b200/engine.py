"""Python host layer over the C ABI: resident systems (`System`) and the legacy one-shot calls.

`run_on_args` / `run_ising_args` take the reference's positional MCMainFunction tuples
(heisenbergLib.c:504-512, isingLib.c:277-284) and return tuples of the reference's layout
(heisenbergLib.c:853-884: 29 items; isingLib.c:435-447: 11 items).
"""
import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import McgError, check, f64, i32, ptr  # noqa: F401

METROPOLIS, WOLFF = 0, 1
ISING, XY, HEISENBERG = 1, 2, 3


def default_precision():
    """fp64 unless MCSOLVER_B200_PRECISION=32 (fp32 spin state: twice the throughput)."""
    return int(os.environ.get("MCSOLVER_B200_PRECISION", "64"))


def default_seed():
    return int(os.environ.get("MCSOLVER_B200_SEED", "1"))


def device_count():
    n = C.c_int(0)
    check(_ffi.lib().mcg_device_count(C.byref(n)))
    return n.value


class _TableArrays:
    """Keeps the numpy buffers behind an mcg_tables struct alive."""

    def __init__(self, model, S, nlink, J, nbr, pairs, D=None, tri=None, groups=None, nG=0, maxG=1, rOrb=None,
                 rCl=None, rNbr=None, ignoreOffDiag=0):
        self.model = int(model)
        self.S = f64(S).reshape(-1)
        N = self.S.size
        self.N = N
        self.nlink = i32(nlink).reshape(-1)
        if self.nlink.size != N:
            raise ValueError("nlink has %d entries, expected %d" % (self.nlink.size, N))
        nbr = i32(nbr).reshape(-1)
        if N == 0 or nbr.size % N:
            raise ValueError("linkedOrb size %d is not a multiple of N=%d" % (nbr.size, N))
        self.maxL = nbr.size // N
        self.nbr = nbr
        jw = 1 if self.model == ISING else 9
        self.J = f64(J).reshape(-1)
        if self.J.size != N * self.maxL * jw:
            raise ValueError("linkStrength has %d entries, expected %d" % (self.J.size, N * self.maxL * jw))
        self.D = f64(D).reshape(-1) if D is not None else None
        if self.D is not None and self.D.size != 3 * N:
            raise ValueError("initD has %d entries, expected %d" % (self.D.size, 3 * N))
        self.pairs = i32(pairs).reshape(-1)
        if self.pairs.size == 0 or self.pairs.size % 2:
            raise ValueError("corrOrbPair must hold a positive even number of indices")
        self.tri = i32(tri if tri is not None else []).reshape(-1)
        if self.tri.size % 3:
            raise ValueError("localCircuits size must be a multiple of 3")
        self.groups = i32(groups if groups is not None else []).reshape(-1)
        self.nG, self.maxG = int(nG), int(maxG)
        self.rOrb = i32(rOrb if rOrb is not None else []).reshape(-1)
        self.rCl = i32(rCl if rCl is not None else []).reshape(-1)
        self.rNbr = i32(rNbr if rNbr is not None else []).reshape(-1)
        self.nR = self.rOrb.size
        self.nC = self.rCl.size // self.nR if self.nR else 0
        if self.nR and (self.rCl.size != self.nR * self.nC or self.rNbr.size != self.nR * self.maxL):
            raise ValueError("rOrbCluster needs nR*nC and linkedOrb_rnorm nR*maxNLinking entries (got %d, %d for nR=%d)"
                             % (self.rCl.size, self.rNbr.size, self.nR))
        self.ignoreOffDiag = int(ignoreOffDiag)

    @classmethod
    def from_tables(cls, t):
        return cls(t.model, t.S, t.nlink, t.J, t.nbr, t.pairs, D=t.D if t.model != ISING else None, tri=t.tri,
                   groups=t.groups, nG=t.nG, maxG=t.maxG, rOrb=t.rOrb, rCl=t.rCluster, rNbr=t.rNbr,
                   ignoreOffDiag=t.ignoreOffDiag)

    def struct(self):
        t = _ffi.Tables()
        t.model, t.N, t.maxL = self.model, self.N, self.maxL
        t.S, t.D, t.nlink, t.J, t.nbr = ptr(self.S), ptr(self.D) if self.D is not None else None, ptr(self.nlink), ptr(self.J), ptr(self.nbr)
        t.nTri, t.tri = self.tri.size // 3, ptr(self.tri)
        t.nLat, t.pairs = self.pairs.size // 2, ptr(self.pairs)
        t.nG, t.maxG, t.groups = self.nG, self.maxG, ptr(self.groups)
        t.nR, t.nC, t.rOrb, t.rCl, t.rNbr = self.nR, self.nC, ptr(self.rOrb), ptr(self.rCl), ptr(self.rNbr)
        t.ignoreOffDiag = self.ignoreOffDiag
        return t


def _config(precision, nReplica, beta, field, seed, replica_offset, device, keep):
    cfg = _ffi.Config()
    cfg.precision, cfg.nReplica = int(precision), int(nReplica)
    b = f64(beta if beta is not None else np.ones(nReplica)).reshape(-1)
    h = f64(field if field is not None else np.zeros(nReplica)).reshape(-1)
    if b.size != nReplica or h.size != nReplica:
        raise ValueError("beta/field need one entry per replica")
    keep.extend([b, h])
    cfg.beta, cfg.field = ptr(b), ptr(h)
    cfg.seed, cfg.replica_offset, cfg.device = int(seed), int(replica_offset), int(device)
    return cfg


class System:
    """A simulation resident on one GPU: lattice tables + nReplica spin configurations."""

    def __init__(self, handle, model, N, nReplica, nG=0):
        self._h = handle
        self.model, self.N, self.R, self.nG = model, N, nReplica, nG

    @classmethod
    def from_tables(cls, tables, precision=64, nReplica=1, beta=None, field=None, seed=1, replica_offset=0, device=-1):
        """tables: mcsolver_b200.lattice.Tables or _TableArrays."""
        ta = tables if isinstance(tables, _TableArrays) else _TableArrays.from_tables(tables)
        keep = [ta]
        cfg = _config(precision, nReplica, beta, field, seed, replica_offset, device, keep)
        h = C.c_void_p()
        st = ta.struct()
        check(_ffi.lib().mcg_create_tables(C.byref(st), C.byref(cfg), C.byref(h)))
        return cls(h, ta.model, ta.N, nReplica, ta.nG)

    @staticmethod
    def _desc(spec, model, keep, block_spin=False):
        d = _ffi.LatticeDesc()
        d.block_spin = int(bool(block_spin))
        d.model = int(model)
        d.L = (C.c_int32 * 3)(*spec.L)
        d.norb = spec.norb
        S, D = f64(spec.S), f64(spec.D).reshape(-1)
        bonds = (_ffi.Bond * max(1, len(spec.bonds)))()
        for k, (src, tgt, dl, J9) in enumerate(spec.bonds):
            bonds[k].src, bonds[k].tgt = src, tgt
            bonds[k].d = (C.c_int32 * 3)(*dl)
            bonds[k].J = (C.c_double * 9)(*J9)
        circ = i32([[v[0], v[1][0], v[1][1], v[1][2]] for c in spec.circuits for v in c]).reshape(-1)
        keep.extend([S, D, bonds, circ])
        d.S, d.D, d.nbond, d.bonds = ptr(S), ptr(D), len(spec.bonds), C.cast(bonds, C.c_void_p)
        d.pair_s, d.pair_t = spec.pair[0], spec.pair[1]
        d.pair_d = (C.c_int32 * 3)(*spec.pair[2])
        d.ncircuit, d.circuits = len(spec.circuits), ptr(circ)
        mask = np.zeros((len(spec.groups), spec.norb), dtype=np.int32)
        for k, grp in enumerate(spec.groups):
            mask[k, list(grp)] = 1
        keep.append(mask)
        d.ngroup, d.group_mask, d.group_in_sc = (len(spec.groups) if model != ISING else 0), ptr(mask), int(bool(spec.groupInSC))
        return d

    @classmethod
    def from_spec(cls, spec, model, precision=32, nReplica=1, beta=None, field=None, seed=1, replica_offset=0, device=-1,
                  block_spin=False):
        """Structured path: spec is a mcsolver_b200.lattice.LatticeSpec (bond templates + supercell).
        block_spin=True also evaluates the block-spin statistics (tuple slots 11-19; Ising 6-7) every measured sweep."""
        keep = []
        d = cls._desc(spec, model, keep, block_spin)
        cfg = _config(precision, nReplica, beta, field, seed, replica_offset, device, keep)
        h = C.c_void_p()
        check(_ffi.lib().mcg_create_lattice(C.byref(d), C.byref(cfg), C.byref(h)))
        return cls(h, int(model), spec.nsite, nReplica, len(spec.groups) if model != ISING else 0)

    @classmethod
    def from_spec_slab(cls, spec, model, rank, world, comm_id=None, precision=32, nReplica=1, beta=None, field=None, seed=1,
                       replica_offset=0, device=-1):
        """This rank's slab of ONE lattice cut along its first axis over `world` ranks (mcg_create_lattice_slab): spec is the
        whole lattice, comm_id the bytes of mcsolver_b200.pt.comm_id(rank, world) (None for world == 1).  Sweeps, measurements
        and energy() are collective; results() are the whole lattice's on every rank.  self.N counts the local planes, ghost
        planes included (slab_info())."""
        keep = []
        d = cls._desc(spec, model, keep, False)
        cfg = _config(precision, nReplica, beta, field, seed, replica_offset, device, keep)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(comm_id, len(comm_id)) if comm_id else None
        check(_ffi.lib().mcg_create_lattice_slab(C.byref(d), C.byref(cfg), int(rank), int(world), idbuf, C.byref(h)))
        info = (C.c_int32 * 6)()
        check(_ffi.lib().mcg_slab_info(h, info))
        nloc = (info[3] + 2 * info[4]) * spec.L[1] * spec.L[2] * spec.norb
        obj = cls(h, int(model), nloc, nReplica, len(spec.groups) if model != ISING else 0)
        obj.slab = dict(rank=info[0], world=info[1], x0=info[2], nx=info[3], ghost=info[4], Lx=info[5], Ly=spec.L[1], Lz=spec.L[2], norb=spec.norb)
        return obj

    def slab_sync(self):
        """collective: refresh the ghost planes from the neighbouring slabs (after set_spins)"""
        check(_ffi.lib().mcg_slab_sync(self._h))

    def own_spins(self, replica=0):
        """a slab's own planes (ghosts stripped) in reference order: rows x0 .. x0+nx of the whole lattice's get_spins()"""
        sl = self.slab
        per_x = sl["Ly"] * sl["Lz"] * sl["norb"]
        sp = self.get_spins(replica)
        return sp[sl["ghost"] * per_x:(sl["ghost"] + sl["nx"]) * per_x]

    def close(self):
        if self._h is not None:
            _ffi.lib().mcg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def _shape(self):
        return (self.N,) if self.model == ISING else (self.N, 3)

    def num_colours(self):
        n = C.c_int(0)
        check(_ffi.lib().mcg_num_colours(self._h, C.byref(n)))
        return n.value

    def colour_order(self):
        o = np.zeros(self.N, dtype=np.int32)
        check(_ffi.lib().mcg_colour_order(self._h, ptr(o)))
        return o

    def rng_layout(self):
        """(stride, group) of the Philox streams of the Metropolis sweeps (csrc/rng.cuh); (0, 0) = one block per site."""
        a, b = C.c_int32(0), C.c_int32(0)
        check(_ffi.lib().mcg_rng_layout(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_params(self, beta=None, field=None):
        b = f64(beta) if beta is not None else None
        h = f64(field) if field is not None else None
        check(_ffi.lib().mcg_set_params(self._h, ptr(b) if b is not None else None, ptr(h) if h is not None else None))

    def recycle(self, beta=None, field=None, seed=1, replica_offset=0):
        """Back to the state of a freshly created system with these replica parameters and Philox streams (mcg_recycle):
        after init_spins() a run reproduces a fresh system's bit for bit, without paying for the creation."""
        b = f64(beta).reshape(-1) if beta is not None else None
        h = f64(field).reshape(-1) if field is not None else None
        if (b is not None and b.size != self.R) or (h is not None and h.size != self.R):
            raise ValueError("beta/field need one entry per replica")
        check(_ffi.lib().mcg_recycle(self._h, ptr(b) if b is not None else None, ptr(h) if h is not None else None, int(seed), int(replica_offset)))

    def init_spins(self, flunc=0.0):
        check(_ffi.lib().mcg_init_spins(self._h, float(flunc)))

    def set_spins(self, spins, replica=0):
        sp = f64(spins).reshape(self._shape)
        check(_ffi.lib().mcg_set_spins(self._h, int(replica), ptr(sp)))

    def get_spins(self, replica=0):
        sp = np.zeros(self._shape)
        check(_ffi.lib().mcg_get_spins(self._h, int(replica), ptr(sp)))
        return sp

    def energy(self, replica=0, per_site=False):
        E = C.c_double(0)
        if per_site:
            eb, eo = np.zeros(self.N), np.zeros(self.N)
            check(_ffi.lib().mcg_energy(self._h, int(replica), C.byref(E), ptr(eb), ptr(eo)))
            return E.value, eb, eo
        check(_ffi.lib().mcg_energy(self._h, int(replica), C.byref(E), None, None))
        return E.value

    def metropolis_sweeps(self, n, p_attempt=1.0):
        check(_ffi.lib().mcg_metropolis_sweeps(self._h, int(n), float(p_attempt)))

    def timed_sweeps(self, n, p_attempt=1.0, with_measure=False):
        """Device time (ms, CUDA events on the launch stream) of n Metropolis sweeps."""
        ms = C.c_double(0)
        check(_ffi.lib().mcg_timed_sweeps(self._h, int(n), float(p_attempt), int(bool(with_measure)), C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64(0)
        check(_ffi.lib().mcg_launch_count(self._h, C.byref(n)))
        return n.value

    def jit_launch_count(self):
        """Launches of NVRTC-specialised kernels so far (0 = the offline runtime-table kernels ran)."""
        n = C.c_int64(0)
        check(_ffi.lib().mcg_jit_launch_count(self._h, C.byref(n)))
        return n.value

    def jit_module_key(self, colour=0):
        """'%016x' cache key of the colour's NVRTC-specialised module (name of its cubin file)."""
        k = C.c_uint64(0)
        check(_ffi.lib().mcg_jit_module_key(self._h, int(colour), C.byref(k)))
        return "%016x" % k.value

    def profile_passes(self, on=True):
        check(_ffi.lib().mcg_profile_passes(self._h, int(bool(on))))

    def profile_read(self):
        """(total device ms, launches) of the colour-pass kernel since the last read."""
        ms, n = C.c_double(0), C.c_int64(0)
        check(_ffi.lib().mcg_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def wolff_steps(self, n):
        check(_ffi.lib().mcg_wolff_steps(self._h, int(n)))

    def measure(self):
        check(_ffi.lib().mcg_measure(self._h))

    def reset_measurements(self):
        check(_ffi.lib().mcg_reset_measurements(self._h))

    def results(self, replica=0):
        out = np.zeros(10 if self.model == ISING else 27)
        g = np.zeros((self.nG + 2) * (self.nG + 1))
        check(_ffi.lib().mcg_results(self._h, int(replica), ptr(out), ptr(g)))
        return (out, g) if self.model != ISING else (out, None)

    def counters(self, replica=0):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(_ffi.lib().mcg_counters(self._h, int(replica), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def wolff_frontier_steps(self, replica=0):
        """Wolff steps of this replica completed by frontier growth from the seed (the others ran the global bond passes)"""
        a = C.c_int64(0)
        check(_ffi.lib().mcg_wolff_frontier_steps(self._h, int(replica), C.byref(a)))
        return a.value

    def run(self, algorithm, nthermal, nsweep, ninterval, spinFrame=0):
        fr = None
        if spinFrame > 0:
            fr = np.zeros((self.R, spinFrame) + self._shape)
        check(_ffi.lib().mcg_run(self._h, int(algorithm), int(nthermal), int(nsweep), int(ninterval), int(spinFrame),
                                 ptr(fr) if fr is not None else None))
        return fr


# ---------------------------------------------------------------------------------------------
# legacy one-shot calls: positional tuples in, reference-layout tuples out
# ---------------------------------------------------------------------------------------------
def slab_plan(spec, model, rank, world, precision=32):
    """How mcg_create_lattice_slab would cut the lattice for rank `rank` of `world` (host only, no GPU): dict(x0, nx, ghost, Lx);
    raises McgError when the lattice cannot be cut that way."""
    keep = []
    d = System._desc(spec, model, keep, False)
    info = (C.c_int32 * 6)()
    check(_ffi.lib().mcg_slab_plan(C.byref(d), int(precision), int(rank), int(world), info))
    return dict(rank=info[0], world=info[1], x0=info[2], nx=info[3], ghost=info[4], Lx=info[5])


def jit_check(spec, model, precision=32, block_spin=False):
    """Host-only: build the class tables of a lattice (incl. the block-spin tables when asked) and NVRTC-compile its
    specialised pass kernels; returns (n_modules, report)."""
    keep = []
    d = System._desc(spec, model, keep, block_spin)
    n = C.c_int(0)
    buf = C.create_string_buffer(8192)
    check(_ffi.lib().mcg_jit_check(C.byref(d), int(precision), C.byref(n), buf, len(buf)))
    return n.value, buf.value.decode()


def _as_int(x, name):
    if isinstance(x, bool) or not isinstance(x, (int, np.integer)):
        raise TypeError("%s must be an int, got %r" % (name, type(x).__name__))
    return int(x)


def run_on_args(model, args, seed=None, precision=None):
    """MCMainFunction of xylib (model=2) / heisenberglib (model=3): 23 positional args -> 29-tuple."""
    if len(args) != 23:
        raise TypeError("MCMainFunction() takes exactly 23 positional arguments (%d given)" % len(args))
    (algorithm, initSpin, initD, nthermal, nsweep, ninterval, maxNLinking, nlink, linkStrength, linkedOrb,
     localCircuits, corrOrbPair, nOrbGroup, maxOrbGroupSize, orbGroupList, flunc, h, rOrb, rOrbCluster,
     linkedOrb_rnorm, spinFrame, ignoreNonDiagonalJ, _callback) = args
    algorithm = _as_int(algorithm, "algorithm")
    nthermal, nsweep, ninterval = _as_int(nthermal, "nthermal"), _as_int(nsweep, "nsweep"), _as_int(ninterval, "ninterval")
    spinFrame = _as_int(spinFrame, "spinFrame")
    ta = _TableArrays(model, initSpin, nlink, linkStrength, linkedOrb, corrOrbPair, D=initD, tri=localCircuits,
                      groups=orbGroupList, nG=_as_int(nOrbGroup, "nOrbGroup"), maxG=_as_int(maxOrbGroupSize, "maxOrbGroupSize"),
                      rOrb=rOrb, rCl=rOrbCluster, rNbr=linkedOrb_rnorm, ignoreOffDiag=_as_int(ignoreNonDiagonalJ, "ignoreNonDiagonalJ"))
    if ta.maxL != _as_int(maxNLinking, "maxNLinking"):
        raise ValueError("maxNLinking=%d does not match the table sizes (%d)" % (maxNLinking, ta.maxL))
    out = np.zeros(27)
    frames = np.zeros((max(spinFrame, 0), ta.N, 3))
    group = np.zeros((ta.nG + 2) * (ta.nG + 1))
    st = ta.struct()
    check(_ffi.lib().mcg_run_on(C.byref(st), algorithm, nthermal, nsweep, ninterval, float(flunc), float(h), spinFrame,
                                default_seed() if seed is None else int(seed),
                                default_precision() if precision is None else int(precision),
                                ptr(out), ptr(frames), ptr(group)))
    # slot 27: tuple[spinFrame] of tuple[N] of (x,y,z) or 0.0; slot 28: group tuple or 0.0 (heisenbergLib.c:647-653, 849-851)
    fr = tuple(tuple(map(tuple, f.tolist())) for f in frames) if spinFrame > 0 else 0.0
    gr = tuple(group.tolist()) if ta.nG > 0 else 0.0
    return tuple(out.tolist()) + (fr, gr)


def run_ising_args(args, seed=None, precision=None):
    """MCMainFunction of isinglib: 16 positional args -> 11-tuple."""
    if len(args) != 16:
        raise TypeError("MCMainFunction() takes exactly 16 positional arguments (%d given)" % len(args))
    (algorithm, initSpin, nthermal, nsweep, ninterval, maxNLinking, nlink, linkStrength, linkedOrb, corrOrbPair, h,
     rOrb, rOrbCluster, linkedOrb_rnorm, spinFrame, _callback) = args
    algorithm = _as_int(algorithm, "algorithm")
    nthermal, nsweep, ninterval = _as_int(nthermal, "nthermal"), _as_int(nsweep, "nsweep"), _as_int(ninterval, "ninterval")
    spinFrame = _as_int(spinFrame, "spinFrame")
    ta = _TableArrays(ISING, initSpin, nlink, linkStrength, linkedOrb, corrOrbPair, rOrb=rOrb, rCl=rOrbCluster,
                      rNbr=linkedOrb_rnorm)
    if ta.maxL != _as_int(maxNLinking, "maxNLinking"):
        raise ValueError("maxNLinking=%d does not match the table sizes (%d)" % (maxNLinking, ta.maxL))
    out = np.zeros(10)
    frames = np.zeros((max(spinFrame, 0), ta.N))
    st = ta.struct()
    check(_ffi.lib().mcg_run_ising(C.byref(st), algorithm, nthermal, nsweep, ninterval, float(h), spinFrame,
                                   default_seed() if seed is None else int(seed),
                                   default_precision() if precision is None else int(precision), ptr(out), ptr(frames)))
    fr = tuple(tuple(f.tolist()) for f in frames) if spinFrame > 0 else 0.0
    return tuple(out.tolist()) + (fr,)
