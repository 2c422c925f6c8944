"""TEST INFRASTRUCTURE ONLY - ctypes front end of oracle/liboracle.so (oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  The product (mcsolver_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), so])
    return so


class _Sys(C.Structure):
    _fields_ = [("model", C.c_int), ("N", C.c_int), ("maxL", C.c_int),
                ("S", C.c_void_p), ("D", C.c_void_p), ("nlink", C.c_void_p), ("J", C.c_void_p),
                ("nbr", C.c_void_p), ("nTri", C.c_int), ("tri", C.c_void_p), ("nLat", C.c_int),
                ("pairs", C.c_void_p), ("nG", C.c_int), ("maxG", C.c_int), ("groups", C.c_void_p),
                ("nR", C.c_int), ("nC", C.c_int), ("rOrb", C.c_void_p), ("rCl", C.c_void_p),
                ("rNbr", C.c_void_p), ("h", C.c_double), ("ignoreOffDiag", C.c_int),
                ("isingStrideBug", C.c_int), ("wolffHalfMove", C.c_int), ("rngStride", C.c_int), ("rngGroup", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        dp, ip = C.c_void_p, C.c_void_p
        L.orc_total_energy.restype = C.c_double
        L.orc_total_energy.argtypes = [C.POINTER(_Sys), dp]
        L.orc_site_energies.argtypes = [C.POINTER(_Sys), dp, dp, dp]
        L.orc_delta_energy.restype = C.c_double
        L.orc_delta_energy.argtypes = [C.POINTER(_Sys), dp, C.c_int, dp]
        L.orc_ising_flip_corr.restype = C.c_double
        L.orc_ising_flip_corr.argtypes = [C.POINTER(_Sys), dp, C.c_int]
        L.orc_signed_area.restype = C.c_double
        L.orc_signed_area.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_double]
        L.orc_topological_q.restype = C.c_double
        L.orc_topological_q.argtypes = [C.POINTER(_Sys), dp]
        L.orc_philox4x32.argtypes = [dp, dp, dp]
        L.orc_observe_on.argtypes = [C.POINTER(_Sys), dp, dp, dp]
        L.orc_init_spins_philox.argtypes = [C.POINTER(_Sys), C.c_double, C.c_uint64, C.c_uint32, C.c_int, dp]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(_Sys), C.c_int, C.c_long, C.c_long, C.c_long, C.c_double, C.c_int, ip,
                              C.c_uint64, C.c_uint32, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.orc_run_ising.restype = C.c_int
        L.orc_run_ising.argtypes = [C.POINTER(_Sys), C.c_int, C.c_long, C.c_long, C.c_long, C.c_int, ip,
                                    C.c_uint64, C.c_uint32, C.c_int, dp, dp, dp, dp]
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class System:
    """The flat tables of one MCMainFunction call (model 1/2/3), held as numpy arrays."""

    def __init__(self, model, S, nlink, J, nbr, pairs, D=None, tri=None, groups=None, rOrb=None, rCl=None,
                 rNbr=None, h=0.0, ignoreOffDiag=0, isingStrideBug=0, wolffHalfMove=0):
        self.model = int(model)
        self.S = _f64(S).reshape(-1)
        self.N = self.S.size
        self.nlink = _i32(nlink).reshape(-1)
        self.nbr = _i32(nbr).reshape(self.N, -1) if self.N else _i32(nbr).reshape(0, 0)
        self.maxL = self.nbr.shape[1]
        self.J = _f64(J).reshape(self.N, self.maxL) if model == 1 else _f64(J).reshape(self.N, self.maxL, 9)
        self.D = _f64(D if D is not None else np.zeros((self.N, 3))).reshape(self.N, 3)
        self.pairs = _i32(pairs).reshape(-1, 2)
        self.tri = _i32(tri if tri is not None else np.zeros((0, 3))).reshape(-1, 3)
        g = _i32(groups) if groups is not None else None
        self.groups = g.reshape(g.shape[0], -1) if (g is not None and g.ndim == 2 and g.shape[0]) else _i32(np.zeros((0, 1)))
        self.nG = self.groups.shape[0]
        self.maxG = self.groups.shape[1] if self.nG else 1
        self.rOrb = _i32(rOrb if rOrb is not None else []).reshape(-1)
        self.nR = self.rOrb.size
        self.rCl = _i32(rCl if rCl is not None else []).reshape(self.nR, -1) if self.nR else _i32([]).reshape(0, 0)
        self.nC = self.rCl.shape[1] if self.nR else 0
        self.rNbr = _i32(rNbr if rNbr is not None else []).reshape(self.nR, -1) if self.nR else _i32([]).reshape(0, 0)
        self.h = float(h)
        self.ignoreOffDiag = int(ignoreOffDiag)
        self.isingStrideBug = int(isingStrideBug)
        self.wolffHalfMove = int(wolffHalfMove)
        self.rng_layout = (0, 0)   # (stride, group) of the Philox colour sweeps: engine.System.rng_layout() of the path under test

    @classmethod
    def from_tables(cls, t, h_over_T=0.0, **kw):
        return cls(t.model, t.S, t.nlink, t.J, t.nbr, t.pairs, D=t.D, tri=t.tri,
                   groups=t.groups if t.nG else None, rOrb=t.rOrb, rCl=t.rCluster, rNbr=t.rNbr, h=h_over_T,
                   ignoreOffDiag=t.ignoreOffDiag if t.model != 1 else 0, **kw)

    @classmethod
    def from_on_args(cls, model, args, **kw):
        """From the 23 positional arguments of the O(n) MCMainFunction."""
        (algo, S, D, nth, nsw, nint, maxL, nlink, J, nbr, tri, pairs, nG, maxG, groups, flunc, h, rOrb, rCl,
         rNbr, spinFrame, ign, cb) = args
        N = len(S)
        nR = len(rOrb)
        g = np.array(groups, dtype=np.int32).reshape(nG, maxG) if nG else None
        return cls(model, S, nlink, np.array(J).reshape(N, maxL, 9), np.array(nbr).reshape(N, maxL), pairs, D=D,
                   tri=tri, groups=g, rOrb=rOrb, rCl=np.array(rCl).reshape(nR, -1), rNbr=np.array(rNbr).reshape(nR, -1),
                   h=h, ignoreOffDiag=ign, **kw)

    @classmethod
    def from_ising_args(cls, args, **kw):
        (algo, S, nth, nsw, nint, maxL, nlink, J, nbr, pairs, h, rOrb, rCl, rNbr, spinFrame, cb) = args
        N = len(S)
        nR = len(rOrb)
        return cls(1, S, nlink, np.array(J).reshape(N, maxL), np.array(nbr).reshape(N, maxL), pairs, rOrb=rOrb,
                   rCl=np.array(rCl).reshape(nR, -1), rNbr=np.array(rNbr).reshape(nR, -1), h=h, **kw)

    def c(self):
        s = _Sys()
        s.model, s.N, s.maxL = self.model, self.N, self.maxL
        s.S, s.D, s.nlink, s.J, s.nbr = _p(self.S), _p(self.D), _p(self.nlink), _p(self.J), _p(self.nbr)
        s.nTri, s.tri = self.tri.shape[0], _p(self.tri)
        s.nLat, s.pairs = self.pairs.shape[0], _p(self.pairs)
        s.nG, s.maxG, s.groups = self.nG, self.maxG, _p(self.groups)
        s.nR, s.nC, s.rOrb, s.rCl, s.rNbr = self.nR, self.nC, _p(self.rOrb), _p(self.rCl), _p(self.rNbr)
        s.h, s.ignoreOffDiag = self.h, self.ignoreOffDiag
        s.isingStrideBug, s.wolffHalfMove = self.isingStrideBug, self.wolffHalfMove
        s.rngStride, s.rngGroup = int(self.rng_layout[0]), int(self.rng_layout[1])
        return s

    # -- configuration-level functions ------------------------------------------------------
    def _sp(self, spins):
        sp = _f64(spins)
        return sp.reshape(self.N) if self.model == 1 else sp.reshape(self.N, 3)

    def total_energy(self, spins):
        sp = self._sp(spins)
        return lib().orc_total_energy(C.byref(self.c()), _p(sp))

    def site_energies(self, spins):
        sp = self._sp(spins)
        eb, eo = np.zeros(self.N), np.zeros(self.N)
        lib().orc_site_energies(C.byref(self.c()), _p(sp), _p(eb), _p(eo))
        return eb, eo

    def delta_energy(self, spins, i, trans):
        sp = self._sp(spins)
        tr = _f64(trans)
        return lib().orc_delta_energy(C.byref(self.c()), _p(sp), int(i), _p(tr))

    def ising_flip_corr(self, spins, i):
        return lib().orc_ising_flip_corr(C.byref(self.c()), _p(self._sp(spins)), int(i))

    def topological_q(self, spins):
        return lib().orc_topological_q(C.byref(self.c()), _p(self._sp(spins)))

    def observe(self, spins):
        """Reference-definition observables of one configuration: (out27, groupOut)."""
        sp = self._sp(spins)
        out = np.zeros(27)
        g = np.zeros((self.nG + 2) * (self.nG + 1))
        lib().orc_observe_on(C.byref(self.c()), _p(sp), _p(out), _p(g))
        return out, g

    def init_spins_philox(self, flunc, seed, replica=0, f32=False):
        sp = np.zeros((self.N, 3))
        lib().orc_init_spins_philox(C.byref(self.c()), float(flunc), int(seed), int(replica), int(f32), _p(sp))
        return sp

    # -- whole runs ---------------------------------------------------------------------------
    def run(self, mode, nthermal, nsweep, ninterval, flunc=0.0, spinFrame=0, order=None, seed=1, replica=0,
            f32=False, spins=None):
        """orc_run / orc_run_ising.  mode: 0 ref Metropolis (rand), 1 ref Wolff (rand),
        2 Philox colour sweeps, 3 Philox Wolff.  Returns dict(out, frames, group, spins, counters)."""
        cnt = np.zeros(3, dtype=np.int64)
        ordp = _i32(order) if order is not None else None
        if self.model == 1:
            sp = _f64(spins if spins is not None else self.S).reshape(self.N).copy()
            out = np.zeros(10)
            fr = np.zeros((max(spinFrame, 0), self.N))
            lib().orc_run_ising(C.byref(self.c()), mode, nthermal, nsweep, ninterval, spinFrame, _p(ordp) if ordp is not None else None,
                                int(seed), int(replica), int(f32), _p(sp), _p(out), _p(fr), _p(cnt))
            return dict(out=out, frames=fr, group=None, spins=sp, counters=cnt)
        given = spins is not None
        sp = _f64(spins).reshape(self.N, 3).copy() if given else np.zeros((self.N, 3))
        out = np.zeros(27)
        fr = np.zeros((max(spinFrame, 0), self.N, 3))
        g = np.zeros((self.nG + 2) * (self.nG + 1))
        lib().orc_run(C.byref(self.c()), mode, nthermal, nsweep, ninterval, float(flunc), spinFrame,
                      _p(ordp) if ordp is not None else None, int(seed), int(replica), int(f32), int(given),
                      _p(sp), _p(out), _p(fr), _p(g), _p(cnt))
        return dict(out=out, frames=fr, group=g, spins=sp, counters=cnt)


def philox4x32(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(_p(c), _p(k), _p(o))
    return o


def signed_area(s1, s2, s3, l1=1.0, l2=1.0, l3=1.0):
    a, b, c = _f64(s1), _f64(s2), _f64(s3)
    return lib().orc_signed_area(_p(a), _p(b), _p(c), l1, l2, l3)
