"""Reader of the reference's parameter ("save") file, version 3.0.

Mirrors the grammar that `fileio.py:169-292` actually enforces: sections are
located by the FIRST ALPHABETIC WORD of a line, values are taken POSITIONALLY from the numbers on the
following line(s) - labels such as "Dz"/"Jx" are ignored, exactly as in the reference (so
`samples/CrI3With2NNCoupling`'s "Dz -3.12 Dx 0 Dy 0" lands in D[0]).  The reference's own parser
cannot be imported headless (it needs tkinter); this one has no GUI dependency.
"""
import re
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .lattice import LatticeSpec

VERSION = "3.0"


@dataclass
class Params:
    LMatrix: List[List[float]]
    LPack: List[int]
    pos: List[List[float]]
    S: List[float]
    DList: List[List[float]]
    bondList: list                # [src, tgt, [n1,n2,n3], J0..J8]
    T0: float
    T1: float
    nT: int
    H0: float
    H1: float
    nH: int
    dipoleAlpha: float
    nthermal: int
    nsweep: int
    ninterval: int
    xAxisType: str
    modelType: str
    algorithm: str
    GcOrb: list                   # [[s,t],[v1,v2,v3]]
    ncores: int
    spinFrame: int
    orbGroupList: list
    groupInSC: bool
    localCircuitList: list

    @property
    def model(self):
        return {"Ising": 1, "XY": 2, "Heisenberg": 3}[self.modelType]

    def spec(self) -> LatticeSpec:
        bonds = [(b[0], b[1], tuple(b[2]), list(b[3:12])) for b in self.bondList]
        if self.model == 1:      # win.py:58-63: Bond(..., On=False) keeps only Jxx
            bonds = [(b[0], b[1], b[2], [b[3][0]] + [0.0] * 8) for b in bonds]
        return LatticeSpec(L=tuple(self.LPack), S=self.S, D=self.DList, bonds=bonds, LMatrix=self.LMatrix, pos=self.pos,
                           pair=(self.GcOrb[0][0], self.GcOrb[0][1], tuple(self.GcOrb[1])), groups=self.orbGroupList,
                           groupInSC=self.groupInSC, circuits=[tuple((o, tuple(d)) for o, d in c) for c in self.localCircuitList])

    def grid(self):
        """(T, H) per task in the reference's order: H outer, T inner (win.py:75-78, 116-119)."""
        TList = np.linspace(self.T0, self.T1, self.nT)
        HList = np.linspace(self.H0, self.H1, self.nH)
        T = np.tile(TList, len(HList))
        H = np.repeat(HList, len(TList))
        return T, H


# ---- the grammar as data -------------------------------------------------------------------------------------------------
# A section is the LAST line whose first alphabetic word is the key; its values sit on the lines after it (or on the header line
# itself).  Fixed-shape sections are described here; the three repeated blocks (orbitals, bonds, circuits) and the group list have
# their own readers below.
_FLOATS, _INTS, _UNSIGNED, _WORD = r"[0-9\.\-]+", r"[0-9\-]+", r"[0-9]+", None
_SCALARS = {
    # attribute(s)                     section        line offset   token pattern   converters
    ("T0", "T1", "nT"):               ("Temperature",  1,           r"[0-9\.]+",    (float, float, int)),
    ("H0", "H1", "nH"):               ("Field",        1,           _FLOATS,        (float, float, int)),
    ("dipoleAlpha",):                 ("Dipole",       1,           _FLOATS,        (float,)),
    ("nthermal", "nsweep", "ninterval"): ("Sweeps",    1,           _FLOATS,        (int, int, int)),
    ("spinFrame",):                   ("Distribution", 0,           _UNSIGNED,      (int,)),
    ("xAxisType",):                   ("XAxis",        1,           _WORD,          (str,)),
    ("modelType",):                   ("Model",        1,           _WORD,          (str,)),
    ("algorithm",):                   ("Algorithm",    1,           _WORD,          (str,)),
    ("ncores",):                      ("Ncores",       1,           _WORD,          (int,)),
}
_SECTIONS = ("Lattice", "Supercell", "Orbitals", "Bonds", "Measurement", "OrbGroup", "LocalCircuit") + tuple(v[0] for v in _SCALARS.values())


class _Doc:
    """The non-empty lines of a save file, addressed as (section key, offset)."""

    def __init__(self, text):
        self.lines = [ln for ln in text.split("\n") if ln]
        found = re.findall(r"[0-9\.]+", self.lines[0]) if self.lines else []
        if not found or found[0] != VERSION:
            raise ValueError("unknown file or version (only support v%s)" % VERSION)
        self.at = {}
        for n, ln in enumerate(self.lines):
            m = re.search(r"[a-zA-Z]+", ln)
            if m:
                self.at[m.group(0)] = n          # a later header of the same name wins, as in the reference
        absent = [k for k in _SECTIONS if not self.at.get(k)]
        if absent:
            raise ValueError("cannot find some tags: %s" % ", ".join(absent))

    def line(self, key, offset=0):
        return self.lines[self.at[key] + offset]

    def tokens(self, key, offset, pattern):
        return re.findall(pattern, self.line(key, offset))

    def count(self, key, offset):
        return int(self.tokens(key, offset, _UNSIGNED)[0])


def _orbitals(doc):
    """'orb k: type t spin S pos [x y z] Dx a Dy b Dz c ...' - numbers are positional, the labels are decoration"""
    S, pos, D = [], [], []
    for k in range(doc.count("Orbitals", 1)):
        v = [float(x) for x in doc.tokens("Orbitals", 3 + k, _FLOATS)]
        S.append(v[2]); pos.append(v[3:6]); D.append(v[6:9])
    return S, pos, D


def _bonds(doc):
    """'bond k: Jx .. (nine numbers) orb s to orb t over [a b c]' -> [s, t, [a, b, c], J0 .. J8]"""
    out = []
    for k in range(doc.count("Bonds", 1)):
        v = doc.tokens("Bonds", 3 + k, _FLOATS)
        out.append([int(v[10]), int(v[11]), [int(x) for x in v[12:15]]] + [float(x) for x in v[1:10]])
    return out


def _groups(doc):
    """'OrbGroup:n' / Supergroup|... / 'groupk orbA-orbB' -> inclusive orbital ranges"""
    ranges = []
    for k in range(doc.count("OrbGroup", 0)):
        _, lo, hi = doc.tokens("OrbGroup", 2 + k, _UNSIGNED)
        ranges.append(list(range(int(lo), int(hi) + 1)))
    return ranges, doc.line("OrbGroup", 1) == "Supergroup"


def _circuits(doc):
    """'LocalCircuit per cell: n' then n lines 'id o1 a b c o2 a b c o3 a b c'"""
    out = []
    for k in range(doc.count("LocalCircuit", 0)):
        v = [int(x) for x in doc.tokens("LocalCircuit", 1 + k, _INTS)]
        out.append([(v[1 + 4 * j], tuple(v[2 + 4 * j:5 + 4 * j])) for j in range(3)])
    return out


def parse(path) -> Params:
    with open(path, "r") as f:
        doc = _Doc(f.read())
    kw = {}
    for names, (key, offset, pattern, conv) in _SCALARS.items():
        raw = [doc.line(key, offset)] if pattern is _WORD else doc.tokens(key, offset, pattern)
        for name, fn, tok in zip(names, conv, raw):
            kw[name] = fn(tok)
    S, pos, D = _orbitals(doc)
    groups, in_sc = _groups(doc)
    corr = [int(x) for x in doc.tokens("Measurement", 1, _INTS)]      # orbS orbT over [a b c]
    return Params(LMatrix=[[float(x) for x in doc.tokens("Lattice", 1 + k, _FLOATS)] for k in range(3)],
                  LPack=[int(x) for x in doc.tokens("Supercell", 1, _UNSIGNED)], pos=pos, S=S, DList=D, bondList=_bonds(doc),
                  GcOrb=[corr[0:2], corr[2:5]], orbGroupList=groups, groupInSC=in_sc, localCircuitList=_circuits(doc), **kw)
