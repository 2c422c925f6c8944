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
_NUM = r"[0-9\.\-]+"


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


def parse(path) -> Params:
    with open(path, "r") as f:
        data = [line for line in f.read().split("\n") if line]
    version = re.findall(r"[0-9\.]+", data[0])[0]
    if version != VERSION:
        raise ValueError("unknown file or version (only support v%s)" % VERSION)
    tags = {}
    for i, line in enumerate(data):
        kw = re.findall(r"[a-zA-Z]+", line)
        if kw:
            tags[kw[0]] = i          # later lines overwrite earlier ones, as in the reference
    need = ["Lattice", "Supercell", "Orbitals", "Bonds", "Temperature", "Sweeps", "Model", "Algorithm", "Ncores", "Field", "Dipole",
            "Distribution", "XAxis", "OrbGroup", "LocalCircuit", "Measurement"]
    missing = [k for k in need if not tags.get(k)]
    if missing:
        raise ValueError("cannot find some tags: %s" % ", ".join(missing))
    t = tags
    LMatrix = [[float(x) for x in re.findall(_NUM, data[t["Lattice"] + 1 + i])] for i in range(3)]
    LPack = [int(x) for x in re.findall(r"[0-9]+", data[t["Supercell"] + 1])]
    norb = int(re.findall(r"[0-9]+", data[t["Orbitals"] + 1])[0])
    pos, S, DList = [], [], []
    for i in range(norb):
        e = re.findall(_NUM, data[t["Orbitals"] + 3 + i])
        S.append(float(e[2]))
        pos.append([float(e[3]), float(e[4]), float(e[5])])
        DList.append([float(e[6]), float(e[7]), float(e[8])])
    nb = int(re.findall(r"[0-9]+", data[t["Bonds"] + 1])[0])
    bondList = []
    for i in range(nb):
        e = re.findall(_NUM, data[t["Bonds"] + 3 + i])
        bondList.append([int(e[10]), int(e[11]), [int(e[12]), int(e[13]), int(e[14])]] + [float(v) for v in e[1:10]])
    s_, t_, v1, v2, v3 = [int(x) for x in re.findall(r"[0-9\-]+", data[t["Measurement"] + 1])]
    nG = int(re.findall(r"[0-9]+", data[t["OrbGroup"]])[0])
    groupInSC = data[t["OrbGroup"] + 1] == "Supergroup"
    groups = []
    for i in range(nG):
        _, a, b = re.findall(r"[0-9]+", data[t["OrbGroup"] + i + 2])
        groups.append(list(range(int(a), int(b) + 1)))
    nC = int(re.findall(r"[0-9]+", data[t["LocalCircuit"]])[0])
    circuits = []
    for i in range(nC):
        e = [int(x) for x in re.findall(r"[0-9\-]+", data[t["LocalCircuit"] + 1 + i])]
        circuits.append([(e[1], (e[2], e[3], e[4])), (e[5], (e[6], e[7], e[8])), (e[9], (e[10], e[11], e[12]))])
    Tp = re.findall(r"[0-9\.]+", data[t["Temperature"] + 1])
    Hp = re.findall(_NUM, data[t["Field"] + 1])
    sw = [int(x) for x in re.findall(_NUM, data[t["Sweeps"] + 1])]
    return Params(LMatrix=LMatrix, LPack=LPack, pos=pos, S=S, DList=DList, bondList=bondList, T0=float(Tp[0]), T1=float(Tp[1]),
                  nT=int(Tp[2]), H0=float(Hp[0]), H1=float(Hp[1]), nH=int(Hp[2]),
                  dipoleAlpha=float(re.findall(_NUM, data[t["Dipole"] + 1])[0]), nthermal=sw[0], nsweep=sw[1], ninterval=sw[2],
                  xAxisType=data[t["XAxis"] + 1], modelType=data[t["Model"] + 1], algorithm=data[t["Algorithm"] + 1],
                  GcOrb=[[s_, t_], [v1, v2, v3]], ncores=int(data[t["Ncores"] + 1]),
                  spinFrame=int(re.findall(r"[0-9]+", data[t["Distribution"]])[0]), orbGroupList=groups, groupInSC=groupInSC,
                  localCircuitList=circuits)
