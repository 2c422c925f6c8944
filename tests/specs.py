"""Lattice models used by the tests and by tests/golden/make_golden.py (the BASELINE configs at
reference-feasible sizes plus one fully anisotropic test model)."""
from mcsolver_b200.lattice import LatticeSpec

HEX = [[1, 0, 0], [-0.5, 0.8660254, 0], [0, 0, 1]]
HEXPOS = [[1 / 3, 2 / 3, 0], [2 / 3, 1 / 3, 0]]
# samples/SkyrmionOnHexLattice (bonds, D, circuits verbatim)
SKYR = dict(S=[1, 1], D=[[0, 0, -0.1]] * 2, LMatrix=HEX, pos=HEXPOS,
            bonds=[(0, 1, (0, 0, 0), [-1, -1, -1, 0, -0.8660254, 0.5, 0, 0.8660254, -0.5]),
                   (0, 1, (0, 1, 0), [-1, -1, -1, 0, 0, -1, 0, 0, 1]),
                   (0, 1, (-1, 0, 0), [-1, -1, -1, 0, 0.8660254, 0.5, 0, -0.8660254, -0.5])],
            circuits=[((1, (0, 0, 0)), (1, (1, 0, 0)), (0, (1, 0, 0))), ((1, (1, 0, 0)), (1, (1, 1, 0)), (0, (1, 0, 0))),
                      ((1, (0, 0, 0)), (0, (1, 0, 0)), (1, (1, 1, 0))), ((1, (0, 0, 0)), (1, (1, 1, 0)), (1, (0, 1, 0)))],
            groups=[[0]], groupInSC=True)
# samples/CrI3With2NNCoupling with the numbers AS PARSED by fileio (positional: first three J numbers
# land in xx,yy,zz; D = (-3.12,0,0) -> index 0 = x)  (SURVEY 5 "config / flags")
_J1 = [-19.49182553875, -18.47237479, -18.47237479] + [0] * 6
_J2 = [-5.7387258225, -6.1364053675, -6.1364053675] + [0] * 6
_J3 = [4.57737083, 4.474132450625, 4.474132450625] + [0] * 6
CRI3 = dict(S=[1.5, 1.5], D=[[-3.12276251875, 0, 0]] * 2, LMatrix=HEX, pos=HEXPOS,
            bonds=[(0, 1, (0, 0, 0), _J1), (0, 1, (-1, 0, 0), _J1), (0, 1, (0, 1, 0), _J1),
                   (0, 0, (1, 0, 0), _J2), (0, 0, (1, 1, 0), _J2), (0, 0, (0, 1, 0), _J2),
                   (1, 1, (1, 0, 0), _J2), (1, 1, (1, 1, 0), _J2), (1, 1, (0, 1, 0), _J2),
                   (0, 1, (-1, 1, 0), _J3), (0, 1, (1, 1, 0), _J3), (0, 1, (-1, -1, 0), _J3)],
            groups=[[0], [1]], groupInSC=True)
_JI = [-1, -1, -1] + [0] * 6
SQUARE = dict(S=[1.0], bonds=[(0, 0, (1, 0, 0), _JI), (0, 0, (0, 1, 0), _JI)], groups=[[0]], groupInSC=True)
CUBIC = dict(S=[1.0], bonds=[(0, 0, (1, 0, 0), _JI), (0, 0, (0, 1, 0), _JI), (0, 0, (0, 0, 1), _JI)])
# an anisotropic test model exercising every tensor slot, D and h (not a sample; for KATs)
_JA = [-1.0, -0.8, -1.3, 0.21, -0.33, 0.12, -0.17, 0.29, 0.05]
_JB = [0.4, -0.6, 0.5, -0.11, 0.07, 0.19, 0.23, -0.31, 0.13]
ANISO = dict(S=[1.0, 1.5], D=[[0.1, -0.2, 0.3], [-0.15, 0.05, 0.25]], LMatrix=HEX, pos=HEXPOS,
             bonds=[(0, 1, (0, 0, 0), _JA), (0, 1, (0, 1, 0), _JB), (0, 1, (-1, 0, 0), _JA), (0, 0, (1, 0, 0), _JB),
                    (1, 1, (0, 1, 0), _JA)], pair=(0, 1, (1, 0, 0)),
             circuits=[((0, (0, 0, 0)), (1, (0, 0, 0)), (0, (1, 0, 0))), ((1, (0, 0, 0)), (0, (0, 1, 0)), (1, (1, 1, 0)))],
             groups=[[0], [1]], groupInSC=True)

SPECS = {"skyrmion": SKYR, "cri3": CRI3, "square": SQUARE, "cubic": CUBIC, "aniso": ANISO}


def spec_of(name, L, **over):
    d = dict(SPECS[name])
    d.update(over)
    return LatticeSpec(L=L, **d)


