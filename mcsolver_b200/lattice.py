"""Lattice descriptor and neighbour-table generator (host side of the engine boundary).

What it replaces in the reference: the Python object-graph builder `Lattice.py:155-284`
(`establishLattice`, `establishLinking`) plus the tuple flattening in `mcMain.py:150-223`
(O(n)) / `mcMain.py:56-105` (Ising).  The reference builds one Python object per orbital
(~80 us and ~3.3 kB each) and cannot express 4096^2 / 256^3; this module keeps a lattice
as *bond templates + supercell dims* (`LatticeSpec`) and only expands to flat numpy tables
(`build_tables`) when the legacy per-site layout is wanted.  The CUDA engine's structured path
consumes the compact `LatticeSpec` directly (include/mcsolver_b200.h: mcg_lattice_desc).

Conventions reproduced from the reference (SURVEY 8 "Sign/units conventions"):
  * site id = ((x*Ly + y)*Lz + z)*norb + o                      (Lattice.py:171-184)
  * every coupling handed to the engine is divided by T, T floored at 0.1 (mcMain.py:21-31)
  * J flat order xx,yy,zz,xy,xz,yz,yx,zx,zy; the target's copy of a bond carries J^T
    (`invStrength`, Lattice.py:136-137, 260-262)
  * periodic boundaries always; a bond whose target is already linked is merged/skipped
    (Lattice.py:36-52); per-site link ORDER equals the reference's insertion order
  * "chosen" block-spin sites: x,y,z all even (Lattice.py:185), 2x2x2 clusters (:197-205)
  * correlated pairs (ki_s in cell, ki_t in cell+overLat) (Lattice.py:273)
  * triangle circuits with PBC (Lattice.py:223-234); orbital groups (:207-211)
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

# transpose permutation of the 9-vector (xx,yy,zz,xy,xz,yz,yx,zx,zy) -> Lattice.py:136-137
_T9 = np.array([0, 1, 2, 6, 7, 8, 3, 4, 5])


@dataclass
class LatticeSpec:
    """Compact, translation-invariant description of a spin lattice (what a parameter file holds)."""
    L: Tuple[int, int, int]
    S: Sequence[float]                                   # per orbital, signed initial spin
    D: Sequence[Sequence[float]] = None                  # per orbital (Dx,Dy,Dz) as parsed
    bonds: Sequence[tuple] = ()                          # (src, tgt, (n1,n2,n3), J9 | J)
    LMatrix: Sequence[Sequence[float]] = ((1, 0, 0), (0, 1, 0), (0, 0, 1))
    pos: Sequence[Sequence[float]] = None                # fractional positions per orbital
    pair: tuple = (0, 0, (0, 0, 0))                      # (ki_s, ki_t, overLat)   fileio GcOrb
    groups: Sequence[Sequence[int]] = ()                 # orbGroupList
    groupInSC: bool = False
    circuits: Sequence[tuple] = ()                       # ((o1,(a,b,c)),(o2,..),(o3,..))

    def __post_init__(self):
        self.L = tuple(int(v) for v in self.L)
        self.S = [float(s) for s in self.S]
        if self.D is None:
            self.D = [[0.0, 0.0, 0.0] for _ in self.S]
        self.D = [[float(v) for v in d] for d in self.D]
        if self.pos is None:
            self.pos = [[0.0, 0.0, 0.0] for _ in self.S]
        nb = []
        for b in self.bonds:
            J = b[3]
            if np.isscalar(J):
                J = [float(J)] * 3 + [0.0] * 6
            J = [float(v) for v in J]
            if len(J) == 3:
                J = J + [0.0] * 6
            assert len(J) == 9, "bond J needs 1, 3 or 9 numbers"
            nb.append((int(b[0]), int(b[1]), tuple(int(v) for v in b[2]), J))
        self.bonds = nb
        ks, kt, kl = self.pair
        if ks >= self.norb or kt >= self.norb:
            # mcMain.py:33-35 raises a str; we raise a real exception
            raise ValueError("pair orbital index out of range ki_s=%d ki_t=%d norb=%d" % (ks, kt, self.norb))
        for b in self.bonds:
            if not (0 <= b[0] < self.norb and 0 <= b[1] < self.norb):
                raise ValueError("bond orbital index out of range: %r" % (b[:3],))

    @property
    def norb(self):
        return len(self.S)

    @property
    def ncell(self):
        return self.L[0] * self.L[1] * self.L[2]

    @property
    def nsite(self):
        return self.ncell * self.norb


@dataclass
class Tables:
    """Flat per-site tables, i.e. the payload of the reference's MCMainFunction call."""
    model: int                     # 1 Ising, 2 XY, 3 Heisenberg
    N: int
    maxL: int
    S: np.ndarray                  # [N] f64 signed
    D: np.ndarray                  # [N,3] f64 (already /T)
    nlink: np.ndarray              # [N] i32
    J: np.ndarray                  # [N,maxL,9] f64 (Ising: [N,maxL]) (already /T)
    nbr: np.ndarray                # [N,maxL] i32, -1 padded
    tri: np.ndarray                # [nTri,3] i32
    pairs: np.ndarray              # [nLat,2] i32
    groups: np.ndarray             # [nG,maxG] i32 (-1 padded)
    nG: int
    maxG: int
    rOrb: np.ndarray               # [nR] i32
    rCluster: np.ndarray           # [nR,nC] i32
    rNbr: np.ndarray               # [nR,maxL] i32
    ignoreOffDiag: int = 1
    T: float = 1.0

    def on_args(self, algorithm, nthermal, nsweep, ninterval, flunc, h_over_T, spinFrame, callback=None):
        """The 23 positional arguments of heisenbergLib.c:504-512 / xyLib.c:439-447 (python tuples)."""
        return (int(algorithm), tuple(self.S.tolist()), tuple(self.D.reshape(-1).tolist()),
                int(nthermal), int(nsweep), int(ninterval), int(self.maxL), tuple(self.nlink.tolist()),
                tuple(self.J.reshape(-1).tolist()), tuple(self.nbr.reshape(-1).tolist()),
                tuple(self.tri.reshape(-1).tolist()), tuple(self.pairs.reshape(-1).tolist()),
                int(self.nG), int(self.maxG), tuple(self.groups.reshape(-1).tolist()),
                float(flunc), float(h_over_T), tuple(self.rOrb.tolist()),
                tuple(self.rCluster.reshape(-1).tolist()), tuple(self.rNbr.reshape(-1).tolist()),
                int(spinFrame), int(self.ignoreOffDiag), callback or (lambda k: None))

    def ising_args(self, algorithm, nthermal, nsweep, ninterval, h_over_T, spinFrame, callback=None):
        """The 16 positional arguments of isingLib.c:277-284."""
        return (int(algorithm), tuple(self.S.tolist()), int(nthermal), int(nsweep), int(ninterval),
                int(self.maxL), tuple(self.nlink.tolist()), tuple(self.J.reshape(-1).tolist()),
                tuple(self.nbr.reshape(-1).tolist()), tuple(self.pairs.reshape(-1).tolist()),
                float(h_over_T), tuple(self.rOrb.tolist()), tuple(self.rCluster.reshape(-1).tolist()),
                tuple(self.rNbr.reshape(-1).tolist()), int(spinFrame), callback or (lambda k: None))


def _site_id(spec, x, y, z, o):
    Lx, Ly, Lz = spec.L
    return ((x % Lx * Ly + y % Ly) * Lz + z % Lz) * spec.norb + o


def _merge_links(entries, ising):
    """Sequential merge rule of Orbital.addLinking (Lattice.py:34-55) on an ordered entry list."""
    ids, Js = [], []
    for t, J in entries:
        if t in ids:
            k = ids.index(t)
            diff = abs(Js[k] - J) if ising else float(np.sum(np.abs(Js[k] - J)))
            if diff < 1e-5:
                continue
            Js[k] = Js[k] + J
            continue
        ids.append(t)
        Js.append(J)
    return ids, Js


def _ordered_link_entries(spec, T, ising):
    """Vectorised: for every site the ordered list of (partner id, J) 'addLinking' calls the
    reference would issue (Lattice.py:247-262).  Returns partner[N,E], Jv[N,E,9|1], valid[N,E]."""
    Lx, Ly, Lz = spec.L
    no = spec.norb
    nc = spec.ncell
    cx, cy, cz = np.meshgrid(np.arange(Lx), np.arange(Ly), np.arange(Lz), indexing="ij")
    cx, cy, cz = cx.ravel(), cy.ravel(), cz.ravel()
    clin = (cx * Ly + cy) * Lz + cz
    nb = len(spec.bonds)
    per_orb = []
    Emax = 0
    for o in range(no):
        ent = [(b, 0) for b in range(nb) if spec.bonds[b][0] == o] + \
              [(b, 1) for b in range(nb) if spec.bonds[b][1] == o]
        per_orb.append(ent)
        Emax = max(Emax, len(ent))
    N = spec.nsite
    jd = 1 if ising else 9
    partner = np.full((N, Emax), -1, dtype=np.int64)
    Jv = np.zeros((N, Emax, jd))
    key = np.full((N, Emax), np.iinfo(np.int64).max, dtype=np.int64)
    for o in range(no):
        sid = clin * no + o
        for e, (b, act) in enumerate(per_orb[o]):
            src, tgt, d, J9 = spec.bonds[b]
            J9 = np.asarray(J9) * (1.0 / T)
            if act == 0:                      # this site is the bond's source
                px, py, pz = (cx + d[0]) % Lx, (cy + d[1]) % Ly, (cz + d[2]) % Lz
                pid = ((px * Ly + py) * Lz + pz) * no + tgt
                k = (clin * no + o) * nb + b
                partner[sid, e] = pid
                Jv[sid, e] = J9[:jd] if not ising else J9[0]
                key[sid, e] = k * 2
            else:                             # this site is the bond's target: link back with J^T
                px, py, pz = (cx - d[0]) % Lx, (cy - d[1]) % Ly, (cz - d[2]) % Lz
                pid = ((px * Ly + py) * Lz + pz) * no + src
                k = (((px * Ly + py) * Lz + pz) * no + src) * nb + b
                ok = pid != sid               # Lattice.py:261: no back link when source is target
                partner[sid, e] = np.where(ok, pid, -1)
                Jv[sid, e] = (J9[_T9] if not ising else J9[0])
                key[sid, e] = np.where(ok, k * 2 + 1, np.iinfo(np.int64).max)
    order = np.argsort(key, axis=1, kind="stable")
    partner = np.take_along_axis(partner, order, axis=1)
    Jv = np.take_along_axis(Jv, order[:, :, None], axis=1)
    return partner, Jv


def build_tables(spec: LatticeSpec, T: float = 1.0, model: int = 3) -> Tables:
    """Expand a LatticeSpec into the reference's flat per-site tables at temperature T
    (everything pre-multiplied by 1/T exactly as mcMain.py:21-31 does)."""
    T = 0.1 if T < 0.1 else float(T)
    ising = model == 1
    Lx, Ly, Lz = spec.L
    no, N = spec.norb, spec.nsite
    partner, Jv = _ordered_link_entries(spec, T, ising)
    E = partner.shape[1]
    # duplicates (only when a supercell dim is 1 or 2, or two templates hit the same pair)
    srt = np.sort(np.where(partner < 0, -np.arange(1, E + 1)[None, :], partner), axis=1)
    dup_rows = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0] if E > 1 else np.array([], dtype=int)
    nlink = (partner >= 0).sum(axis=1).astype(np.int32)
    if len(dup_rows):
        merged = {}
        for i in dup_rows:
            ents = [(int(partner[i, e]), (float(Jv[i, e, 0]) if ising else Jv[i, e].copy()))
                    for e in range(E) if partner[i, e] >= 0]
            merged[int(i)] = _merge_links(ents, ising)
            nlink[i] = len(merged[int(i)][0])
    maxL = int(nlink.max()) if N else 0
    jd = 1 if ising else 9
    nbr = np.full((N, maxL), -1, dtype=np.int32)
    J = np.zeros((N, maxL, jd))
    # compact valid entries to the left preserving order
    valid = partner >= 0
    pos_in_row = np.cumsum(valid, axis=1) - 1
    rows = np.nonzero(valid)
    keep = np.ones(len(rows[0]), dtype=bool)
    if len(dup_rows):
        keep = ~np.isin(rows[0], dup_rows)
    r, c = rows[0][keep], rows[1][keep]
    nbr[r, pos_in_row[r, c]] = partner[r, c]
    J[r, pos_in_row[r, c]] = Jv[r, c]
    for i, (ids, Js) in (merged.items() if len(dup_rows) else ()):
        for k, (t, jj) in enumerate(zip(ids, Js)):
            nbr[i, k] = t
            J[i, k] = jj
    if ising:
        J = J[:, :, 0]
    ignore = 1
    if not ising and np.any(np.abs(J[:, :, 3:]) > 1e-6):      # mcMain.py:174
        ignore = 0

    sid = np.arange(N)
    o_of = sid % no
    cell = sid // no
    cz = cell % Lz
    cy = (cell // Lz) % Ly
    cx = cell // (Lz * Ly)
    S = np.asarray(spec.S, dtype=float)[o_of]
    D = (np.asarray(spec.D, dtype=float) / T)[o_of]

    # correlated pairs, one per cell (Lattice.py:273)
    ks, kt, kl = spec.pair
    c = np.arange(spec.ncell)
    pz, py, px = c % Lz, (c // Lz) % Ly, c // (Lz * Ly)
    pi = c * no + ks
    pj = ((((px + kl[0]) % Lx) * Ly + (py + kl[1]) % Ly) * Lz + (pz + kl[2]) % Lz) * no + kt
    pairs = np.stack([pi, pj], axis=1).astype(np.int32)

    # triangle circuits (Lattice.py:223-234): per cell, per circuit template
    tri = np.zeros((0, 3), dtype=np.int32)
    if len(spec.circuits):
        cols = []
        for circ in spec.circuits:
            ids = []
            for (orb, dl) in circ:
                ids.append(((((px + dl[0]) % Lx) * Ly + (py + dl[1]) % Ly) * Lz + (pz + dl[2]) % Lz) * no + orb)
            cols.append(np.stack(ids, axis=1))
        tri = np.stack(cols, axis=1).reshape(-1, 3).astype(np.int32)

    # orbital groups (Lattice.py:207-211): the comprehension order is id, x, y, z
    glist = []
    for sub in spec.groups:
        if spec.groupInSC:
            g = np.concatenate([c * no + int(o) for o in sub]) if len(sub) else np.zeros(0, dtype=int)
        else:
            g = np.array([int(o) for o in sub], dtype=int)
        glist.append(g)
    nG = len(glist)
    maxG = 1
    if nG > 0:
        maxG = max(len(g) for g in glist)          # mcMain.py:186-191
    groups = np.full((nG, maxG), -1, dtype=np.int32)
    for k, g in enumerate(glist):
        groups[k, :len(g)] = g

    # block-spin ("renormalised") tables: chosen sites, 2x2x2 clusters, doubled-bond links
    chosen = (cx % 2 + cy % 2 + cz % 2) == 0
    rOrb = sid[chosen].astype(np.int32)
    rx, ry, rz, ro = cx[chosen], cy[chosen], cz[chosen], o_of[chosen]
    offs = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)]  # Lattice.py:198-205
    cl = np.stack([((((rx + a) % Lx) * Ly + (ry + b) % Ly) * Lz + (rz + cc) % Lz) * no + ro for a, b, cc in offs], axis=1)
    # addOrbIntoCluster dedupes (dimension of size 1): keep first occurrences, same count for all rows
    keepc = [0]
    for k in range(1, 8):
        if not any(np.array_equal(cl[:, k], cl[:, j]) for j in keepc):
            keepc.append(k)
    rCluster = cl[:, keepc].astype(np.int32)
    rNbr = _rnorm_links(spec, rOrb, maxL)
    return Tables(model=model, N=N, maxL=maxL, S=S, D=D, nlink=nlink, J=J, nbr=nbr, tri=tri, pairs=pairs,
                  groups=groups, nG=nG, maxG=maxG, rOrb=rOrb, rCluster=rCluster, rNbr=rNbr,
                  ignoreOffDiag=ignore, T=T)


def _rnorm_links(spec, rOrb, maxL):
    """linkedOrb_rnorm of the chosen sites (Lattice.py:265-271): bonds with doubled overLat,
    insertion-ordered, duplicates by id skipped (addLinking_rnorm :60-66); -1 padded to maxL."""
    Lx, Ly, Lz = spec.L
    no, nb = spec.norb, len(spec.bonds)
    nR = len(rOrb)
    out = np.full((nR, maxL), -1, dtype=np.int32)
    if nR == 0 or maxL == 0:
        return out
    row_of = {int(s): i for i, s in enumerate(rOrb)}
    chosen = set(row_of)
    ids = [[] for _ in range(nR)]
    all_sites_links = {}
    for s in range(spec.nsite):                 # reference visits every orbital in id order
        if s not in chosen:
            continue
        o = s % no
        c = s // no
        z, y, x = c % Lz, (c // Lz) % Ly, c // (Lz * Ly)
        for b in range(nb):
            src, tgt, d, _ = spec.bonds[b]
            if src != o:
                continue
            t = _site_id(spec, x + 2 * d[0], y + 2 * d[1], z + 2 * d[2], tgt)
            lst = all_sites_links.setdefault(s, [])
            if t not in lst:
                lst.append(t)
            if t != s:
                lt = all_sites_links.setdefault(t, [])
                if s not in lt:
                    lt.append(s)
    for s, i in row_of.items():
        lst = all_sites_links.get(s, [])[:maxL]
        out[i, :len(lst)] = lst
    return out


def positions(spec: LatticeSpec) -> np.ndarray:
    """Cartesian positions in site-id order (Lattice.py:179): (cell + frac) @ LMatrix."""
    Lx, Ly, Lz = spec.L
    no = spec.norb
    c = np.arange(spec.ncell)
    cell = np.stack([c // (Lz * Ly), (c // Lz) % Ly, c % Lz], axis=1).astype(float)
    frac = np.asarray(spec.pos, dtype=float)
    p = (cell[:, None, :] + frac[None, :, :]).reshape(-1, 3)
    return p @ np.asarray(spec.LMatrix, dtype=float)


# ---------------------------------------------------------------------------------------------
# dipole-dipole coupling
# ---------------------------------------------------------------------------------------------
def dipole_tensor(r, alpha, ising=False):
    """J_dip = alpha/|r|^3 (1 - 3 r^ r^T) in the 9-vector order xx,yy,zz,xy,xz,yz,yx,zx,zy
    (Lattice.py:300-305); Ising keeps only alpha/|r|^3 (Lattice.py:296-298)."""
    r = np.asarray(r, dtype=float)
    d = float(np.sqrt(np.dot(r, r)))
    a = alpha / d ** 3
    if ising:
        return a
    x, y, z = r / d
    return np.array([a * (1 - 3 * x * x), a * (1 - 3 * y * y), a * (1 - 3 * z * z), -3 * a * x * y, -3 * a * x * z, -3 * a * y * z,
                     -3 * a * y * x, -3 * a * z * x, -3 * a * z * y])


def add_dipole_all_pairs(spec: LatticeSpec, t: Tables, alpha_over_T: float) -> Tables:
    """The reference's DEFINITION of the dipole term (Lattice.py:286-305): every ordered pair of distinct
    orbitals gets an extra link appended (forceAdd: never merged) with open-boundary Cartesian separation
    r = r_source - r_target.  O(N^2) links: reference-feasible sizes only (the reference's own code path
    raises TypeError; tests pin this against the reference loop with the missing `distance` argument supplied)."""
    N = t.N
    ising = t.model == 1
    pos = positions(spec)
    extra = N - 1
    maxL = t.maxL + extra
    jd = () if ising else (9,)
    nbr = np.full((N, maxL), -1, dtype=np.int32)
    J = np.zeros((N, maxL) + jd)
    nlink = t.nlink.copy()
    nbr[:, :t.maxL] = t.nbr
    J[:, :t.maxL] = t.J
    for i in range(N):
        k = nlink[i]
        for j in range(N):
            if i == j:
                continue
            nbr[i, k] = j
            J[i, k] = dipole_tensor(pos[i] - pos[j], alpha_over_T, ising)
            k += 1
        nlink[i] = k
    ignore = t.ignoreOffDiag
    if not ising and np.any(np.abs(J[:, :, 3:]) > 1e-6):
        ignore = 0
    rNbr = np.full((t.rNbr.shape[0], maxL), -1, dtype=np.int32)
    rNbr[:, :t.rNbr.shape[1]] = t.rNbr
    return Tables(model=t.model, N=N, maxL=maxL, S=t.S, D=t.D, nlink=nlink.astype(np.int32), J=J, nbr=nbr, tri=t.tri, pairs=t.pairs,
                  groups=t.groups, nG=t.nG, maxG=t.maxG, rOrb=t.rOrb, rCluster=t.rCluster, rNbr=rNbr, ignoreOffDiag=ignore, T=t.T)


def add_dipole_stencil(spec: LatticeSpec, alpha: float, rcut: float, ising=False) -> LatticeSpec:
    """Cut-off, periodic (minimum-image) reading of the dipole term for lattices the all-pairs definition
    cannot reach: one extra bond template per unordered orbital pair with 0 < |r| <= rcut (Cartesian, in the
    units of LMatrix).  Templates that coincide with an exchange bond are merged by the engine (J adds up)."""
    LM = np.asarray(spec.LMatrix, dtype=float)
    frac = np.asarray(spec.pos, dtype=float)
    no = spec.norb
    inv = np.linalg.inv(LM)
    need = [int(np.floor(rcut * np.linalg.norm(inv[:, k]) + 1e-9)) for k in range(3)]   # cells a bond can span along axis k
    reach = [0, 0, 0]
    for k in range(3):
        if spec.L[k] == 1:
            continue
        if 2 * need[k] + 1 > spec.L[k]:    # an image and its mirror would be the same site
            raise ValueError("dipole cutoff %g needs a supercell of at least %d cells along axis %d" % (rcut, 2 * need[k] + 1, k))
        reach[k] = min(need[k] + 1, (spec.L[k] - 1) // 2)
    bonds = list(spec.bonds)
    for o in range(no):
        for o2 in range(no):
            for dx in range(-reach[0], reach[0] + 1):
                for dy in range(-reach[1], reach[1] + 1):
                    for dz in range(-reach[2], reach[2] + 1):
                        # each unordered pair once: (o,o2,d) and (o2,o,-d) are the same bond
                        if o2 < o or (o2 == o and (dx, dy, dz) <= (0, 0, 0)):
                            continue
                        r = (np.array([dx, dy, dz]) + frac[o2] - frac[o]) @ LM
                        d = np.sqrt(np.dot(r, r))
                        if d < 1e-9 or d > rcut + 1e-9:
                            continue
                        Jd = dipole_tensor(r, alpha, ising)
                        J9 = [float(Jd)] + [0.0] * 8 if ising else [float(v) for v in Jd]
                        bonds.append((o, o2, (dx, dy, dz), J9))
    return LatticeSpec(L=spec.L, S=spec.S, D=spec.D, bonds=bonds, LMatrix=spec.LMatrix, pos=spec.pos, pair=spec.pair, groups=spec.groups,
                       groupInSC=spec.groupInSC, circuits=spec.circuits)
