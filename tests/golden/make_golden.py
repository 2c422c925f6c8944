"""Generates the committed golden fixtures from the UNMODIFIED reference, in the build container.

Needs /root/reference (host Python, imported with tkinter/matplotlib stubbed) and oracle/_ref
(the reference C engines compiled by oracle/Makefile; xylib with the 2-line seedID fix).
Run:  python tests/golden/make_golden.py [tables] [kat] [runs] [stats] [dipole] [u4cross] [stats32] [refhost]
Writes tests/golden/{tables,kat,runs,stats}.json(+npz).  None of the tests reads
/root/reference at run time: they read these files.

  tables : the reference's own flattened MCMainFunction arguments (Lattice.py + mcMain.py) for
           small lattices -> pins mcsolver_b200.lattice.build_tables
  kat    : config-level known answers: a pseudo-random configuration produced by the
           reference's own init (flunc=1e6, srand(k)) and the reference's fp64 observables of
           exactly that configuration (nthermal=0, ninterval=0, nsweep=1, spinFrame=1)
  runs   : whole seeded runs of the reference engines (srand(k)) -> pins the oracle restatement
           (same rand() call sequence => same trajectory => same result tuple)
  stats  : K independent seeded reference runs per (model, T, H) point -> mean and sigma of the
           equilibrium observables for the 3-sigma statistical parity tests of the CUDA engine
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness as rh  # noqa: E402
from mcsolver_b200.lattice import build_tables  # noqa: E402

from tests.specs import SPECS, spec_of  # noqa: E402,F401


def _ref_args(spec, T, model, h=0.0, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return rh.reference_tables(spec.LMatrix, spec.pos, spec.S, spec.D, spec.bonds, T=T, L=spec.L, ki=spec.pair,
                                   orbGroupList=spec.groups, groupInSC=spec.groupInSC, h=h, On=model,
                                   circuits=spec.circuits, **kw)


def _jsonable(a):
    def cv(x):
        if isinstance(x, tuple):
            return [cv(v) for v in x]
        if isinstance(x, (np.integer,)):
            return int(x)
        if isinstance(x, (np.floating,)):
            return float(x)
        return x
    return [cv(x) for x in a if not callable(x)]


def make_tables():
    cases = [("skyrmion", (4, 6, 1), 0.3, 3, 0.2), ("skyrmion", (2, 2, 1), 0.3, 3, 0.0), ("skyrmion", (1, 1, 1), 0.3, 3, 0.0),
             ("cri3", (4, 4, 1), 35.0, 3, 0.0), ("cri3", (2, 2, 1), 35.0, 3, 0.0), ("cri3", (3, 3, 1), 35.0, 3, 0.0),
             ("square", (6, 4, 1), 0.9, 2, 0.0), ("square", (2, 2, 1), 2.2, 1, 0.1), ("square", (1, 4, 1), 2.2, 1, 0.1),
             ("square", (5, 3, 1), 2.2, 1, 0.0), ("cubic", (4, 4, 4), 1.4, 3, 0.0), ("cubic", (2, 4, 6), 4.4, 1, 0.0),
             ("cubic", (3, 3, 3), 1.4, 3, 0.0), ("aniso", (4, 4, 2), 0.7, 3, 0.3), ("aniso", (3, 4, 1), 0.7, 2, 0.3)]
    out = []
    for name, L, T, model, h in cases:
        a = _ref_args(spec_of(name, L), T, model, h=h)
        out.append(dict(spec=name, L=L, T=T, model=model, h=h, args=_jsonable(a)))
    json.dump(out, open(os.path.join(HERE, "tables.json"), "w"))
    print("tables.json:", len(out), "cases")


def make_kat():
    cases = [("skyrmion", (8, 8, 1), 0.3, 3, 0.2, 5), ("cri3", (6, 6, 1), 35.0, 3, 0.0, 6), ("cubic", (6, 6, 6), 1.4, 3, 0.0, 7),
             ("aniso", (6, 6, 2), 0.7, 3, 0.3, 8), ("aniso", (6, 6, 1), 0.7, 2, 0.3, 9), ("square", (8, 8, 1), 0.9, 2, 0.05, 10)]
    meta, arrays = [], {}
    for idx, (name, L, T, model, h, seed) in enumerate(cases):
        a = list(_ref_args(spec_of(name, L), T, model, h=h, spinFrame=1, flunc=1e6))
        a[3], a[4], a[5] = 0, 1, 0  # nthermal=0, nsweep=1, ninterval=0: no update, one measurement
        res = rh.run_ref_engine(model, tuple(a), seed=seed)
        arrays["spins%d" % idx] = np.array(res[27][0])
        arrays["out%d" % idx] = np.array(res[:27])
        arrays["group%d" % idx] = np.array(res[28]) if not isinstance(res[28], float) else np.zeros(0)
        meta.append(dict(spec=name, L=L, T=T, model=model, h=h, seed=seed))
    # Ising: the configuration is the input; absolute energy through one accepted Wolff flip
    # (isingLib.c:230-233): algorithm=1, nsweep=1, ninterval=1, spinFrame=1 -> frame = post-flip state
    for name, L, T, h, seed in [("square", (8, 8, 1), 2.0, 0.0, 11), ("cubic", (6, 6, 6), 4.0, 0.0, 12)]:
        idx = len(meta)
        a = list(_ref_args(spec_of(name, L), T, 1, h=h, spinFrame=1, algo="Wolff"))
        rng = np.random.RandomState(seed)
        a[1] = tuple(float(v) for v in rng.choice([-1.0, 1.0], size=len(a[1])))
        a[2], a[3], a[4] = 0, 1, 1
        res = rh.run_ref_engine(1, tuple(a), seed=seed)
        arrays["spins%d" % idx] = np.array(res[10][0])
        arrays["out%d" % idx] = np.array(res[:10])
        arrays["group%d" % idx] = np.zeros(0)
        meta.append(dict(spec=name, L=L, T=T, model=1, h=h, seed=seed))
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("kat:", len(meta), "cases")


RUN_CASES = [
    # name, L, T, model, algo, nthermal, nsweep, ninterval, h, flunc, frames, seed
    ("skyrmion", (6, 6, 1), 0.3, 3, 0, 50, 100, 72, 0.2, 0.0, 2, 3),
    ("skyrmion", (6, 6, 1), 0.3, 3, 0, 5, 10, 72, 0.2, 0.7, 1, 3),
    ("cubic", (6, 6, 6), 1.4, 3, 0, 20, 50, 216, 0.1, 0.0, 0, 4),
    ("cubic", (6, 6, 6), 1.4, 3, 1, 20, 50, 3, 0.0, 0.0, 0, 5),
    ("skyrmion", (6, 6, 1), 0.3, 3, 1, 20, 50, 2, 0.2, 0.0, 0, 6),
    ("aniso", (6, 6, 1), 0.7, 3, 0, 20, 60, 72, 0.3, 0.2, 0, 7),
    ("aniso", (6, 6, 1), 0.7, 2, 0, 20, 60, 72, 0.3, 0.2, 0, 8),
    ("square", (8, 8, 1), 0.9, 2, 0, 20, 100, 64, 0.05, 0.0, 0, 9),
    ("square", (8, 8, 1), 0.9, 2, 1, 100, 200, 1, 0.0, 0.0, 0, 10),
    ("square", (8, 8, 1), 2.2, 1, 0, 20, 100, 64, 0.05, 0.0, 2, 11),
    ("square", (8, 8, 1), 2.2, 1, 1, 100, 200, 1, 0.02, 0.0, 0, 12),
    ("cubic", (6, 6, 6), 4.4, 1, 1, 50, 100, 2, 0.0, 0.0, 0, 13),
]


def make_runs():
    meta, arrays = [], {}
    for idx, (name, L, T, model, algo, nth, nsw, nint, h, flunc, frames, seed) in enumerate(RUN_CASES):
        t = build_tables(spec_of(name, L), T, model)
        hT = h / max(T, 0.1)
        if model == 1:
            args = t.ising_args(algo, nth, nsw, nint, hT, frames)
            res = rh.run_ref_engine(1, args, seed=seed)
            arrays["out%d" % idx] = np.array(res[:10])
            fr = res[10]
        else:
            args = t.on_args(algo, nth, nsw, nint, flunc, hT, frames)
            res = rh.run_ref_engine(model, args, seed=seed)
            arrays["out%d" % idx] = np.array(res[:27])
            arrays["group%d" % idx] = np.array(res[28]) if not isinstance(res[28], float) else np.zeros(0)
            fr = res[27]
        arrays["frames%d" % idx] = np.array(fr) if frames else np.zeros(0)
        meta.append(dict(spec=name, L=L, T=T, model=model, algo=algo, nthermal=nth, nsweep=nsw, ninterval=nint, h=h,
                         flunc=flunc, frames=frames, seed=seed))
    np.savez_compressed(os.path.join(HERE, "runs.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "runs.json"), "w"), indent=1)
    print("runs:", len(meta), "cases")


STAT_POINTS = [
    # tag, spec, L, model, algo, T, H, nthermal(sweeps), nsweep, K seeds.  48: with 8 or 16 seeds the sample means of two points (C4 T=0.6, Ising T=2.4 next to
    # Tc) sat 3 sigma off the 96-seed mean of the same reference engine and the 3-sigma test raised false alarms
    ("C1_xy_metro", "square", (16, 16, 1), 2, 0, 0.7, 0.0, 1000, 4000, 48),
    ("C1_xy_metro", "square", (16, 16, 1), 2, 0, 0.9, 0.0, 1000, 4000, 48),
    ("C1_xy_metro", "square", (16, 16, 1), 2, 0, 1.2, 0.0, 1000, 4000, 48),
    ("C1_xy_wolff", "square", (16, 16, 1), 2, 1, 0.9, 0.0, 4000, 16000, 48),
    ("C2_ising_metro", "square", (24, 24, 1), 1, 0, 2.0, 0.0, 1000, 4000, 48),
    ("C2_ising_metro", "square", (24, 24, 1), 1, 0, 2.4, 0.0, 1000, 4000, 48),
    ("C2_ising_metro", "square", (24, 24, 1), 1, 0, 3.0, 0.1, 1000, 4000, 48),
    ("C2_ising_wolff", "square", (24, 24, 1), 1, 1, 2.269, 0.0, 2000, 8000, 48),
    ("C3_cri3", "cri3", (8, 8, 1), 3, 0, 30.0, 0.0, 1000, 4000, 48),
    ("C3_cri3", "cri3", (8, 8, 1), 3, 0, 45.0, 0.0, 1000, 4000, 48),
    ("C3_cri3", "cri3", (8, 8, 1), 3, 0, 60.0, 0.0, 1000, 4000, 48),
    ("C4_skyrmion", "skyrmion", (12, 12, 1), 3, 0, 0.3, 0.0, 2000, 4000, 48),
    ("C4_skyrmion", "skyrmion", (12, 12, 1), 3, 0, 0.3, 0.3, 2000, 4000, 48),
    ("C4_skyrmion", "skyrmion", (12, 12, 1), 3, 0, 0.6, 0.3, 2000, 4000, 48),
    ("C5_cubic", "cubic", (8, 8, 8), 3, 0, 1.2, 0.0, 1000, 3000, 48),
    ("C5_cubic", "cubic", (8, 8, 8), 3, 0, 1.45, 0.0, 1000, 3000, 48),
    ("C5_cubic", "cubic", (8, 8, 8), 3, 0, 1.8, 0.0, 1000, 3000, 48),
    ("C5_cubic_wolff", "cubic", (8, 8, 8), 3, 1, 1.45, 0.0, 500, 2000, 48),
    ("aniso_heis", "aniso", (8, 8, 1), 3, 0, 0.7, 0.3, 1000, 4000, 48),
    ("aniso_xy", "aniso", (8, 8, 1), 2, 0, 0.7, 0.3, 1000, 4000, 48),
]


def make_stats():
    out = []
    t0 = time.time()
    for tag, name, L, model, algo, T, H, nth, nsw, K in STAT_POINTS:
        spec = spec_of(name, L)
        t = build_tables(spec, T, model)
        N = t.N
        nint = N if algo == 0 else (1 if model != 3 or name != "cubic" else 4)
        hT = H / max(T, 0.1)
        rows = []
        for k in range(1, K + 1):
            if model == 1:
                res = rh.run_ref_engine(1, t.ising_args(algo, nth, nsw, nint, hT, 0), seed=k)
                rows.append(list(res[:10]))
            else:
                res = rh.run_ref_engine(model, t.on_args(algo, nth, nsw, nint, 0.0, hT, 0), seed=k)
                rows.append(list(res[:27]))
        rows = np.array(rows)
        out.append(dict(tag=tag, spec=name, L=L, model=model, algo=algo, T=T, H=H, nthermal=nth, nsweep=nsw,
                        ninterval=nint, K=K, mean=rows.mean(axis=0).tolist(), sigma=rows.std(axis=0, ddof=1).tolist(),
                        rows=rows.tolist()))
        print("%s T=%g H=%g done (%.0fs)" % (tag, T, H, time.time() - t0), flush=True)
    json.dump(out, open(os.path.join(HERE, "stats.json"), "w"))
    print("stats:", len(out), "points")


def make_dipole():
    """Ising + dipole through the reference's own all-pairs loop (with the missing `distance` argument
    supplied at run time, refharness.patch_reference_dipole): the flattened tables."""
    rh.patch_reference_dipole()
    out = []
    arrays = {}
    for idx, (name, L, T, alpha, seed) in enumerate([("square", (4, 4, 1), 2.0, 0.3, 21), ("cubic", (3, 3, 2), 3.0, 0.5, 22)]):
        spec = spec_of(name, L)
        with contextlib.redirect_stdout(io.StringIO()):
            a = list(rh.reference_tables(spec.LMatrix, spec.pos, spec.S, spec.D, spec.bonds, T=T, L=spec.L, ki=spec.pair,
                                         orbGroupList=spec.groups, groupInSC=spec.groupInSC, h=0.0, On=1, dipoleAlpha=alpha))
        # The reference ENGINE cannot run these tables: the block-spin link table is -1 padded up to the
        # new maxNLinking and isingLib.c:111-113 dereferences lattice[-1] (segfault).  Only the tables are pinned.
        out.append(dict(spec=name, L=L, T=T, alpha=alpha, args=_jsonable(a)))
    json.dump(out, open(os.path.join(HERE, "dipole.json"), "w"))
    print("dipole:", len(out), "cases")


U4_T = [1.30, 1.36, 1.42, 1.48, 1.54, 1.60]
U4_SIZES = [6, 10]
U4_SWEEPS = (2000, 8000)
U4_K = 6


def _u4_job(job):
    L, T, seed = job
    t = build_tables(spec_of("cubic", (L, L, L)), T, 3)
    res = rh.run_ref_engine(3, t.on_args(0, U4_SWEEPS[0], U4_SWEEPS[1], t.N, 0.0, 0.0, 0), seed=seed)
    return L, T, seed, res[10], res[8] * T


def make_u4cross():
    """Tc from the U4 crossing (north_star): the reference's 3D Heisenberg runs at two sizes over a T grid,
    K seeds each.  U4 here is the reference's <M^2>^2/<M^4> (heisenbergLib.c:833)."""
    import multiprocessing as mp
    jobs = [(L, T, k) for L in U4_SIZES for T in U4_T for k in range(1, U4_K + 1)]
    t0 = time.time()
    with mp.get_context("fork").Pool(processes=min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_u4_job, jobs, chunksize=1)
    out = dict(T=U4_T, sizes=U4_SIZES, nthermal=U4_SWEEPS[0], nsweep=U4_SWEEPS[1], K=U4_K,
               U4={str(L): [[r[3] for r in res if r[0] == L and r[1] == T] for T in U4_T] for L in U4_SIZES},
               E={str(L): [[r[4] for r in res if r[0] == L and r[1] == T] for T in U4_T] for L in U4_SIZES})
    json.dump(out, open(os.path.join(HERE, "u4cross.json"), "w"))
    print("u4cross: %d runs in %.0fs" % (len(jobs), time.time() - t0))


# ---- the size at which the engine's DEFAULT path is the NVRTC-specialised colour pass (N*R >= 2^20): sc 32^3, three
# temperatures x 16 seeds = 48 replicas in one batch on the GPU side ----
S32_T = [1.30, 1.45, 1.70]
S32_L = 32
S32_SWEEPS = (600, 2400)
S32_K = 16


def _s32_job(job):
    T, seed = job
    t = build_tables(spec_of("cubic", (S32_L,) * 3), T, 3)
    res = rh.run_ref_engine(3, t.on_args(0, S32_SWEEPS[0], S32_SWEEPS[1], t.N, 0.0, 0.0, 0), seed=seed)
    return T, seed, list(res[:27])


def make_stats32():
    import multiprocessing as mp
    jobs = [(T, k) for T in S32_T for k in range(1, S32_K + 1)]
    t0 = time.time()
    with mp.get_context("fork").Pool(processes=min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_s32_job, jobs, chunksize=1)
    out = []
    for T in S32_T:
        rows = np.array([r[2] for r in res if r[0] == T])
        out.append(dict(tag="C5_cubic32", spec="cubic", L=(S32_L,) * 3, model=3, algo=0, T=T, H=0.0, nthermal=S32_SWEEPS[0],
                        nsweep=S32_SWEEPS[1], ninterval=S32_L ** 3, K=S32_K, mean=rows.mean(axis=0).tolist(),
                        sigma=rows.std(axis=0, ddof=1).tolist(), rows=rows.tolist()))
    json.dump(out, open(os.path.join(HERE, "stats32.json"), "w"))
    print("stats32: %d runs in %.0fs" % (len(jobs), time.time() - t0))


# ---- the REAL call chain, reference host + reference engines: win.startSimulation(updateGUI=False, rpath=<sample>) ----
REFHOST_K = 12


def _refhost_run(job):
    """One whole run of the reference (host Python + its compiled engines) on an edited sample file, in a scratch
    directory; returns the three output files verbatim."""
    name, seed = job
    import tempfile
    from tests.refhost_cases import CASES, edited
    sys.path.insert(0, rh.REF_DIR)                       # `from xylib import MCMainFunction` -> the reference's C engine
    for m in ("isinglib", "xylib", "heisenberglib"):
        sys.modules.pop(m, None)
    Lattice, mcMain, win, fileio = rh.load_reference_host()
    d = tempfile.mkdtemp(prefix="refhost_")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        open("param", "w").write(edited(open(rh.sample_file(name)).read(), **CASES[name]))
        rh.srand(seed)                                   # forked Pool workers inherit the stream
        fd = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(fd, 1)
        try:
            win.startSimulation(updateGUI=False, rpath="param")
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(fd)
            os.close(saved)
        return name, seed, {f: open(f).read() for f in ("result.txt", "out", "spinDotSpin.txt")}
    finally:
        os.chdir(cwd)


def make_refhost():
    """Reference-produced golden result.txt / out / spinDotSpin.txt (SURVEY 4 'interface' row; win.py:152-156,
    mcMain.py:261-265, 274-289): K seeded whole runs per sample for the 3-sigma comparison of every result.txt column,
    and the files of seed 1 verbatim for the byte-level layout checks."""
    import multiprocessing as mp
    from tests.refhost_cases import CASES, parse_result
    out = {}
    t0 = time.time()
    for name in CASES:
        runs = []
        for seed in range(1, REFHOST_K + 1):             # one at a time, each in a fresh process: startSimulation brings its own Pool
            ctx = mp.get_context("fork")
            rx, tx = ctx.Pipe(duplex=False)
            pr = ctx.Process(target=lambda: tx.send(_refhost_run((name, seed))))
            pr.start()
            runs.append(rx.recv())
            pr.join()
        rows = [parse_result(r[2]["result.txt"])[1] for r in runs]
        out[name] = dict(edits=CASES[name], K=REFHOST_K, rows=rows, files=runs[0][2])
        print("%s: %d runs (%.0fs)" % (name, len(runs), time.time() - t0), flush=True)
    json.dump(out, open(os.path.join(HERE, "refhost.json"), "w"))


if __name__ == "__main__":
    what = sys.argv[1:] or ["tables", "kat", "runs", "stats", "dipole"]
    assert rh.have_reference_host() and rh.have_ref_engine(), "needs /root/reference and oracle/_ref (make -f oracle/Makefile)"
    for w in what:
        {"tables": make_tables, "kat": make_kat, "runs": make_runs, "stats": make_stats, "dipole": make_dipole, "u4cross": make_u4cross, "stats32": make_stats32, "refhost": make_refhost}[w]()
