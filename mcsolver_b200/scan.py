"""(T,H) grid scans on the GPU: the data-parallel axis of the reference, batched.

What it replaces: `win.py:75-144` builds one task per (H,T) point and farms them over a
`multiprocessing.Pool`, one forked process (and one full Python lattice object graph) per point.
Here every point is a replica of ONE resident system: all points of a GPU advance in the same
kernel launches (replica = a grid dimension), sharing the lattice tables; with several GPUs the
points are sharded over ranks with no data-path collective - only the final result
rows are gathered by the caller.

Couplings are passed UNSCALED; replica r uses beta_r = 1/max(T_r, 0.1) (mcMain.py:21) and field H_r.
"""
import numpy as np

from . import engine


def shard(n_points, rank, world):
    """Contiguous block partition of the grid points over ranks (PT neighbours stay local)."""
    base, rem = divmod(n_points, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---- created systems kept for the next job on the same lattice ----
# A scan is usually followed by another scan of the same lattice (next seed, refined grid, next field value): creation - device
# allocation, table build and upload, colouring search, loading the specialised kernels - then costs more than a short job's
# sweeps.  run_points() keeps the systems it created (per process, keyed by everything creation depends on, at most POOL_MAX of
# them) and recycles them through mcg_recycle, which restarts every RNG counter: results are bit for bit those of a fresh system
# (tests/test_gpu_pool.py).  MCG_POOL=0 or clear_pool() turn it off / give the memory back.
import os as _os

POOL_MAX = 2
_pool = {}          # key -> System (insertion order = age)


def clear_pool():
    for s in _pool.values():
        s.close()
    _pool.clear()


def _spec_key(spec):
    return (tuple(spec.L), tuple(np.asarray(spec.S, float).ravel()), tuple(np.asarray(spec.D, float).ravel()),
            tuple((b[0], b[1], tuple(b[2]), tuple(float(x) for x in b[3])) for b in spec.bonds), repr(spec.pair), repr(spec.circuits),
            repr(spec.groups), bool(spec.groupInSC))


def _acquire(key, make, beta, field, seed, lo):
    if _os.environ.get("MCG_POOL", "1") == "0":
        return make(), False
    s = _pool.pop(key, None)
    if s is not None:
        try:
            s.recycle(beta=beta, field=field, seed=seed, replica_offset=lo)
            return s, True
        except Exception:
            s.close()
    return make(), True


def _release(key, s, pooled):
    if not pooled:
        s.close()
        return
    _pool[key] = s
    while len(_pool) > POOL_MAX:
        _pool.pop(next(iter(_pool))).close()


def run_points(spec, model, T, H, nthermal, nsweep, ninterval=0, algorithm=engine.METROPOLIS, precision=32, seed=1,
               rank=0, world=1, device=-1, flunc=0.0, spin_frames=0, tables=False, want_groups=False, block_spin=False,
               info=None):
    """Run the points (T[i], H[i]) owned by `rank` and return (indices, results[n,27|10], frames).

    spec: LatticeSpec (bond templates + supercell).  ninterval<=0 means N (mcMain.py:145).
    tables=True forces the table-driven engine (always full tuples); default is the structured path, which fills the
    block-spin slots 11-19 (Ising 6-7) only when block_spin=True (one more read of the configuration per sweep).
    Result rows have the reference's tuple layout with E, E2 still in beta units (caller rescales
    exactly as mcMain.py:251 does).  info: optional dict, filled with the launch counters of the job."""
    T = np.atleast_1d(np.asarray(T, dtype=float))
    H = np.atleast_1d(np.asarray(H, dtype=float))
    if T.shape != H.shape:
        raise ValueError("T and H must list the grid points pairwise")
    lo, hi = shard(T.size, rank, world)
    idx = np.arange(lo, hi)
    if idx.size == 0:
        empty = np.zeros((0, 10 if model == engine.ISING else 27))
        return idx, ((empty, None) if want_groups else empty), None
    Tl = np.maximum(T[lo:hi], 0.1)
    beta, field = 1.0 / Tl, H[lo:hi]
    N = spec.nsite
    nint = N if ninterval <= 0 else int(ninterval)
    key = (_spec_key(spec), int(model), int(precision), int(idx.size), int(device), bool(tables), bool(block_spin))

    def make():
        if tables:
            from .lattice import build_tables
            t = build_tables(spec, 1.0, model)     # unscaled tables, beta per replica
            return engine.System.from_tables(t, precision=precision, nReplica=idx.size, beta=beta, field=field, seed=seed,
                                             replica_offset=lo, device=device)
        return engine.System.from_spec(spec, model, precision=precision, nReplica=idx.size, beta=beta, field=field,
                                       seed=seed, replica_offset=lo, device=device, block_spin=block_spin)

    s, pooled = _acquire(key, make, beta, field, seed, lo)
    try:
        s.init_spins(flunc)
        frames = s.run(algorithm, nthermal, nsweep, nint, spinFrame=spin_frames)
        res = [s.results(r) for r in range(idx.size)]
        out = np.stack([r[0] for r in res])
        groups = np.stack([r[1] for r in res]) if (model != engine.ISING and res[0][1] is not None and res[0][1].size) else None
        if info is not None:
            info.update(launches=s.launch_count(), jit_launches=s.jit_launch_count(), colours=s.num_colours())
    except BaseException:
        s.close()
        raise
    _release(key, s, pooled)
    return idx, ((out, groups) if want_groups else out), frames


def observables(rows, T, N, model):
    """Post-processing of win.py:133-143 / mcMain.py:250-257 on result rows (beta units -> K)."""
    rows = np.asarray(rows, dtype=float)
    T = np.maximum(np.asarray(T, dtype=float), 0.1)
    if model == engine.ISING:
        si, sj, sij, auto, E, E2, U4 = rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 4] * T, rows[:, 5] * T ** 2, rows[:, 8]
        return dict(Si=si, Sj=sj, Susc=(sij - si * sj) / T, Energy=E, Capacity=(E2 - E * E) / T ** 2 * N, TopoQ=np.zeros_like(E),
                    U4=U4, AutoCorr=auto)
    si, sj = rows[:, 0:3], rows[:, 3:6]
    E, E2 = rows[:, 8] * T, rows[:, 9] * T ** 2
    return dict(Si=np.linalg.norm(si, axis=1), Sj=np.linalg.norm(sj, axis=1),
                Susc=(rows[:, 6] - np.sum(si * sj, axis=1)) / T, Energy=E, Capacity=(E2 - E * E) / T ** 2 * N,
                TopoQ=rows[:, 26], U4=rows[:, 10], AutoCorr=rows[:, 7])


def run_field_sweep(spec, model, T, H_path, nthermal, nsweep, ninterval=0, precision=32, seed=1, device=-1, tables=False):
    """True field hysteresis (opt-in extension): unlike the reference - whose "H scan" runs every
    field value as an independent simulation from the polarised state (win.py:116-119) - the configuration is
    carried from one field value of `H_path` to the next.  One replica per temperature in `T`; at every field
    value: nthermal intervals of relaxation, then nsweep measured sweeps.  Returns rows[len(H_path), len(T), 27|10]."""
    T = np.maximum(np.atleast_1d(np.asarray(T, dtype=float)), 0.1)
    H_path = np.atleast_1d(np.asarray(H_path, dtype=float))
    N = spec.nsite
    nint = N if ninterval <= 0 else int(ninterval)
    kw = dict(precision=precision, nReplica=T.size, beta=1.0 / T, field=np.full(T.size, H_path[0]), seed=seed, device=device)
    if tables:
        from .lattice import build_tables
        sysm = engine.System.from_tables(build_tables(spec, 1.0, model), **kw)
    else:
        sysm = engine.System.from_spec(spec, model, **kw)
    out = []
    with sysm as s:
        s.init_spins(0.0)
        for h in H_path:
            s.set_params(field=np.full(T.size, h))
            s.reset_measurements()
            s.run(engine.METROPOLIS, nthermal, nsweep, nint)
            out.append(np.stack([s.results(r)[0] for r in range(T.size)]))
    return np.stack(out)
