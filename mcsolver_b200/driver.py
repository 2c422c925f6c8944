"""`loadMC(parameterfile)` on the GPU: the reference's headless entry (mcsolver/__init__.py:16-17 ->
win.py:40-159) with the whole (H,T) grid run as one replica batch instead of a process pool.

Writes the same files in the current directory: `./out` (one labelled line per point, mcMain.py:128-132 /
:261-265), `./result.txt` (win.py:152-156, 10 columns %15.6E) and `./spinDotSpin.txt` (mcMain.py:274-289).
Rows appear in grid order (the reference writes them in task-completion order).
"""
import os
import time

import numpy as np

from . import engine, paramfile, scan


def _write_out_line(f, model, T, h, r, extra=None):
    if model == engine.ISING:     # mcMain.py:126-132
        si, sj, sij, auto, E, E2, Er, E2r, U4, stot = r[:10]
        E, E2, Er, E2r = E * T, E2 * T ** 2, Er * T, E2r * T ** 2
        f.write("%.3f %.3f %.3f %.3f %.3f %.3f %.3f %.3f %.3f %.3f %.6f\n" % (T, h, si, sj, stot, sij, E, E2, Er, E2r, U4))
        return
    si, sj = r[0:3], r[3:6]          # mcMain.py:250-265
    sir, sjr = r[11:14], r[14:17]
    E, E2, Er, E2r = r[8] * T, r[9] * T ** 2, r[18] * T, r[19] * T ** 2
    f.write('T= %.6E h= %.6E <Siz>= %.6E <Sjz>= %.6E <Sz>= %.6E <Sih>= %.6E <Sjh>= %.6E <Sh>= %.6E <Si>= %.6E <Sj>= %.6E <SiSj>= %.6E '
            '<Si><Sj>= %.6E <E>= %.6E <E^2>= %.6E <U4>= %.6E <SiSj>r= %.6E <Si>r<Sj>r= %.6E <Er>= %.6E <E^2_r>= %.6E <Q>= %.6E\n' % (
                T, h, r[20], r[21], r[22], r[23], r[24], r[25], np.linalg.norm(si), np.linalg.norm(sj), r[6], np.dot(si, sj), E, E2, r[10],
                r[17], np.dot(sir, sjr), Er, E2r, r[26]))


def _gather_rows(workdir, tag, rank, world, payload, timeout=3600.0):
    """Result rows of all ranks on rank 0, through files in `workdir` (the product has no PyTorch and the grid points are
    independent: nothing else ever crosses between the ranks of a scan).  Other ranks return None."""
    mine = os.path.join(workdir, ".mcg_%s_rank%d_of%d.npz" % (tag, rank, world))
    np.savez(mine + ".tmp.npz", **{k: v for k, v in payload.items() if v is not None})
    os.replace(mine + ".tmp.npz", mine)
    if rank != 0:
        return None
    parts, t0 = [], time.time()
    for r in range(world):
        f = os.path.join(workdir, ".mcg_%s_rank%d_of%d.npz" % (tag, r, world))
        while not os.path.exists(f):
            if time.time() - t0 > timeout:
                raise TimeoutError("rank %d never delivered its rows (%s)" % (r, f))
            time.sleep(0.05)
        with np.load(f) as z:
            parts.append({k: z[k] for k in z.files})
        os.remove(f)
    out = {}
    for k in payload:
        have = [p[k] for p in parts if k in p]
        out[k] = np.concatenate(have) if have and len(have) == world else None
    return out


def loadMC(rpath, precision=None, seed=None, rank=None, world=None, device=None, workdir=".", table_limit=200000, quiet=False, dipole_rcut=2.0):
    """Run the simulation described by a reference parameter file.  Returns the result table
    (dict of columns, as written to result.txt).
    Several GPUs: start one process per GPU under any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK (torchrun does); each
    rank runs its contiguous share of the (H,T) grid - the reference's process-pool axis, win.py:90-91 - on GPU LOCAL_RANK, rank 0
    collects the rows and writes the files; the files are byte for byte those of a one-GPU run (the Philox streams follow the grid
    point, not the rank)."""
    t0 = time.time()
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    if device is None:
        device = int(os.environ["LOCAL_RANK"]) % max(1, engine.device_count()) if (world > 1 and "LOCAL_RANK" in os.environ) else -1
    p = paramfile.parse(rpath)
    if p.algorithm not in ("Metropolis", "Wolff"):
        raise ValueError("only Metropolis and Wolff algorithm is supported")
    model = p.model
    spec = p.spec()
    if abs(p.dipoleAlpha) > 1e-5:
        # mcMain.py:37-40 would call generateDipoleBondings (all pairs, open boundary; raises in the reference).
        # Here: periodic cut-off stencil of radius dipole_rcut (units of LMatrix), SURVEY 8 f2.
        from .lattice import add_dipole_stencil
        spec = add_dipole_stencil(spec, p.dipoleAlpha, dipole_rcut, ising=(model == engine.ISING))
    T, H = p.grid()
    Tf = np.maximum(T, 0.1)
    algo = engine.WOLFF if p.algorithm == "Wolff" else engine.METROPOLIS
    prec = engine.default_precision() if precision is None else precision
    sd = engine.default_seed() if seed is None else seed
    use_tables = spec.nsite <= table_limit     # per-site tables exactly as the reference builds them, while affordable
    block_spin = all(l == 1 or l % 2 == 0 for l in spec.L)   # structured path: `out` columns <..>r, <Er>, <E^2_r>
    ninterval = spec.nsite if p.ninterval <= 0 else p.ninterval
    idx, rows, frames = scan.run_points(spec, model, T, H, p.nthermal, p.nsweep, ninterval=ninterval, algorithm=algo, precision=prec,
                                        seed=sd, rank=rank, world=world, device=device, spin_frames=p.spinFrame, tables=use_tables,
                                        want_groups=True, block_spin=block_spin)
    rows, groups = rows
    if world > 1:
        import zlib
        tag = "%08x" % zlib.crc32(("%s|%d|%d|%d" % (os.path.abspath(rpath), p.nthermal, p.nsweep, len(T))).encode())
        allr = _gather_rows(workdir, tag, rank, world, dict(idx=idx, rows=rows, groups=groups, frames=frames))
        if rank != 0:
            return dict(T=Tf[idx], H=H[idx], rows=rows, **scan.observables(rows, Tf[idx], spec.nsite, model))
        idx, rows, groups, frames = allr["idx"], allr["rows"], allr["groups"], allr["frames"]
    obs = scan.observables(rows, Tf[idx], spec.nsite, model)
    if rank == 0:
        with open(os.path.join(workdir, "out"), "w") as f:
            f.write("#T #H\n")
            for k, i in enumerate(idx):
                _write_out_line(f, model, Tf[i], H[i], rows[k])
        with open(os.path.join(workdir, "spinDotSpin.txt"), "w") as f:
            f.write("#T #H\n")
            if model != engine.ISING and groups is not None and len(p.orbGroupList) > 0:
                for k, i in enumerate(idx):
                    f.write("%.3f %.3f " % (Tf[i], H[i]) + "".join("%.6f " % v for v in groups[k]) + "\n")
        with open(os.path.join(workdir, "result.txt"), "w") as f:
            f.write("#Temp          #Field         #<Si>          #<Sj>          #Susc          #Energy(K)     #Capacity(K/K) #Topo.Q        #U4            #Auto-corr.    \n")
            for k, i in enumerate(idx):
                f.write("%15.6E%15.6E%15.6E%15.6E%15.6E%15.6E%15.6E%15.6E%15.6E%15.6E\n" % (
                    Tf[i], H[i], obs["Si"][k], obs["Sj"][k], obs["Susc"][k], obs["Energy"][k], obs["Capacity"][k], obs["TopoQ"][k],
                    obs["U4"][k], obs["AutoCorr"][k]))
        if p.spinFrame > 0 and frames is not None:
            from .lattice import positions
            xyz = positions(spec)
            for k, i in enumerate(idx):
                for fr in range(p.spinFrame):
                    if model == engine.ISING:    # mcMain.py:291-297
                        name = "IsingSpinDistribution.T%.3f.H%.3f.%d.txt" % (Tf[i], H[i], fr)
                        np.savetxt(os.path.join(workdir, name), np.column_stack([xyz, frames[k, fr]]), fmt="%.6f %.6f %.6f %.3f",
                                   header="x       #y       #z       #spin", comments="#")
                    else:                        # mcMain.py:317-323
                        name = "OnSpinDistribution.T%.3f.H%.3f.%d.txt" % (Tf[i], H[i], fr)
                        np.savetxt(os.path.join(workdir, name), np.column_stack([xyz, frames[k, fr]]), fmt="%.6f",
                                   header="x       #y       #z       #spinx  #spiny  #spinz", comments="#")
    if not quiet and rank == 0:
        print("time elapsed %.3f s" % (time.time() - t0))
    out = dict(T=Tf[idx], H=H[idx], **obs)
    out["rows"] = rows
    return out
