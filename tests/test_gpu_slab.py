"""Slab decomposition of one lattice along its first axis (mcg_create_lattice_slab; SURVEY 8e, optional in the reference's terms -
it has no decomposition, README.md:99).  The slabs' Philox counters are keyed by the global site ids and their ghost planes are
refreshed after every colour pass, so together they must perform the undivided lattice's trajectory bit for bit, and every rank
must accumulate the whole lattice's observables."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.specs import spec_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [("cubic", (16, 8, 8), 3, 0.0, 64), ("cubic", (16, 8, 16), 3, 0.3, 32), ("cubic", (8, 8, 8), 2, 0.0, 64), ("aniso", (8, 6, 8), 3, 0.2, 64),
         ("cubic", (32, 32, 64), 3, 0.0, 32), ("cubic", (16, 16, 16), 1, 0.1, 32)]
IDS = ["%s-%s-m%d-fp%d" % (c[0], "x".join(map(str, c[1])), c[2], c[4]) for c in CASES]


def _whole(spec, model, prec, T, H, flunc, nth, nsw):
    from mcsolver_b200 import engine
    R = len(T)
    with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1 / T, field=H, seed=5) as s:
        s.init_spins(flunc)
        s.run(0, nth, nsw, spec.nsite)
        return [s.get_spins(r) for r in range(R)], np.stack([s.results(r)[0] for r in range(R)]), [s.counters(r) for r in range(R)], s.energy(0)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_one_rank_slab_with_ghost_planes_reproduces_the_whole_lattice(case, monkeypatch):
    """world = 1: the slab owns every plane, its ghosts are its own periodic images - the row range, the ghost refresh after
    every colour pass, the global-id Philox counters and the folded sums are all in play on one GPU.  JIT on: the specialised
    kernels (literal row range and x offset) are the ones that run at size."""
    from mcsolver_b200 import engine
    name, L, model, h, prec = case
    monkeypatch.setenv("MCG_JIT", "1")
    spec = spec_of(name, L, circuits=[], pair=(0, 0, (0, 0, 0))) if name == "aniso" else spec_of(name, L)
    T = np.array([0.8, 1.5, 3.0]) * (1.0 if model != 1 else 3.0)
    H = np.full(3, h)
    flunc = 0.0 if model == 1 else 0.6
    sp, rows, cnt, E = _whole(spec, model, prec, T, H, flunc, 3, 6)
    with engine.System.from_spec_slab(spec, model, 0, 1, precision=prec, nReplica=3, beta=1 / T, field=H, seed=5) as s:
        assert s.slab["nx"] == L[0] and s.slab["x0"] == 0
        s.init_spins(flunc)
        s.run(0, 3, 6, spec.nsite)
        for r in range(3):
            assert np.array_equal(s.own_spins(r), sp[r]), r
            assert s.counters(r) == cnt[r]
        got = np.stack([s.results(r)[0] for r in range(3)])
        assert np.max(np.abs(got - rows) / np.maximum(1.0, np.abs(rows))) < (1e-11 if prec == 64 else 2e-6)
        assert abs(s.energy(0) - E) <= (1e-11 if prec == 64 else 2e-6) * abs(E)
        if s.rng_layout()[1] > 1:
            assert s.jit_launch_count() > 0


@pytest.mark.parametrize("prec", [32, 64])
def test_one_rank_slab_of_the_dipole_stencil_lattice(prec, monkeypatch):
    """BASELINE config 5's Hamiltonian (sc + dipole stencil r <= 2: 32 full-tensor links, period 4, 16 colours) as a slab: the ghost
    is a whole period (4 planes) wide, fp32 runs the asynchronous link pipeline (k_struct_async) over the slab's row range."""
    from mcsolver_b200 import engine
    from mcsolver_b200.lattice import add_dipole_stencil
    spec = add_dipole_stencil(spec_of("cubic", (16, 8, 16)), 0.1, 2.0)
    T, H = np.array([0.9, 1.6]), np.array([0.0, 0.2])
    sp, rows, cnt, E = _whole(spec, 3, prec, T, H, 0.5, 2, 4)
    with engine.System.from_spec_slab(spec, 3, 0, 1, precision=prec, nReplica=2, beta=1 / T, field=H, seed=5) as s:
        assert s.slab["ghost"] == 4 and s.num_colours() == 16
        s.init_spins(0.5)
        s.run(0, 2, 4, spec.nsite)
        for r in range(2):
            assert np.array_equal(s.own_spins(r), sp[r]), r
            assert s.counters(r) == cnt[r]
        got = np.stack([s.results(r)[0] for r in range(2)])
        assert np.max(np.abs(got - rows) / np.maximum(1.0, np.abs(rows))) < (1e-11 if prec == 64 else 2e-6)


def test_slab_refuses_what_is_not_decomposed():
    from mcsolver_b200 import engine
    with pytest.raises(engine.McgError):
        engine.System.from_spec_slab(spec_of("square", (16, 16, 1)), 2, 0, 1)            # two-dimensional supercell
    with pytest.raises(engine.McgError):
        engine.System.from_spec_slab(spec_of("cubic", (6, 8, 8)), 3, 0, 4)               # 6 planes do not split over 4 ranks
    with engine.System.from_spec_slab(spec_of("cubic", (8, 8, 8)), 3, 0, 1, precision=32) as s:
        s.init_spins(0.0)
        with pytest.raises(engine.McgError):
            s.wolff_steps(1)


WORKER = r'''
import json, os, sys
sys.path.insert(0, os.environ["MCG_ROOT"])
import numpy as np
from mcsolver_b200 import engine, pt
from tests.specs import spec_of
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
prec = int(os.environ["MCG_PREC"])
spec = spec_of("cubic", (32, 16, 32))
T = np.array([0.9, 1.44, 2.5]); H = np.array([0.0, 0.1, 0.0])
cid = pt.comm_id(rank, world)
with engine.System.from_spec_slab(spec, 3, rank, world, comm_id=cid, precision=prec, nReplica=3, beta=1 / T, field=H, seed=5, device=rank) as s:
    s.init_spins(0.5)
    s.run(0, 4, 8, spec.nsite)
    out = dict(info=s.slab, rows=[s.results(r)[0].tolist() for r in range(3)], counters=[s.counters(r) for r in range(3)],
               E=s.energy(1), torch_loaded="torch" in sys.modules)
    np.save(os.environ["MCG_OUT"] + ".%d.npy" % rank, np.stack([s.own_spins(r) for r in range(3)]))
json.dump(out, open(os.environ["MCG_OUT"] + ".%d.json" % rank, "w"))
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("prec", [64, 32])
def test_slabs_on_several_gpus_reproduce_the_whole_lattice(prec, world, tmp_path):
    """One process per GPU, halo planes over NCCL send/recv after every colour pass (two ranks: both neighbours are the same
    rank; four: distinct left and right neighbours), raw sums all-reduced: the slabs, put together, are the single-GPU
    configuration bit for bit; every rank reports the whole lattice's observables."""
    from mcsolver_b200 import engine
    if engine.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    spec = spec_of("cubic", (32, 16, 32))
    T = np.array([0.9, 1.44, 2.5]); H = np.array([0.0, 0.1, 0.0])
    sp, rows, cnt, _ = _whole(spec, 3, prec, T, H, 0.5, 4, 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "out")
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, MCG_ROOT=ROOT, MCG_OUT=out, MCG_PREC=str(prec), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o[-3000:]
    res = [json.load(open(out + ".%d.json" % r)) for r in range(world)]
    parts = [np.load(out + ".%d.npy" % r) for r in range(world)]
    assert not res[0]["torch_loaded"]
    assert [res[r]["info"]["x0"] for r in range(world)] == [r * 32 // world for r in range(world)]
    for r in range(3):
        assert np.array_equal(np.concatenate([p[r] for p in parts]), sp[r]), r
        assert tuple(np.sum([res[k]["counters"][r] for k in range(world)], axis=0)) == cnt[r]
    assert all(res[k]["rows"] == res[0]["rows"] and res[k]["E"] == res[0]["E"] for k in range(world))
    got = np.array(res[0]["rows"])
    assert np.max(np.abs(got - rows) / np.maximum(1.0, np.abs(rows))) < (1e-11 if prec == 64 else 2e-6)
