"""CPU tests of the N>1 host logic with a world_size-2 gloo group: grid sharding, the parallel-
tempering ladder bookkeeping and the determinism of the swap decisions across ranks.  No GPU: the
engine is replaced by a synthetic energy model; mcg_pt_decide is pure host code in the C-ABI library."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["MCG_ROOT"])
import torch.distributed as dist
from mcsolver_b200 import pt, scan
from tests import dist_util
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 16
T = np.linspace(1.0, 2.5, n)
lad = pt.Ladder(1.0 / T, np.zeros(n), rank=rank, world=world, seed=7)
ag = dist_util.torch_allgather(None)
rng = np.random.RandomState(100 + rank)
lo, hi = lad.lo, lad.hi
hist = []
for step in range(40):
    labels = lad.local_labels()
    # synthetic replica state: energy fluctuating around -N*T_label-dependent mean (overlapping distributions)
    e_local = -100.0 / T[labels] + 3.0 * rng.randn(hi - lo)
    st = np.stack([e_local, np.zeros(hi - lo)], axis=1)
    lad.exchange(ag(st.reshape(-1)))
    hist.append(lad.holder.tolist())
out = dict(rank=rank, lo=lo, hi=hi, holder=lad.holder.tolist(), hist=hist, acc=lad.accepts.tolist(), att=lad.attempts.tolist(),
           shard=[list(scan.shard(21, r, world)) for r in range(world)])
json.dump(out, open(os.environ["MCG_OUT"] + ".%d" % rank, "w"))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_ladder_is_consistent_across_two_gloo_ranks(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "out.json")
    env = dict(os.environ, MCG_ROOT=ROOT, MCG_OUT=out, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    subprocess.run(cmd, env=env, check=True, timeout=300, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    r0, r1 = (json.load(open(out + ".%d" % r)) for r in range(2))
    # both ranks drew the same decisions from the same allgathered state
    assert r0["hist"] == r1["hist"] and r0["holder"] == r1["holder"]
    assert sorted(r0["holder"]) == list(range(16))                     # labels stay a permutation of replicas
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 8, 8, 16)   # contiguous blocks
    assert sum(r0["acc"]) > 0 and sum(r0["acc"]) < sum(r0["att"])      # some, not all, swaps accepted
    assert r0["shard"] == [[0, 11], [11, 21]]                          # ragged split of 21 grid points
    # labels actually travel between the two ranks' blocks
    assert any(h != list(range(16)) for h in r0["hist"])
    assert any(set(h[:8]) != set(range(8)) for h in r0["hist"])


def test_decide_detailed_balance_limits():
    from mcsolver_b200 import pt
    beta = np.array([1.0, 0.5])
    # replica 0 (cold label) has HIGHER energy than replica 1: swapping lowers the action -> always accepted
    hold, acc = pt.decide(beta, [0, 0], [5.0, -5.0], [0, 0], [0, 1], 0, 1, 0)
    assert hold.tolist() == [1, 0] and acc[0] == 1
    # the reverse costs Delta = (b0-b1)*(E1-E0) = 0.5*40 = 20: accepted with prob e^-20 -> essentially never
    n_acc = sum(pt.decide(beta, [0, 0], [-20.0, 20.0], [0, 0], [0, 1], 0, 1, s)[1][0] for s in range(200))
    assert n_acc == 0
    # equal energies: always accepted; parity 1 leaves pair (0,1) untouched
    assert pt.decide(beta, [0, 0], [1.0, 1.0], [0, 0], [0, 1], 0, 1, 3)[1][0] == 1
    assert pt.decide(beta, [0, 0], [5.0, -5.0], [0, 0], [0, 1], 1, 1, 0)[0].tolist() == [0, 1]
    # acceptance frequency matches exp(-Delta) for a moderate Delta = 0.5*(1.0) = 0.5 -> 0.6065
    n = 4000
    f = sum(pt.decide(beta, [0, 0], [0.0, 1.0], [0, 0], [0, 1], 0, 11, s)[1][0] for s in range(n)) / n
    assert abs(f - np.exp(-0.5)) < 4 * np.sqrt(0.6065 * 0.3935 / n)


ID_WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, os.environ["MCG_ROOT"])
from mcsolver_b200 import pt
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
cid = pt.comm_id(rank, world)
assert len(cid) == pt.COMM_ID_BYTES and "torch" not in sys.modules
open(os.environ["MCG_OUT"] + ".%d" % rank, "w").write(hashlib.sha256(cid).hexdigest())
'''


def test_communicator_id_reaches_every_rank_without_torch(tmp_path):
    """The NCCL communicator id of the in-library tempering driver: made by rank 0 (ncclGetUniqueId through the C ABI,
    works without a GPU) and passed over a TCP socket at the launcher's MASTER_ADDR / MASTER_PORT + 29."""
    from mcsolver_b200 import _ffi
    import ctypes
    buf = ctypes.create_string_buffer(128)
    if _ffi.lib().mcg_comm_unique_id(buf, 128) != 0:
        import pytest
        pytest.skip("libnccl not loadable here: " + _ffi.lib().mcg_last_error().decode())
    script = tmp_path / "idworker.py"
    script.write_text(ID_WORKER)
    out = str(tmp_path / "id")
    port = _free_port()
    world = 3
    procs = []
    for rank in reversed(range(world)):      # clients first: they must retry until rank 0 listens
        env = dict(os.environ, MCG_ROOT=ROOT, MCG_OUT=out, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=120)
        assert p.returncode == 0, o[-2000:]
    ids = [open(out + ".%d" % r).read() for r in range(world)]
    assert len(set(ids)) == 1 and len(ids[0]) == 64


def test_slab_plan_tiles_the_first_axis_for_every_world_size():
    """Host half of the slab decomposition (mcg_slab_plan, no GPU): for every rank count the slabs tile [0, Lx) without gap or
    overlap, the ghost width is the colouring period of the WHOLE lattice (the same on every rank), and lattices that cannot be
    cut into whole periods are refused with a message instead of being cut wrongly."""
    from mcsolver_b200 import engine
    from tests.specs import spec_of
    for name, L, model, prec in (("cubic", (64, 32, 32), 3, 32), ("cubic", (48, 8, 8), 2, 64), ("aniso", (16, 6, 8), 3, 64)):
        spec = spec_of(name, L, circuits=[], pair=(0, 0, (0, 0, 0)))
        period = engine.slab_plan(spec, model, 0, 1, precision=prec)["ghost"]
        for world in (1, 2, 4, 8):
            if L[0] % world or (L[0] // world) % period or L[0] // world < 2 * period:
                with pytest.raises(engine.McgError):
                    engine.slab_plan(spec, model, 0, world, precision=prec)
                continue
            plans = [engine.slab_plan(spec, model, r, world, precision=prec) for r in range(world)]
            assert [p["x0"] for p in plans] == [r * L[0] // world for r in range(world)]
            assert sum(p["nx"] for p in plans) == L[0] and len({p["ghost"] for p in plans}) == 1
            assert all(p["nx"] % p["ghost"] == 0 and p["nx"] >= 2 * p["ghost"] for p in plans)
    for bad, world in (((6, 8, 8), 4), ((16, 16, 1), 2), ((10, 8, 8), 4)):
        with pytest.raises(engine.McgError):
            engine.slab_plan(spec_of("cubic" if bad[2] > 1 else "square", bad), 3, 0, world)
    with pytest.raises(engine.McgError):      # topological charge is not decomposed
        engine.slab_plan(spec_of("aniso", (16, 6, 8)), 3, 0, 2)


def test_loadmc_row_gather_between_ranks(tmp_path):
    """driver._gather_rows: the only thing that crosses between the ranks of a sharded loadMC scan - rank 0 receives every rank's
    block of result rows in rank (= grid) order, optional parts that no rank produced stay None, the exchange files are removed."""
    import threading
    from mcsolver_b200 import driver
    world, got = 3, {}

    def work(rank):
        rows = np.full((2, 4), float(rank))
        payload = dict(idx=np.arange(2 * rank, 2 * rank + 2), rows=rows, groups=None, frames=None)
        got[rank] = driver._gather_rows(str(tmp_path), "t", rank, world, payload, timeout=60.0)

    th = [threading.Thread(target=work, args=(r,)) for r in (2, 0, 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert got[1] is None and got[2] is None
    assert got[0]["idx"].tolist() == [0, 1, 2, 3, 4, 5] and got[0]["rows"][:, 0].tolist() == [0, 0, 1, 1, 2, 2]
    assert got[0]["groups"] is None and got[0]["frames"] is None
    assert list(tmp_path.iterdir()) == []
