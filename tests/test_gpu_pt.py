"""Parallel tempering driven by the library (csrc/pt.cu: mcg_pt_setup / mcg_pt_run / mcg_pt_reduce): the device-side
decide-and-relabel kernel takes the same decisions as the host restatement (mcg_pt_decide), and a ladder sharded over two
GPUs - replica energies exchanged by ncclAllGather inside the library, no PyTorch - gives what one GPU gives."""
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


def test_in_library_tempering_equals_host_driven_exchange():
    """Same ladder, same seed: the in-library loop (pack kernel -> gather -> k_pt_decide, nothing synchronises with the host)
    and the host-driven loop (mcg_pt_state -> mcg_pt_decide -> mcg_pt_set_labels) are the same Markov chain."""
    from mcsolver_b200 import pt
    spec = spec_of("cubic", (8, 8, 8))
    T = np.linspace(1.2, 1.9, 8)
    H = np.linspace(0.0, 0.07, 8)
    a = pt.ParallelTempering(spec, 3, T, H, precision=64, seed=5)
    assert a.in_library
    ra = a.run(40, 160, sweeps_per_swap=2)
    b = pt.ParallelTempering(spec, 3, T, H, precision=64, seed=5, allgather=pt.local_allgather)
    assert not b.in_library
    rb = b.run(40, 160, sweeps_per_swap=2)
    assert np.array_equal(a.holders(), b.holders())
    assert np.allclose(a.swap_rates(), b.swap_rates(), atol=0) and a.swap_rates().max() > 0.05
    assert any(a.holders() != np.arange(8))                                 # labels did move
    assert np.max(np.abs(ra - rb) / np.maximum(1.0, np.abs(rb))) < 1e-12    # every slot incl. autoCorr (one rank: same series)
    for r in range(8):
        assert np.array_equal(a.sys.get_spins(r), b.sys.get_spins(r))
    a.close()
    b.close()


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.environ["MCG_ROOT"])
from mcsolver_b200 import pt
from tests.specs import spec_of
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
assert "torch" not in sys.modules
spec = spec_of("cubic", (8, 8, 16))
n = 8
T = np.linspace(1.2, 1.9, n)
p = pt.ParallelTempering(spec, 3, T, precision=64, seed=5, rank=rank, world=world, device=rank if world > 1 else 0)
rows = p.run(40, 200, sweeps_per_swap=2)
out = dict(rows=rows.tolist(), holders=p.holders().tolist(), rates=p.swap_rates().tolist(),
           spins0=p.sys.get_spins(0)[:64].tolist(), torch_loaded="torch" in sys.modules)
json.dump(out, open(os.environ["MCG_OUT"] + ".%d.%d" % (world, rank), "w"))
p.close()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_ladder_sharded_over_two_gpus_by_nccl_equals_one_gpu(tmp_path):
    from mcsolver_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "out.json")
    res = {}
    for world in (1, 2):
        port = _free_port()
        procs = []
        for rank in range(world):
            env = dict(os.environ, MCG_ROOT=ROOT, MCG_OUT=out, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        for p in procs:
            o, _ = p.communicate(timeout=600)
            assert p.returncode == 0, o[-3000:]
        res[world] = [json.load(open(out + ".%d.%d" % (world, r))) for r in range(world)]
    one, (a, b) = res[1][0], res[2]
    assert not a["torch_loaded"] and not b["torch_loaded"]                 # the product's multi-GPU path is PyTorch-free
    assert a["holders"] == b["holders"] == one["holders"]                  # same decisions on every rank and on one GPU
    assert a["rates"] == one["rates"] and max(a["rates"]) > 0.05
    assert a["spins0"] == one["spins0"]                                    # replica 0: same trajectory (GPU-count independent streams)
    ra, rb, r1 = np.array(a["rows"]), np.array(b["rows"]), np.array(one["rows"])
    assert np.array_equal(ra, rb)                                          # allreduce: both ranks hold the same sums
    assert np.max(np.abs(ra - r1) / np.maximum(1.0, np.abs(r1))) < 1e-11   # incl. autoCorr: |M| travels with the label
