"""Parallel tempering across ranks (torchrun): per-replica (E0, M) cross NVLink through an NCCL allgather."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from mcsolver_b200 import pt
from mcsolver_b200.lattice import LatticeSpec
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
L = int(os.environ.get("L3", "64"))
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
n = 8 * world
T = 0.8 * 1.443 * (1.3 / 0.8) ** (np.arange(n) / (n - 1))
p = pt.ParallelTempering(spec, 3, T, precision=32, seed=3, rank=rank, world=world, device=local,
                         allgather=pt.torch_allgather(dev), allreduce_sum=pt.torch_allreduce_sum(dev))
t0 = time.time()
rows = p.run(200, 800, sweeps_per_swap=5)
dt = time.time() - t0
if rank == 0:
    print(json.dumps({"world": world, "L": L, "ladder": n, "wall_s": dt, "attempts_per_s": n * spec.nsite * 1000 / dt,
                      "swap_rates": np.round(p.swap_rates(), 3).tolist(), "e_over_kT": np.round(rows[:, 8], 4).tolist(),
                      "U4": np.round(rows[:, 10], 4).tolist()}))
p.close()
dist.barrier(); dist.destroy_process_group()
