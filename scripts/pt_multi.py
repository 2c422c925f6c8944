"""Parallel tempering across ranks (torchrun): per-replica (E0, M) cross NVLink through an NCCL allgather."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from mcsolver_b200 import pt
from tests import dist_util
from mcsolver_b200.lattice import LatticeSpec, add_dipole_stencil
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
L = int(os.environ.get("L3", "64"))
J = [-1, -1, -1] + [0] * 6
spec = LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J), (0, 0, (0, 1, 0), J), (0, 0, (0, 0, 1), J)])
DIP = float(os.environ.get("DIPOLE", "0"))        # dipole strength alpha (0 = exchange only), cut-off stencil r <= RCUT
RCUT = float(os.environ.get("RCUT", "2.0"))
if DIP:
    spec = add_dipole_stencil(spec, DIP, RCUT)
NTH, NSW, SPS = int(os.environ.get("NTHERMAL", "200")), int(os.environ.get("NSWEEP", "800")), int(os.environ.get("SWEEPS_PER_SWAP", "5"))
n = 8 * world
T = 0.8 * 1.443 * (1.3 / 0.8) ** (np.arange(n) / (n - 1))
p = pt.ParallelTempering(spec, 3, T, precision=32, seed=3, rank=rank, world=world, device=local,
                         allgather=dist_util.torch_allgather(dev), allreduce_sum=dist_util.torch_allreduce_sum(dev))
p.sys.timed_sweeps(1, with_measure=True)   # first launch of each specialised kernel (NVRTC compile or cubin load) stays untimed
p.sys.reset_measurements()
dist.barrier(); torch.cuda.synchronize()
t0 = time.time()
rows = p.run(NTH, NSW, sweeps_per_swap=SPS)
torch.cuda.synchronize(); dist.barrier()
dt = time.time() - t0
ncol = p.sys.num_colours()
if rank == 0:
    att = n * spec.nsite * (NTH + NSW) / dt
    z = 2 * len({(b[0], b[1], tuple(b[2])) for b in spec.bonds})   # coinciding templates are merged by the engine
    balg = (2 + min(ncol - 1, z)) * 12
    print(json.dumps({"world": world, "L": L, "ladder": n, "dipole_alpha": DIP, "rcut": RCUT if DIP else None, "links_per_site": z, "colours": ncol,
                      "sweeps": [NTH, NSW, SPS], "wall_s": dt, "attempts_per_s": att,
                      "roofline_frac_per_gpu": att / world * balg / 6547.5e9, "bytes_per_attempt": balg,
                      "what": "whole ladder incl. every-sweep measurement, swap steps (allgather of 2 doubles per replica over NCCL) and host relabelling; wall clock between barriers",
                      "swap_rates": np.round(p.swap_rates(), 3).tolist(), "e_over_kT": np.round(rows[:, 8], 4).tolist(),
                      "U4": np.round(rows[:, 10], 4).tolist()}))
p.close()
dist.barrier(); dist.destroy_process_group()
