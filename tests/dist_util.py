"""Collectives for the host-driven mode of mcsolver_b200.pt (`allgather=` / `allreduce_sum=` callables) over torch.distributed: gloo
on CPU (tests/test_multirank_cpu.py), NCCL when a CUDA device is given (scripts/pt_multi.py).  Test infrastructure: the product's own
multi-GPU path talks to NCCL inside the library and never imports PyTorch."""
import numpy as np


def torch_allgather(device=None):
    """allgather over torch.distributed: NCCL (NVLink/NVSwitch) when `device` is a CUDA device, gloo on CPU."""
    import torch
    import torch.distributed as dist

    def ag(x):
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        if device is not None:
            t = t.to(device)
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return torch.cat(out).cpu().numpy()
    return ag


def torch_allreduce_sum(device=None):
    import torch
    import torch.distributed as dist

    def ar(x):
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    return ar
