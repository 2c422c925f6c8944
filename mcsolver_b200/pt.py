"""Parallel tempering over a (beta, H) ladder, sharded over the GPUs of a node.

New capability with no reference counterpart (the reference runs independent (T,H) points,
win.py:116-144).  Each rank holds a contiguous block of replicas resident on its GPU; every
`sweeps_per_swap` measured sweeps the ranks allgather two doubles per replica (E0, M) - the only
data-path exchange, an NCCL allgather over NVLink when the ranks are GPUs - evaluate the same
deterministic swap decisions (mcg_pt_decide) and relabel their own replicas.  Spin configurations
never cross the interconnect.  Observables are accumulated per temperature label.
"""
import ctypes as C

import numpy as np

from . import _ffi, engine
from ._ffi import check, f64, i32, ptr
from .scan import shard


def decide(beta, field, E0, M, holder, parity, seed, step):
    """One exchange step on the whole ladder (pure host, identical on every rank).
    holder[k] = global replica carrying label k; returns (new holder, accepted flags)."""
    b, h, e, m = f64(beta), f64(field), f64(E0), f64(M)
    hold = i32(holder).copy()
    acc = np.zeros(len(b), dtype=np.int32)
    check(_ffi.lib().mcg_pt_decide(len(b), ptr(b), ptr(h), ptr(e), ptr(m), ptr(hold), int(parity), int(seed), int(step), ptr(acc)))
    return hold, acc


def local_allgather(x):
    return np.asarray(x)


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


class Ladder:
    """Bookkeeping of which replica carries which label; independent of the engine (CPU-testable)."""

    def __init__(self, beta, field, rank=0, world=1, seed=1):
        self.beta, self.field = f64(beta).reshape(-1), f64(field).reshape(-1)
        self.n = self.beta.size
        if self.field.size != self.n:
            raise ValueError("beta and field must have one entry per ladder label")
        if self.n % world:
            raise ValueError("ladder size must be divisible by the number of ranks (equal allgather blocks)")
        self.rank, self.world, self.seed = rank, world, seed
        self.lo, self.hi = shard(self.n, rank, world)
        self.holder = np.arange(self.n, dtype=np.int32)     # label k is carried by global replica holder[k]
        self.step = 0
        self.attempts = np.zeros(self.n, dtype=np.int64)
        self.accepts = np.zeros(self.n, dtype=np.int64)

    def label_of_replica(self):
        lab = np.empty(self.n, dtype=np.int32)
        lab[self.holder] = np.arange(self.n, dtype=np.int32)
        return lab

    def local_labels(self):
        return self.label_of_replica()[self.lo:self.hi]

    def exchange(self, state_global):
        """state_global: [n,2] (E0, M) ordered by global replica index.  Returns local labels after the step."""
        st = np.asarray(state_global, dtype=float).reshape(self.n, 2)
        parity = self.step & 1
        self.holder, acc = decide(self.beta, self.field, st[:, 0], st[:, 1], self.holder, parity, self.seed, self.step)
        ks = np.arange(parity, self.n - 1, 2)
        self.attempts[ks] += 1
        self.accepts[ks] += acc[ks]
        self.step += 1
        return self.local_labels()


class ParallelTempering:
    def __init__(self, spec, model, T, H=None, precision=32, seed=1, rank=0, world=1, device=-1, allgather=None,
                 allreduce_sum=None, tables=False):
        T = np.maximum(np.asarray(T, dtype=float).reshape(-1), 0.1)
        H = np.zeros_like(T) if H is None else np.asarray(H, dtype=float).reshape(-1)
        self.ladder = Ladder(1.0 / T, H, rank, world, seed)
        self.T = T
        lo, hi = self.ladder.lo, self.ladder.hi
        self.allgather = allgather or local_allgather
        self.allreduce_sum = allreduce_sum or (lambda x: np.asarray(x))
        if world > 1 and allgather is None:
            raise ValueError("world > 1 needs an allgather callable (e.g. pt.torch_allgather(device))")
        kw = dict(precision=precision, nReplica=hi - lo, beta=self.ladder.beta[lo:hi], field=H[lo:hi], seed=seed,
                  replica_offset=lo, device=device)
        if tables:
            from .lattice import build_tables
            self.sys = engine.System.from_tables(build_tables(spec, 1.0, model), **kw)
        else:
            self.sys = engine.System.from_spec(spec, model, **kw)
        self.model, self.N = model, spec.nsite
        check(_ffi.lib().mcg_pt_configure(self.sys._h, self.ladder.n))
        self.sys.nLabelAll = self.ladder.n
        self._apply(self.ladder.local_labels())
        self.sys.init_spins(0.0)

    def _apply(self, labels):
        lab = i32(labels)
        b, h = f64(self.ladder.beta[lab]), f64(self.ladder.field[lab])
        check(_ffi.lib().mcg_pt_set_labels(self.sys._h, ptr(lab), ptr(b), ptr(h)))

    def _state(self):
        st = np.zeros((self.sys.R, 2))
        check(_ffi.lib().mcg_pt_state(self.sys._h, ptr(st)))
        return st

    def run(self, nthermal, nsweep, sweeps_per_swap=1, measure_thermal=False):
        """nthermal + nsweep sweeps, an exchange step after every `sweeps_per_swap` sweeps.
        Every sweep is measured (the exchange needs E, M); accumulators are cleared after thermalisation."""
        s = self.sys
        done = 0
        total = nthermal + nsweep
        cleared = nthermal == 0
        while done < total:
            limit = nthermal if done < nthermal else total
            n = min(sweeps_per_swap, limit - done)
            s.timed_sweeps(n, with_measure=True)
            done += n
            if not cleared and done >= nthermal:
                s.reset_measurements()      # clears the per-label sums; the last-sweep state (E0, M) is kept
                cleared = True
            st = self.allgather(self._state().reshape(-1)).reshape(-1, 2)
            self._apply(self.ladder.exchange(st))
        return self.results()

    def results(self):
        """Result rows per LABEL (temperature), accumulators summed over ranks."""
        n = self.ladder.n
        nacc = C.c_int(0)
        row0 = np.zeros(64)
        check(_ffi.lib().mcg_acc_get(self.sys._h, 0, ptr(row0), C.byref(nacc)))
        acc = np.zeros((n, nacc.value))
        for k in range(n):
            r = np.zeros(nacc.value)
            check(_ffi.lib().mcg_acc_get(self.sys._h, k, ptr(r), None))
            acc[k] = r
        tot = self.allreduce_sum(acc.reshape(-1)).reshape(n, -1)
        for k in range(n):
            check(_ffi.lib().mcg_acc_set(self.sys._h, k, ptr(f64(tot[k]))))
        width = 10 if self.model == engine.ISING else 27
        rows = np.zeros((n, width))
        for k in range(n):
            out = np.zeros(width)
            g = np.zeros(8)
            check(_ffi.lib().mcg_results(self.sys._h, k, ptr(out), None))
            rows[k] = out
        # restore this rank's own partial sums so that run() can be continued
        for k in range(n):
            check(_ffi.lib().mcg_acc_set(self.sys._h, k, ptr(f64(acc[k]))))
        return rows

    def swap_rates(self):
        a = np.maximum(self.ladder.attempts[:-1], 1)
        return self.ladder.accepts[:-1] / a

    def close(self):
        self.sys.close()
