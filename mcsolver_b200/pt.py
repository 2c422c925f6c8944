"""Parallel tempering over a (beta, H) ladder, sharded over the GPUs of a node.

New capability with no reference counterpart (the reference runs independent (T,H) points,
win.py:116-144).  Each rank holds a contiguous block of replicas resident on its GPU; every
`sweeps_per_swap` measured sweeps the ranks allgather three doubles per replica (E0, M, |M|) - the only
data-path exchange - evaluate the same deterministic swap decisions and relabel their own replicas.
Spin configurations never cross the interconnect.  Observables are accumulated per temperature label.

Two drivers:
  * in-library (default): `mcg_pt_run` enqueues sweeps, an `ncclAllGather` and a decide-and-relabel kernel on the
    system's stream, no host round trip per swap step, no PyTorch anywhere (libnccl is dlopen'ed by the library; the
    128-byte communicator id travels over a plain TCP socket, `comm_id()`);
  * host-driven (`allgather=` callable given): the caller's collective moves the state through the host - used by the
    CPU multi-process tests (gloo) and as the cross-check of the device-side decision kernel.
"""
import ctypes as C
import os
import socket
import sys
import time

import numpy as np

from . import _ffi, engine
from ._ffi import check, f64, i32, ptr
from .scan import shard

COMM_ID_BYTES = 128
ACC_MTMP, ACC_LASTE = 11, 23      # csrc/devmath.cuh: accumulator slots that are state, not sums


def _point_at_nccl():
    """Prefer the NCCL that ships with the Python environment (nvidia/nccl/lib) over the system one; the library
    dlopen()s MCG_NCCL_LIB first, then libnccl.so.2 by name."""
    if os.environ.get("MCG_NCCL_LIB"):
        return
    for d in sys.path:
        cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["MCG_NCCL_LIB"] = cand
            return


def comm_id(rank, world, addr=None, port=None, timeout=300.0):
    """The NCCL communicator id of a `world`-rank job: created by rank 0 (mcg_comm_unique_id) and handed to the other
    ranks over a TCP socket at (MASTER_ADDR, MASTER_PORT + 29) - the launcher's rendezvous variables, nothing else."""
    if world == 1:
        return None
    _point_at_nccl()
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(port or os.environ.get("MCG_PT_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 29))
    if rank == 0:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        check(_ffi.lib().mcg_comm_unique_id(buf, COMM_ID_BYTES))
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr if addr not in ("localhost",) else "127.0.0.1", port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                c, _a = srv.accept()
                c.sendall(buf.raw)
                c.close()
        finally:
            srv.close()
        return buf.raw
    t0 = time.time()
    while True:
        try:
            c = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() - t0 > timeout:
                raise
            time.sleep(0.05)
    data = b""
    while len(data) < COMM_ID_BYTES:
        chunk = c.recv(COMM_ID_BYTES - len(data))
        if not chunk:
            raise ConnectionError("communicator id truncated")
        data += chunk
    c.close()
    return data


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
                 allreduce_sum=None, tables=False, comm=None):
        """allgather/allreduce_sum given -> host-driven exchange through those callables; otherwise the library drives the
        loop itself (world > 1: `comm` = the id from comm_id(), fetched from the launcher's rendezvous if None)."""
        T = np.maximum(np.asarray(T, dtype=float).reshape(-1), 0.1)
        H = np.zeros_like(T) if H is None else np.asarray(H, dtype=float).reshape(-1)
        self.ladder = Ladder(1.0 / T, H, rank, world, seed)
        self.T = T
        lo, hi = self.ladder.lo, self.ladder.hi
        self.in_library = allgather is None
        self.allgather = allgather or local_allgather
        self.allreduce_sum = allreduce_sum or (lambda x: np.asarray(x))
        kw = dict(precision=precision, nReplica=hi - lo, beta=self.ladder.beta[lo:hi], field=H[lo:hi], seed=seed,
                  replica_offset=lo, device=device)
        if self.in_library and world > 1 and comm is None:
            comm = comm_id(rank, world)
        if tables:
            from .lattice import build_tables
            self.sys = engine.System.from_tables(build_tables(spec, 1.0, model), **kw)
        else:
            self.sys = engine.System.from_spec(spec, model, **kw)
        self.model, self.N = model, spec.nsite
        if self.in_library:
            b, h = f64(self.ladder.beta), f64(self.ladder.field)
            idbuf = C.create_string_buffer(comm, COMM_ID_BYTES) if comm is not None else None
            check(_ffi.lib().mcg_pt_setup(self.sys._h, rank, world, idbuf, self.ladder.n, ptr(b), ptr(h)))
        else:
            check(_ffi.lib().mcg_pt_configure(self.sys._h, self.ladder.n))
            self._apply(self.ladder.local_labels())
        self.sys.nLabelAll = self.ladder.n
        self.sys.init_spins(0.0)
        self.device_ms = 0.0

    def _apply(self, labels):
        lab = i32(labels)
        b, h = f64(self.ladder.beta[lab]), f64(self.ladder.field[lab])
        check(_ffi.lib().mcg_pt_set_labels(self.sys._h, ptr(lab), ptr(b), ptr(h)))

    def _state(self):
        st = np.zeros((self.sys.R, 2))
        check(_ffi.lib().mcg_pt_state(self.sys._h, ptr(st)))
        return st

    def run(self, nthermal, nsweep, sweeps_per_swap=1, want_results=True):
        """nthermal + nsweep sweeps, an exchange step after every `sweeps_per_swap` sweeps.
        Every sweep is measured (the exchange needs E, M); accumulators are cleared after thermalisation."""
        s = self.sys
        if self.in_library:
            ms = C.c_double(0)
            check(_ffi.lib().mcg_pt_run(s._h, int(nthermal), int(nsweep), int(sweeps_per_swap), C.byref(ms)))
            self.device_ms = ms.value
            return self.results() if want_results else None
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
        return self.results() if want_results else None

    def results(self):
        """Result rows per LABEL (temperature), accumulators summed over ranks.  COLLECTIVE: every rank calls it.
        autoCorr (slot 7 / Ising 3) is the lag-1 product of the series AT a temperature: exact in the in-library driver
        (the label's last |M| travels with it), undefined in the host-driven one when labels change rank."""
        n = self.ladder.n
        width = 10 if self.model == engine.ISING else 27
        rows = np.zeros((n, width))
        if self.in_library:
            check(_ffi.lib().mcg_pt_reduce(self.sys._h))
            for k in range(n):
                out = np.zeros(width)
                check(_ffi.lib().mcg_pt_results(self.sys._h, k, ptr(out), None))
                rows[k] = out
            return rows
        nacc = C.c_int(0)
        row0 = np.zeros(64)
        check(_ffi.lib().mcg_acc_get(self.sys._h, 0, ptr(row0), C.byref(nacc)))
        acc = np.zeros((n, nacc.value))
        for k in range(n):
            r = np.zeros(nacc.value)
            check(_ffi.lib().mcg_acc_get(self.sys._h, k, ptr(r), None))
            acc[k] = r
        mine = np.zeros(n, dtype=bool)
        mine[self.ladder.local_labels()] = True
        masked = acc.copy()
        masked[~mine, ACC_MTMP] = 0.0       # non-additive state slots count on the holder's rank only
        masked[~mine, ACC_LASTE] = 0.0
        tot = self.allreduce_sum(masked.reshape(-1)).reshape(n, -1)
        for k in range(n):
            check(_ffi.lib().mcg_acc_set(self.sys._h, k, ptr(f64(tot[k]))))
        for k in range(n):
            out = np.zeros(width)
            check(_ffi.lib().mcg_results(self.sys._h, k, ptr(out), None))
            rows[k] = out
        # restore this rank's own partial sums so that run() can be continued
        for k in range(n):
            check(_ffi.lib().mcg_acc_set(self.sys._h, k, ptr(f64(acc[k]))))
        return rows

    def swap_rates(self):
        if self.in_library:
            n = self.ladder.n
            att, acc = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
            check(_ffi.lib().mcg_pt_stats(self.sys._h, ptr(att), ptr(acc), None))
            return acc[:-1] / np.maximum(att[:-1], 1)
        a = np.maximum(self.ladder.attempts[:-1], 1)
        return self.ladder.accepts[:-1] / a

    def holders(self):
        """holder[k] = global replica currently carrying label k."""
        if self.in_library:
            h = np.zeros(self.ladder.n, dtype=np.int32)
            check(_ffi.lib().mcg_pt_stats(self.sys._h, None, None, ptr(h)))
            return h
        return self.ladder.holder.copy()

    def close(self):
        self.sys.close()
