#!/usr/bin/env python
"""bench.py - the headline measurement (BASELINE.json: attempted Metropolis spin updates / s).

Workload (config.workload): 3D Heisenberg, simple cubic 256^3 (N = 16 777 216 spins), J = diag(-1,-1,-1)
on [100],[010],[001], initial state polarised (S,0,0) as in the reference, temperature scan with 8
replicas per GPU (the 64-point ladder of BASELINE configs[4] sharded over 8 GPUs).  One "step" = SWEEPS
colour-class Metropolis sweeps over all replicas of the rank, every sweep followed by the reference's
per-sweep measurement (fused into the colour passes).  Beside the headline the same line carries, at every N:
"configs" (C1-C5 and C5 + dipole at their named sizes, replica-sharded the same way) and "pt" (the C5 ladder with
the dipole stencil as parallel tempering: ncclAllGather inside the library at every exchange step).

  python bench.py --gpus N --steps K --warmup W              # our engine (torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...    # the reference's own C engine on host cores

JSON line keys follow the driver contract; see DESIGN.md "Measurement" for the byte model.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

J_ISO = [-1.0, -1.0, -1.0] + [0.0] * 6
TC = 1.443  # Heisenberg sc, |J| = 1


def cubic_spec(L):
    from mcsolver_b200.lattice import LatticeSpec
    return LatticeSpec(L=(L, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J_ISO), (0, 0, (0, 1, 0), J_ISO), (0, 0, (0, 0, 1), J_ISO)])


def ladder(n_total):
    """geometric temperature ladder in [0.8 Tc, 1.3 Tc]"""
    return 0.8 * TC * (1.3 / 0.8) ** (np.arange(n_total) / max(1, n_total - 1))


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own compiled C engine (oracle/_ref) on host cores
# ---------------------------------------------------------------------------------------------
def _ref_worker(job):
    L, T, nthermal, nsweep, seed = job
    from oracle import refharness as rh
    from mcsolver_b200.lattice import build_tables
    t = build_tables(cubic_spec(L), T, 3)
    args = t.on_args(0, nthermal, nsweep, t.N, 0.0, 0.0, 0)
    t0 = time.time()
    out = rh.run_ref_engine(3, args, seed=seed)
    dt = time.time() - t0
    return t.N * (nthermal + nsweep), dt, out[8] * T


def _quiet_stdout():
    # the reference's PyInit_* printf("Initializing ...") goes to fd 1 of the worker: bench.py's stdout is ONE JSON line
    fd = os.open(os.devnull, os.O_WRONLY)
    os.dup2(fd, 1)
    os.close(fd)


def reference_step(nproc, L, nthermal, nsweep):
    """One bounded sample: nproc independent temperature points, one process each (the reference's
    own parallel axis, win.py:90-91).  Returns (attempts, wall seconds of the engine calls)."""
    import multiprocessing as mp
    Ts = ladder(nproc)
    jobs = [(L, float(Ts[i]), nthermal, nsweep, i + 1) for i in range(nproc)]
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(processes=nproc, initializer=_quiet_stdout) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.time() - t0
    attempts = sum(r[0] for r in res)
    engine_wall = max(r[1] for r in res)
    return attempts, engine_wall, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ref_kind():
    from oracle import refharness as rh
    return "reference" if rh.have_ref_engine() else "port"


def run_reference_arm(a, rank, world, emit):
    if rank != 0:
        return
    cores = min(host_cores(), 128)
    L, nth, nsw = a.ref_L, 2, a.ref_sweeps
    # the engine runs in forked pool workers; map it in this process too, so that whoever lists the shared objects this arm loaded
    # sees oracle/_ref/heisenberglib*.so and not an empty list (its PyInit prints to fd 1: silenced, this arm's stdout is ONE line)
    try:
        from oracle import refharness as rh
        if rh.have_ref_engine():
            keep = os.dup(1)
            _quiet_stdout()
            try:
                rh.load_ref_engine("heisenberglib")
            finally:
                sys.stdout.flush()
                os.dup2(keep, 1)
                os.close(keep)
    except Exception:
        pass
    for _ in range(max(0, min(a.warmup, 1))):
        reference_step(cores, L, 1, 2)
    att, eng, t = 0, 0.0, 0.0
    t0 = time.time()
    for _ in range(a.steps):
        x, e, w = reference_step(cores, L, nth, nsw)
        att += x
        eng += e
        t += w
    wall = time.time() - t0
    val = att / eng
    sample = "sc %d^3 Heisenberg, %d temperature points (1 process each), %d+%d sweeps per point per step, engine-call time only" % (L, cores, nth, nsw)
    # second sample at the largest size a reference process builds and runs inside a minute: the rate hardly depends on the
    # lattice once it leaves the cache (the engine chases pointers through 264-byte structs either way)
    big = None
    try:
        xb, eb, _ = reference_step(cores, a.ref_L2, 1, 6)
        big = {"value": xb / eb, "unit": "attempts/s", "sample": "sc %d^3, %d points x (1+6) sweeps, engine-call time %.1f s" % (a.ref_L2, cores, eb)}
    except Exception as e:      # memory-starved host: the headline sample stands alone
        big = {"error": str(e)[:200]}
    cfg = workload_config(a, world)
    # the metric (a rate) is the one of the GPU arm's workload; the reference cannot BUILD that workload (Python object graph:
    # 80 us + 3.3 kB per orbital), so the lattice it actually runs is named here
    cfg["reference_arm_runs"] = {"lattice": "simple cubic %d^3" % L, "spins": L ** 3, "points": cores, "sweeps_per_point_per_step": nth + nsw,
                                 "precision": "fp64 (the reference has no other)", "inputs": "mcsolver_b200.lattice.build_tables == the reference's own "
                                 "flattening (tests/golden/tables.json); engine-call time only, Python graph build excluded"}
    line = {"impl": "reference", "metric": "attempted Metropolis spin updates per second", "value": val, "unit": "attempts/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * eng / max(1, a.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {"value": val, "unit": "attempts/s", "cores": cores, "kind": ref_kind(), "sample": sample, "larger_sample": big},
            "e2e": {"value": val, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    emit(line)


def workload_config(a, world):
    return {"workload": "heisenberg_sc_%d^3_Tscan_%d_replicas_per_gpu" % (a.L, a.replicas), "lattice": "simple cubic %d^3" % a.L,
            "spins": a.L ** 3, "replicas_per_gpu": a.replicas, "replicas_total": a.replicas * world,
            "sweeps_per_step": a.sweeps, "measurement": "every sweep (fused M,E; reference definitions)",
            "state": "fp32 SoA planes, 12 B/spin", "parallelism": "replica-sharded x%d (no data-path collective)" % world,
            "l2": "inputs larger than L2: %.0f MB of spin state per colour pass" % (a.replicas * a.L ** 3 * 12 / 1e6),
            "dipole": "not in the headline workload; the 'configs' and 'pt' legs of this line run C5 with the dipole stencil"}


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None,
                "power_w_max": max(pw) if pw else None}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def wolff_metric(device, L=4096, steps=100, peak=None):
    """BASELINE.json's second metric (Wolff cluster flips/sec) on configs[1]: 2D Ising 4096^2 across Tc.  Per temperature, 8
    replicas of that temperature in one batch, the engine's default path (adaptive hybrid: frontier growth from the seed when
    the cluster is small, global bond passes + union-find when it percolates) next to the global passes alone."""
    from mcsolver_b200 import engine
    from mcsolver_b200.lattice import LatticeSpec
    spec = LatticeSpec(L=(L, L, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J_ISO), (0, 0, (0, 1, 0), J_ISO)])
    R = 8
    rows = []
    saved = os.environ.get("MCG_WOLFF_FRONTIER")
    try:
        for T in [2.0, 2.2, 2.269, 2.35, 2.6]:
            row = {"T": T}
            for mode, key in (("0", "global_passes_only"), (None, "default")):
                if mode is None:
                    os.environ.pop("MCG_WOLFF_FRONTIER", None)
                else:
                    os.environ["MCG_WOLFF_FRONTIER"] = mode
                n = steps if mode is None else max(10, steps // 4)
                with engine.System.from_spec(spec, 1, precision=32, nReplica=R, beta=np.full(R, 1 / T), seed=1, device=device) as s:
                    s.init_spins(0.0)
                    s.metropolis_sweeps(20)
                    s.wolff_steps(20)
                    c0 = [s.counters(r) for r in range(R)]
                    f0 = sum(s.wolff_frontier_steps(r) for r in range(R))
                    t0 = time.time()
                    s.wolff_steps(n)          # synchronous: returns after the stream has drained
                    dt = time.time() - t0
                    c1 = [s.counters(r) for r in range(R)]
                    flipped = sum(b[2] - a[2] for a, b in zip(c0, c1))
                    nfront = sum(s.wolff_frontier_steps(r) for r in range(R)) - f0
                row[key] = {"cluster_updates_per_s": n * R / dt, "flipped_spins_per_s": flipped / dt,
                            "mean_cluster_fraction": flipped / (n * R) / spec.nsite, "steps_by_frontier_growth": nfront / (n * R)}
                if mode == "0" and peak:
                    # bond pass: spin + own forest word read, forest word written; flip pass: forest word + spin read, spin and the
                    # next step's forest word written: 7 words of 4 B per site per update (finds through the forest come on top)
                    row[key]["hbm_frac_at_28B_per_site"] = n * R * spec.nsite * 28 / dt / (peak * 1e9)
            row["speedup"] = row["default"]["cluster_updates_per_s"] / row["global_passes_only"]["cluster_updates_per_s"]
            rows.append(row)
    finally:
        if saved is None:
            os.environ.pop("MCG_WOLFF_FRONTIER", None)
        else:
            os.environ["MCG_WOLFF_FRONTIER"] = saved
    tot = sum(r["default"]["cluster_updates_per_s"] for r in rows)
    return {"workload": "ising_square_%d^2 (J=-1), Wolff single-cluster updates, %d replicas per temperature" % (L, R),
            "per_temperature": rows, "cluster_updates_per_s": tot / len(rows),
            "flipped_spins_per_s": sum(r["default"]["flipped_spins_per_s"] for r in rows) / len(rows),
            "algorithm": "per step and replica: breadth-first growth from the seed over per-bond Philox words (one thread block, O(cluster)); "
                         "clusters that outgrow the 32768-site queue take the global passes (bond activation + atomicCAS union-find, O(N)); "
                         "same clusters either way (tests/test_gpu_wolff_frontier.py)",
            "bound": "global passes: HBM + atomics; frontier growth: latency (one L2/HBM round trip per breadth-first level)"}


# ---------------------------------------------------------------------------------------------
# every BASELINE.json config at its named size (north_star: "reported at 1, 2, 4 and 8 GPUs"), replica-sharded like the
# headline: each rank runs `R` points of the config's scan, device-timed sweeps with every sweep measured, max over ranks
# ---------------------------------------------------------------------------------------------
def square_spec(L):
    from mcsolver_b200.lattice import LatticeSpec
    return LatticeSpec(L=(L, L, 1), S=[1.0], bonds=[(0, 0, (1, 0, 0), J_ISO), (0, 0, (0, 1, 0), J_ISO)])


def config_table():
    """name -> (spec factory, model, precision, replicas per GPU, T(n), H(n), coordination z, sweeps timed)"""
    from tests.specs import spec_of
    from mcsolver_b200.lattice import add_dipole_stencil
    lin = lambda a, b: (lambda n: np.linspace(a, b, n))
    zero = lambda n: np.zeros(n)
    return [
        ("C1 XY square 4096^2 T-scan", lambda: square_spec(4096), 2, 32, 8, lin(0.9, 1.2), zero, 4, 20),
        ("C2 Ising square 4096^2 T-scan, int8 state", lambda: square_spec(4096), 1, 8, 32, lin(2.0, 2.6), zero, 4, 20),
        ("C2 Ising square 4096^2 T-scan, fp32 state", lambda: square_spec(4096), 1, 32, 32, lin(2.0, 2.6), zero, 4, 20),
        ("C3 CrI3 honeycomb 512^2 x 2 (1NN+2NN+3NN, D)", lambda: spec_of("cri3", (512, 512, 1)), 3, 32, 64, lin(30, 50), zero, 12, 40),
        ("C4 skyrmion hex 1024^2 x 2 (DMI, D, h; Q every sweep)", lambda: spec_of("skyrmion", (1024, 1024, 1)), 3, 32, 32, lambda n: np.full(n, 0.3), lin(0, 0.7), 3, 40),
        ("C5 Heisenberg sc 256^3 T-scan, fp64 state", lambda: cubic_spec(256), 3, 64, 8, ladder, zero, 6, 10),
        ("C5 + dipole stencil r<=2 (32 full-tensor links) sc 256^3", lambda: add_dipole_stencil(cubic_spec(256), 0.1, 2.0), 3, 32, 8, ladder, zero, 32, 8),
    ]


def config_legs(rank, world, local, barrier_max, peak, only=None):
    from mcsolver_b200 import engine
    rows = []
    for name, mk, model, prec, R, Tf, Hf, z, nsw in config_table():
        if only and not any(name.startswith(o) for o in only):
            continue
        try:
            spec = mk()
            T, H = Tf(R * world)[rank * R:(rank + 1) * R], Hf(R * world)[rank * R:(rank + 1) * R]
            with engine.System.from_spec(spec, model, precision=prec, nReplica=R, beta=1 / np.asarray(T, float), field=H, seed=1,
                                         replica_offset=rank * R, device=local) as s:
                C = s.num_colours()
                s.init_spins(0.0)
                s.timed_sweeps(3, with_measure=True)
                barrier_max(0.0)
                ms = s.timed_sweeps(nsw, with_measure=True)
                # share of the colour passes: a second, untimed-for-the-metric run with an event pair around every pass (the
                # events themselves cost the many-small-launch configs ~10 %, so they stay out of the number above)
                s.profile_passes(True)
                ms_prof = s.timed_sweeps(max(2, nsw // 4), with_measure=True)
                pass_ms, npass = s.profile_read()
                s.profile_passes(False)
                jit = s.jit_launch_count() > 0
            t = barrier_max(ms / 1e3)
            att = world * R * spec.nsite * nsw / t
            w = (prec // 8) * model
            balg = (2 + min(C - 1, z)) * w
            rows.append({"config": name, "spins": spec.nsite, "replicas_per_gpu": R, "colours": C, "state": "int8" if prec == 8 else "fp%d" % prec, "sweeps_timed": nsw,
                         "attempts_per_s": att, "bytes_per_attempt": balg, "roofline_frac_per_gpu": att / world * balg / (peak * 1e9),
                         "specialised_kernels": jit, "colour_pass_share_of_sweep": pass_ms / ms_prof, "colour_pass_avg_us": 1e3 * pass_ms / max(1, npass)})
        except Exception as e:        # one config must not take the contract line down
            rows.append({"config": name, "error": str(e)[:300]})
    return rows


def pt_leg(rank, world, local, barrier_max, peak, L=256, R=8, sweeps=40, sps=5):
    """BASELINE configs[4] as named: sc 256^3 Heisenberg + dipole stencil, a ladder of 8 labels per GPU (64 on 8 GPUs), an
    exchange step every `sps` sweeps.  The whole loop runs inside the library (mcg_pt_run): per exchange step one
    ncclAllGather of 3 doubles per replica on the compute stream, then a decide-and-relabel kernel; no host round trip."""
    from mcsolver_b200 import pt
    from mcsolver_b200.lattice import add_dipole_stencil
    out = {}
    spec = add_dipole_stencil(cubic_spec(L), 0.1, 2.0)
    n = R * world
    p = pt.ParallelTempering(spec, 3, ladder(n), precision=32, seed=3, rank=rank, world=world, device=local)
    try:
        p.sys.timed_sweeps(2, with_measure=True)
        p.run(0, sps, sweeps_per_swap=sps, want_results=False)            # first NCCL call (connection set-up) stays untimed
        barrier_max(0.0)
        t_plain = barrier_max(p.sys.timed_sweeps(sweeps, with_measure=True) / 1e3)
        barrier_max(0.0)
        p.run(0, sweeps, sweeps_per_swap=sps, want_results=False)
        t_swap = barrier_max(p.device_ms / 1e3)
        att = n * spec.nsite * sweeps
        C = p.sys.num_colours()
        balg = (2 + min(C - 1, 32)) * 12
        out = {"workload": "heisenberg_sc_%d^3 + dipole stencil r<=2 (32 links), %d-label ladder (%d per GPU), exchange every %d sweeps" % (L, n, R, sps),
               "value": att / t_swap, "unit": "attempts/s", "no_swap_value": att / t_plain, "swap_cost_frac": t_swap / t_plain - 1.0,
               "exchange_steps_timed": sweeps // sps, "bytes_per_attempt": balg, "roofline_frac_per_gpu": att / world / t_swap * balg / (peak * 1e9),
               "collective": "ncclAllGather of 3 doubles per replica inside libmcsolver_b200 (dlopen'ed libnccl, no PyTorch), world=%d" % world,
               "swap_rates_at_this_size": [round(float(x), 3) for x in p.swap_rates()]}
    finally:
        p.close()
    # a ladder that does exchange: same machinery at 64^3 (nearest-neighbour exchange only), temperatures 0.25 % apart around Tc
    spec = cubic_spec(64)
    T = TC * (1.0 + 0.0025 * (np.arange(n) - n / 2))
    p = pt.ParallelTempering(spec, 3, T, precision=32, seed=3, rank=rank, world=world, device=local)
    try:
        p.run(200, 10, sweeps_per_swap=2, want_results=False)
        barrier_max(0.0)
        p.run(0, 400, sweeps_per_swap=2, want_results=False)
        t = barrier_max(p.device_ms / 1e3)
        rows = p.results()
        rates = p.swap_rates()
        out["swapping_ladder"] = {"workload": "heisenberg_sc_64^3, %d labels, dT/T = 0.25 %% around Tc, exchange every 2 sweeps" % n,
                                  "attempts_per_s": n * spec.nsite * 400 / t, "swap_rate_mean": float(rates.mean()), "swap_rate_min": float(rates.min()),
                                  "swap_rate_max": float(rates.max()), "labels_displaced": int(np.sum(p.holders() != np.arange(n))),
                                  "e_over_kT_first_last": [float(rows[0, 8]), float(rows[-1, 8])]}
    finally:
        p.close()
    return out


SHIM_JOB = ("cri3", (32, 32, 1), 40.0, 800, 6400)      # samples/CrI3With2NNCoupling at its own size, sweep counts / 100


def _shim_ref_worker(_):
    """the reference's compiled heisenberglib on the shim job, one host core (runs in a forked child before CUDA is initialised)"""
    from oracle import refharness as rh
    from mcsolver_b200.lattice import build_tables
    from tests.specs import spec_of
    name, L, T, nth, nsw = SHIM_JOB
    t = build_tables(spec_of(name, L), T, 3)
    args = t.on_args(0, nth // 10, nsw // 10, t.N, 0.0, 0.0, 0)     # a tenth of the job: ~0.5 s of CPU
    t0 = time.time()
    out = rh.run_ref_engine(3, args, seed=1)
    return time.time() - t0, t.N * (nth // 10 + nsw // 10), out[8]


def shim_ref_sample():
    if not ref_kind() == "reference":
        return None
    import multiprocessing as mp
    with mp.get_context("fork").Pool(processes=1, initializer=_quiet_stdout) as pool:
        dt, att, e = pool.map(_shim_ref_worker, [0])[0]
    return {"seconds": dt, "attempts": att, "attempts_per_s": att / dt, "e_per_site_over_kT": e}


def shim_leg(ref):
    """End to end through the reference-facing boundary itself: ONE heisenberglib.MCMainFunction(*args) call of the drop-in module
    (mcsolver_b200/lib/heisenberglib.py) with the 23 positional arguments mcMain.py:239-248 builds - host tuples in, the 29-item
    result tuple out, marshalling, upload and download inside the timed call - at the reference's own job size."""
    from mcsolver_b200.lattice import build_tables
    from tests.specs import spec_of
    sys.path.insert(0, os.path.join(ROOT, "mcsolver_b200", "lib"))
    import heisenberglib
    name, L, T, nth, nsw = SHIM_JOB
    t = build_tables(spec_of(name, L), T, 3)
    args = t.on_args(0, nth, nsw, t.N, 0.0, 0.0, 0)
    nbytes = sum(np.asarray(x).nbytes for x in args if isinstance(x, (tuple, list)))
    heisenberglib.MCMainFunction(*t.on_args(0, 10, 20, t.N, 0.0, 0.0, 0))      # context, module load
    times = []
    for _ in range(3):
        t0 = time.time()
        out = heisenberglib.MCMainFunction(*args)
        times.append(time.time() - t0)
    dt = min(times)
    att = t.N * (nth + nsw)
    res = {"workload": "samples/CrI3With2NNCoupling: honeycomb 32x32x2 (z = 12, D), T = 40, %d + %d sweeps (the sample's counts / 100), one (T,H) point "
                       "= one MCMainFunction call, fp64 state (the shim's default)" % (nth, nsw),
           "seconds_per_call": dt, "value": att / dt, "unit": "attempts/s", "h2d_bytes_per_call": int(nbytes), "d2h_bytes_per_call": 29 * 8,
           "result_tuple_items": len(out), "e_per_site_over_kT": float(out[8]),
           "note": "one point occupies one thread block (resident kernel); a scan's points run concurrently on different SMs where the reference "
                   "runs one process per core"}
    if ref:
        res["reference_one_core"] = ref
        res["speedup_vs_reference_one_core"] = res["value"] / ref["attempts_per_s"]
    return res


def slab_leg(rank, world, local, barrier_max, a, peak, sweeps=40):
    """ONE lattice over all GPUs (SURVEY 8e "largest lattice"): Heisenberg sc (256 N) x 256 x 256 cut into N slabs along x, the same
    8 replicas and 256^3 sites per GPU as the headline.  After every colour pass the boundary planes go to the neighbours' ghost
    planes (ncclSend/ncclRecv inside the library, on the compute stream) and the raw measurement sums are all-reduced, so the
    number next to the headline's is the price of the halo exchange; N = 1 runs the same code with its own periodic images."""
    from mcsolver_b200 import engine, pt
    from mcsolver_b200.lattice import LatticeSpec
    L, R = a.L, a.replicas
    spec = LatticeSpec(L=(L * world, L, L), S=[1.0], bonds=[(0, 0, (1, 0, 0), J_ISO), (0, 0, (0, 1, 0), J_ISO), (0, 0, (0, 0, 1), J_ISO)])
    port = int(os.environ.get("MASTER_PORT", "29500")) + 31
    cid = pt.comm_id(rank, world, port=port)
    Ts = ladder(R)                                     # every rank holds a slab of each of the R replicas
    with engine.System.from_spec_slab(spec, 3, rank, world, comm_id=cid, precision=32, nReplica=R, beta=1 / Ts, seed=1, device=local) as s:
        s.init_spins(0.0)
        s.timed_sweeps(5, with_measure=True)
        barrier_max(0.0)
        ms = s.timed_sweeps(sweeps, with_measure=True)
        e_check = s.results(0)[0][8]
        info = dict(s.slab)
    t = barrier_max(ms / 1e3)
    att = R * spec.nsite * sweeps / t
    plane_bytes = R * 3 * 4 * (L // 2) * (L // 2) * 4 * 2      # per colour pass and direction: replicas x components x classes x coarse plane x 4 B ... x2 directions
    return {"workload": "heisenberg_sc_%dx%dx%d (one lattice, %d slabs along x), %d replicas, fp32" % (L * world, L, L, world, R),
            "value": att, "unit": "attempts/s", "sweeps_timed": sweeps, "roofline_frac_per_gpu": att / world * 36 / (peak * 1e9),
            "halo_bytes_per_colour_pass_per_gpu": plane_bytes if world > 1 else 0, "slab_of_rank0": info,
            "collectives": "per colour pass ncclSend/ncclRecv of the boundary coarse planes to both neighbours; per measured sweep one "
                           "ncclAllReduce of 12 doubles per replica (dlopen'ed libnccl, no PyTorch)" if world > 1 else
                           "world = 1: ghost planes are the slab's own periodic images (device copy), no collective",
            "check_e_per_site_over_kT_replica0": float(e_check)}


def fp64_state_leg(rank, world, local, barrier_max, spec, Ts, a, peak):
    """The headline workload with fp64 spin state - the reference's own precision, the shims' default (24 B/spin, 72 B per
    attempt) - timed exactly like the headline: W warm-up steps, then K steps of `sweeps` measured sweeps, CUDA events on the
    launch stream, max over ranks."""
    from mcsolver_b200 import engine
    try:
        R = len(Ts)
        with engine.System.from_spec(spec, 3, precision=64, nReplica=R, beta=1 / np.asarray(Ts), seed=1, replica_offset=rank * R, device=local) as s:
            s.init_spins(0.0)
            for _ in range(a.warmup):
                s.timed_sweeps(a.sweeps, with_measure=True)
            s.profile_passes(True)
            barrier_max(0.0)
            ms = 0.0
            for _ in range(a.steps):
                ms += s.timed_sweeps(a.sweeps, with_measure=True)
            pass_ms, npass = s.profile_read()
            s.profile_passes(False)
            jit = s.jit_launch_count() > 0
        t = barrier_max(ms / 1e3)
        att = world * R * spec.nsite * a.sweeps * a.steps / t
        per_launch = R * spec.nsite / 2 * 72 / (pass_ms / 1e3 / max(1, npass)) / 1e9
        return {"value": att, "unit": "attempts/s", "dtype": "f64", "bytes_per_attempt": 72, "steps": a.steps, "sweeps_per_step": a.sweeps,
                "ms_per_step": 1e3 * t / a.steps, "roofline_frac_per_gpu": att / world * 72 / (peak * 1e9),
                "roofline": {"bound": "hbm", "achieved": per_launch, "peak": peak, "unit": "GB/s", "frac": per_launch / peak,
                             "kernel": "mcg_pass_m1 = pass_body<NC=3,double,diagJ,MODE=1,V=2>, NVRTC-specialised, bulk L2 prefetch of the rows that miss",
                             "avg_launch_ms": pass_ms / max(1, npass), "launches_timed": npass},
                "specialised_kernels": jit}
    except Exception as e:      # never let this leg break the contract line
        return {"error": str(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--L", type=int, default=256)
    ap.add_argument("--replicas", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=100, help="Metropolis sweeps (each measured) per step")
    ap.add_argument("--ref-L", type=int, default=32)
    ap.add_argument("--ref-sweeps", type=int, default=60)
    ap.add_argument("--ref-L2", type=int, default=48, help="second, larger CPU sample of the reference arm")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config legs (C1-C5 at their named sizes)")
    ap.add_argument("--no-pt", action="store_true", help="skip the parallel-tempering leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wolff", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the one-lattice-over-all-GPUs (slab decomposition) leg")
    ap.add_argument("--no-fp64", action="store_true", help="skip the fp64-state leg of the headline workload")
    ap.add_argument("--e2e-sweeps", type=int, default=1000, help="measured sweeps per end-to-end job")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: whatever libraries print to fd 1 on the way (NCCL's version banner, the reference's PyInit
    # printf) goes to stderr instead; the line itself is written to the original descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if a.impl == "reference":
        a.steps = min(a.steps, 100)    # each step is a bounded ~1.5 s CPU sample: K steps stay within a few minutes
        run_reference_arm(a, rank, world, emit)
        return

    # ---- CPU baseline first (rank 0, N=1): forks workers, so it must precede CUDA initialisation
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = min(host_cores(), 128)
        att, eng, _ = reference_step(cores, a.ref_L, 2, a.ref_sweeps)
        cpu = {"value": att / eng, "unit": "attempts/s", "cores": cores, "kind": ref_kind(),
               "sample": "reference C engine (oracle/_ref heisenberglib), sc %d^3, %d temperature points x (2+%d) sweeps, one "
                         "process per point, engine-call time %.1f s" % (a.ref_L, cores, a.ref_sweeps, eng)}

    shim_ref = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            shim_ref = shim_ref_sample()
        except Exception:
            shim_ref = None

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def barrier_max(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from mcsolver_b200 import engine, scan
    spec = cubic_spec(a.L)
    N, R = spec.nsite, a.replicas
    Ts = ladder(R * world)[rank * R:(rank + 1) * R]
    s = engine.System.from_spec(spec, 3, precision=32, nReplica=R, beta=1.0 / Ts, field=np.zeros(R), seed=1,
                                replica_offset=rank * R, device=local)
    s.init_spins(0.0)
    for _ in range(a.warmup):
        s.timed_sweeps(a.sweeps, with_measure=True)
    s.reset_measurements()
    s.profile_passes(True)
    module_key = s.jit_module_key(1) if s.jit_launch_count() > 0 else None     # identity of the kernel the roofline line is about
    l0 = s.launch_count()
    j0 = s.jit_launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    barrier_max(0.0)
    t0 = time.time()
    dev_ms = 0.0
    for _ in range(a.steps):
        dev_ms += s.timed_sweeps(a.sweeps, with_measure=True)      # CUDA events on the launch stream; syncs
    wall = time.time() - t0
    t_max = barrier_max(dev_ms / 1e3)
    clk = clocks.stop()
    launches = s.launch_count() - l0
    jit_launches = s.jit_launch_count() - j0      # of those, launches of the NVRTC-specialised module (cuLaunchKernel of mcg_pass_m1)
    pass_ms, npass = s.profile_read()
    s.profile_passes(False)
    out0 = s.results(0)[0]
    s.close()

    attempts_rank = R * N * a.sweeps * a.steps
    value = world * attempts_rank / t_max

    # ---- roofline of the dominant kernel (colour pass with fused measurement), per launch
    w_bytes = 12                                   # fp32 x,y,z planes
    b_alg = 3 * w_bytes                            # own read + own write + each neighbour-colour spin once (2 colours)
    attempts_per_launch = R * N / 2
    avg_launch_s = pass_ms / 1e3 / max(1, npass)
    achieved = attempts_per_launch * b_alg / avg_launch_s / 1e9
    peak, peak_src = measured_peak()
    # DRAM traffic per launch comes from an ncu capture of EXACTLY this module (profiles/traffic.json is keyed by the module
    # key = the hash of the generated lattice prologue and the kernel headers); any other build of the kernel gets null
    traffic, traffic_note = None, "no ncu capture on record"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if module_key is not None and tj.get("module_key") == module_key:
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("source")
        else:
            traffic_note = "profiles/traffic.json was captured for module %s, this run launched %s" % (tj.get("module_key"), module_key)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": "mcg_pass_m1 = pass_body<NC=3,float,diagJ,MODE=1 (update + fused measurement),V=4>, NVRTC-specialised for the lattice (struct_pass.cuh)", "bytes_per_attempt": b_alg,
            "attempts_per_launch": attempts_per_launch, "avg_launch_ms": avg_launch_s * 1e3, "launches_timed": npass,
            "kernel_share_of_step": pass_ms / dev_ms, "peak_source": peak_src,
            "module_key": module_key, "specialised_kernel_launches_timed": int(jit_launches), "traffic_source": traffic_note,
            "algorithmic_bytes_per_launch": attempts_per_launch * b_alg}

    # ---- e2e: whole jobs through the public API, host descriptors in, host result rows out.  The first job creates the system
    #      (allocation, tables, specialised modules); the following ones find it in scan's pool and recycle it - both are inside
    #      the timed region.
    e2e_steps = 3
    desc_bytes = 8 * (len(spec.S) * 4 + len(spec.bonds) * 14) + 16 * R
    scan.clear_pool()
    barrier_max(0.0)
    t0 = time.time()
    e2e_job_s = []
    for k in range(e2e_steps):
        tj = time.time()
        idx, rows, _ = scan.run_points(spec, 3, ladder(R * world), np.zeros(R * world), 0, a.e2e_sweeps, precision=32, seed=1 + k,
                                       rank=rank, world=world, device=local)
        e2e_job_s.append(round(time.time() - tj, 4))
    e2e_t = barrier_max(time.time() - t0)
    scan.clear_pool()
    e2e = {"value": world * R * N * a.e2e_sweeps * e2e_steps / e2e_t, "unit": "attempts/s", "h2d_bytes_per_step": desc_bytes,
           "d2h_bytes_per_step": int(rows.nbytes), "job_seconds": e2e_job_s, "jobs": e2e_steps, "sweeps_per_job": a.e2e_sweeps,
           "what": "scan.run_points(): system from host descriptor (created by the first job, recycled by the others), init, %d measured "
                   "sweeps, result rows to host; host wall clock over all jobs, max over ranks" % a.e2e_sweeps}

    # ---- secondary metric of BASELINE.json ("Wolff cluster flips/sec"), outside the timed region, N=1 only:
    #      2D Ising 4096^2 (configs[1]) at five temperatures across Tc, union-find cluster updates
    wolff = None
    if rank == 0 and world == 1 and not a.no_wolff:
        wolff = wolff_metric(local, peak=peak)

    # ---- the same workload at the reference's precision (fp64 state), timed over the same K steps, at every N
    f64 = None if a.no_fp64 else fp64_state_leg(rank, world, local, barrier_max, spec, Ts, a, peak)

    # ---- every BASELINE config at its named size and the parallel-tempering ladder (C5 as named), at every N
    configs = None if a.no_configs else config_legs(rank, world, local, barrier_max, peak)
    ptres = None
    if not a.no_pt:
        try:
            ptres = pt_leg(rank, world, local, barrier_max, peak)
        except Exception as e:
            ptres = {"error": str(e)[:300]}

    shim = None
    if rank == 0 and world == 1 and not a.no_wolff:
        try:
            shim = shim_leg(shim_ref)
        except Exception as e:
            shim = {"error": str(e)[:300]}

    slab = None
    if not a.no_slab:
        try:
            slab = slab_leg(rank, world, local, barrier_max, a, peak)
        except Exception as e:
            slab = {"error": str(e)[:300]}

    if rank == 0:
        line = {"metric": "attempted Metropolis spin updates per second", "value": value, "unit": "attempts/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_max / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(a, world), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "wolff": wolff, "fp64_state": f64, "configs": configs, "pt": ptres, "slab": slab, "shim": shim, "host_wall_s": wall,
                "check": {"replica0_T": float(Ts[0]), "e_per_site_over_kT": float(out0[8]), "U4": float(out0[10])}}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
