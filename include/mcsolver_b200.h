/* mcsolver_b200.h - C ABI of libmcsolver_b200.so, the B200-native Monte Carlo engine.
 *
 * Drop-in boundary: the reference's native engines are three CPython extension
 * modules exporting one function each,
 *     heisenberglib.MCMainFunction   /root/reference/mcsolver/heisenbergLib.c:478
 *     xylib.MCMainFunction           /root/reference/mcsolver/xyLib.c:413
 *     isinglib.MCMainFunction        /root/reference/mcsolver/isingLib.c:259
 * called from mcMain.py:119-124 (Ising) and mcMain.py:239-248 (XY/Heisenberg).  The shims
 * mcsolver_b200/lib/{isinglib,xylib,heisenberglib}.py expose the same MCMainFunction(*args)
 * and bind the entry points below through ctypes.  Plain pointers and sizes only; every
 * function returns 0 on success and a non-zero mcg_status otherwise, with a thread-local
 * message available from mcg_last_error() (the reference signals no errors at all and
 * segfaults on bad input: heisenbergLib.c:504 ignores PyArg_ParseTuple's result).
 *
 * There is no CPU fallback: if no CUDA device is usable every entry point fails with
 * MCG_ERR_CUDA.
 */
#ifndef MCSOLVER_B200_H
#define MCSOLVER_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define MCG_API __attribute__((visibility("default")))
#else
#define MCG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MCG_OK = 0,
    MCG_ERR_ARG = 1,      /* invalid argument (bad index, bad size, NULL, unsupported combination) */
    MCG_ERR_CUDA = 2,     /* CUDA runtime failure / no device */
    MCG_ERR_ALLOC = 3,
    MCG_ERR_STATE = 4,    /* call sequence error (e.g. results before any measurement) */
    MCG_ERR_NCCL = 5
} mcg_status;

typedef struct mcg_system mcg_system; /* opaque: tables + replica states resident in HBM */

/* models: the reference's three engines */
#define MCG_ISING 1       /* isingLib.c      */
#define MCG_XY 2          /* xyLib.c         */
#define MCG_HEISENBERG 3  /* heisenbergLib.c */

/* update algorithms (first positional argument of MCMainFunction, mcMain.py:107-108,147-148) */
#define MCG_METROPOLIS 0  /* localUpdate: heisenbergLib.c:441, xyLib.c:382, isingLib.c:238 */
#define MCG_WOLFF 1       /* blockUpdate: heisenbergLib.c:373, xyLib.c:319, isingLib.c:202 */

/* Per-site tables: exactly the payload of one MCMainFunction call (SURVEY 8b "Signature").
 * O(n): heisenbergLib.c:504-585.  Ising: isingLib.c:277-335 (D, tri, groups unused; J has one
 * number per link instead of nine).  All couplings are in units of kT_table (the reference
 * passes everything pre-divided by T, mcMain.py:21-31); per-replica beta[] (mcg_config) scales
 * them further, so tables may also be left unscaled and beta[r] = 1/T_r used instead. */
typedef struct {
    int32_t model;            /* MCG_ISING | MCG_XY | MCG_HEISENBERG */
    int32_t N;                /* totOrbs */
    int32_t maxL;             /* maxNLinking */
    const double *S;          /* [N]   initSpin: signed spin length (Ising: the +-S configuration) */
    const double *D;          /* [N*3] initD (x,y,z as parsed); NULL = 0 */
    const int32_t *nlink;     /* [N] */
    const double *J;          /* O(n): [N*maxL*9] xx,yy,zz,xy,xz,yz,yx,zx,zy ; Ising: [N*maxL] */
    const int32_t *nbr;       /* [N*maxL] linkedOrb, -1 padded */
    int32_t nTri;             /* number of triangle circuits */
    const int32_t *tri;       /* [nTri*3] localCircuits */
    int32_t nLat;             /* number of correlated pairs */
    const int32_t *pairs;     /* [nLat*2] corrOrbPair */
    int32_t nG, maxG;         /* nOrbGroup, maxOrbGroupSize */
    const int32_t *groups;    /* [nG*maxG] orbGroupList, -1 padded */
    int32_t nR, nC;           /* block-spin sites, cluster size */
    const int32_t *rOrb;      /* [nR] */
    const int32_t *rCl;       /* [nR*nC] rOrbCluster */
    const int32_t *rNbr;      /* [nR*maxL] linkedOrb_rnorm */
    int32_t ignoreOffDiag;    /* ignoreNonDiagonalJ: use only Jxx,Jyy,Jzz (heisenbergLib.c:592-593) */
} mcg_tables;

/* Compact translation-invariant lattice: what a parameter file holds (fileio.py:10-22) and what
 * Lattice.py:155-284 expands into one Python object per orbital.  The engine expands it on the
 * device instead (structured path: neighbour indices are computed, not stored). */
typedef struct {
    int32_t src, tgt;         /* orbital indices inside the cell */
    int32_t d[3];             /* overLat: target cell = source cell + d (periodic) */
    double J[9];              /* xx,yy,zz,xy,xz,yz,yx,zx,zy (Ising: J[0]) - Lattice.py:127-137 */
} mcg_bond;

typedef struct {
    int32_t model;
    int32_t L[3];             /* supercell Lx,Ly,Lz ; site id = ((x*Ly+y)*Lz+z)*norb+o (Lattice.py:171-184) */
    int32_t norb;
    const double *S;          /* [norb] */
    const double *D;          /* [norb*3]; NULL = 0 */
    int32_t nbond;
    const mcg_bond *bonds;
    int32_t pair_s, pair_t;   /* correlated pair: orbital pair_s in cell, pair_t in cell+pair_d */
    int32_t pair_d[3];
    int32_t ncircuit;         /* triangles per cell */
    const int32_t *circuits;  /* [ncircuit*3*4]: per vertex (orb, dx, dy, dz) - Lattice.py:223-234 */
    int32_t ngroup;           /* orbital groups (fileio orbGroupList); 0 = none */
    const int32_t *group_mask;/* [ngroup*norb] 1 if the orbital belongs to the group */
    int32_t group_in_sc;      /* 1: "Supergroup" - a group spans the whole supercell (Lattice.py:208-209); 0: cell (0,0,0) only */
    int32_t block_spin;       /* 1: block-spin ("renormalised lattice") statistics every measured sweep - tuple slots 11-19,
                                 Ising 6-7 (heisenbergLib.c:748-803); one more read of the configuration per sweep.
                                 0: those slots stay 0.  Needs even supercell dims; the table path always computes them. */
} mcg_lattice_desc;

typedef struct {
    int32_t precision;        /* 32: fp32 spin state and arithmetic; 64: fp64 (parity level 1);
                                 8 (mcg_create_lattice, Ising, one exchange constant and one |S|): spins as int8, 1 byte per spin,
                                 acceptance by integer thresholds of the fp64 probabilities - same decisions as precision 64 */
    int32_t nReplica;         /* independent (T,H) points / PT replicas resident on this GPU */
    const double *beta;       /* [nReplica] multiplies J and D;  NULL = 1.0 (tables pre-scaled) */
    const double *field;      /* [nReplica] h (units of the table energies; multiplied by beta) */
    uint64_t seed;            /* Philox key */
    int32_t replica_offset;   /* global index of local replica 0 (keeps RNG streams GPU-count independent) */
    int32_t device;           /* CUDA device ordinal, -1 = current */
} mcg_config;

MCG_API const char *mcg_last_error(void);
MCG_API int mcg_version(void);
MCG_API int mcg_device_count(int *count);

/* ---- system life cycle ---- */
MCG_API int mcg_create_tables(const mcg_tables *t, const mcg_config *cfg, mcg_system **out);
MCG_API int mcg_create_lattice(const mcg_lattice_desc *d, const mcg_config *cfg, mcg_system **out);
MCG_API int mcg_destroy(mcg_system *sys);
/* Host-only: build the colouring/class tables of a descriptor and compile its specialised colour-pass
 * kernels with NVRTC for sm_100a (no GPU needed).  ncompiled = kernels modules built (one per colour). */
MCG_API int mcg_jit_check(const mcg_lattice_desc *d, int precision, int *ncompiled, char *report, int report_len);
MCG_API int mcg_num_colours(const mcg_system *sys, int *ncolours);
MCG_API int mcg_colour_order(const mcg_system *sys, int32_t *order /*[N] site ids, colour-major*/);
/* ---- one lattice over several GPUs ("the largest lattice can optionally be domain-decomposed with halo exchange"; the
 * reference has no decomposition, README.md:99) ----
 * desc describes the WHOLE lattice (three-dimensional supercell); rank r of world creates the slab of L[0]/world planes it owns plus
 * one ghost colouring period on either side.  comm_id: mcg_comm_unique_id() of rank 0 (NULL for world == 1).  Metropolis sweeps
 * (mcg_run, mcg_metropolis_sweeps, mcg_timed_sweeps, mcg_measure, mcg_energy) are collective over the ranks: after every colour
 * pass the boundary planes travel to the neighbours' ghosts (ncclSend/ncclRecv on the compute stream) and the raw measurement
 * sums are all-reduced, so mcg_results is the whole lattice's on every rank.  RNG streams are keyed by the global site ids: the
 * slabs together reproduce the undivided lattice's trajectory bit for bit.  mcg_get_spins / mcg_set_spins address the local
 * planes, ghosts included (mcg_slab_info: {rank, world, first own x, own planes, ghost planes per side, global L[0]}); after
 * mcg_set_spins call mcg_slab_sync (collective) to refresh the ghosts.  Not decomposed: Wolff updates, topological charge,
 * block-spin and orbital-group statistics, int8 state. */
MCG_API int mcg_create_lattice_slab(const mcg_lattice_desc *desc, const mcg_config *cfg, int rank, int world, const char *comm_id, mcg_system **out);
MCG_API int mcg_slab_info(const mcg_system *sys, int32_t *info6);
/* the same six numbers for rank r of world without creating anything (host only): how a lattice would be cut, or why it cannot */
MCG_API int mcg_slab_plan(const mcg_lattice_desc *desc, int precision, int rank, int world, int32_t *info6);
MCG_API int mcg_slab_sync(mcg_system *sys);

/* Philox stream layout of the Metropolis sweeps (csrc/rng.cuh), for checkers that restate the trajectory:
 * stride = 0: one block per site (table-built systems, scalar structured pass); stride = S, group = V: the V sites
 * id = base + m*S of one vector item share their blocks. */
MCG_API int mcg_rng_layout(const mcg_system *sys, int32_t *stride, int32_t *group);
MCG_API int mcg_set_params(mcg_system *sys, const double *beta, const double *field); /* per replica */
/* Reuse a created system for another job on the same lattice (the reference creates every grid point from scratch in a fresh
 * pool worker, win.py:90-91 / mcMain.py:150-223): new per-replica beta/field (NULL keeps them), new seed and replica offset, all
 * RNG counters, measurement accumulators and attempt counters back to zero.  After mcg_init_spins the run is bit for bit the one
 * a fresh mcg_create_* with the same configuration performs. */
MCG_API int mcg_recycle(mcg_system *sys, const double *beta, const double *field, uint64_t seed, int replica_offset);

/* ---- state ---- */
/* initial state of establishLattice (heisenbergLib.c:157-172): normalise((S,0,0)+flunc*n)*|S| */
MCG_API int mcg_init_spins(mcg_system *sys, double flunc);
MCG_API int mcg_set_spins(mcg_system *sys, int replica, const double *spins /* O(n):[N*3]  Ising:[N] */);
MCG_API int mcg_get_spins(mcg_system *sys, int replica, double *spins);

/* ---- parity level 1: energy of the resident configuration, fp64 accumulation ----
 * getCorrEnergy / getOnsiteEnergy (heisenbergLib.c:238-253, xyLib.c:191-204, isingLib.c:121-127,232).
 * Etot = sum_i ebond_i/2 + sum_i eonsite_i, in beta*E units of the replica.  Per-site arrays may be NULL. */
MCG_API int mcg_energy(mcg_system *sys, int replica, double *Etot, double *ebond_site, double *eonsite_site);

/* ---- updates ---- */
/* nsweeps colour-class Metropolis sweeps (each site attempted once per sweep with probability
 * pAttempt); one attempt = localUpdate (heisenbergLib.c:441-473 / xyLib.c:382-409 / isingLib.c:238-254) */
MCG_API int mcg_metropolis_sweeps(mcg_system *sys, int64_t nsweeps, double pAttempt);
/* Same as mcg_metropolis_sweeps, bracketed by CUDA events on the stream the kernels are launched on;
 * with_measure != 0 also produces and accumulates the per-sweep measurements after every sweep
 * (fused into the colour passes on structured systems).  elapsed_ms = device time of the region. */
MCG_API int mcg_timed_sweeps(mcg_system *sys, int64_t nsweeps, double pAttempt, int with_measure, double *elapsed_ms);
/* nsteps Wolff single-cluster updates (blockUpdate) by bond activation + union-find labelling */
MCG_API int mcg_wolff_steps(mcg_system *sys, int64_t nsteps);
/* how many of the replica's Wolff steps so far were completed by frontier growth from the seed (O(cluster) work) instead of
 * the global bond passes (O(N)); lattices of >= 65536 sites choose per step and per replica (MCG_WOLFF_FRONTIER=0 disables).
 * Same clusters either way: the reference grows its FIFO from the seed, isingLib.c:165-236, heisenbergLib.c:310-439 */
MCG_API int mcg_wolff_frontier_steps(mcg_system *sys, int replica, int64_t *steps);

/* ---- measurement (heisenbergLib.c:661-831 definitions) ---- */
MCG_API int mcg_measure(mcg_system *sys);              /* accumulate one "sweep" worth of observables */
MCG_API int mcg_reset_measurements(mcg_system *sys);
/* O(n): out[27] = tuple slots 0..26 of heisenbergLib.c:853-881 ; Ising: out[10] = isingLib.c:435-446.
 * groupOut [(nG+2)*(nG+1)] = slot 28 (may be NULL). */
MCG_API int mcg_results(mcg_system *sys, int replica, double *out, double *groupOut);
MCG_API int mcg_counters(mcg_system *sys, int replica, int64_t *attempts, int64_t *accepted, int64_t *cluster_sites);

/* ---- instrumentation (bench.py) ----
 * mcg_launch_count: kernels this system has launched so far.
 * mcg_profile_passes(on): bracket every Metropolis colour-pass launch with CUDA events on the launch stream;
 * mcg_profile_read: total device time (ms) and number of the bracketed launches since the last read.
 * mcg_jit_launch_count: how many of the launches were NVRTC-specialised kernels (mcg_pass_m0/m1, mcg_topo) - lets a
 * test prove which build of the colour pass it exercised. */
MCG_API int mcg_launch_count(mcg_system *sys, int64_t *launches);
MCG_API int mcg_jit_launch_count(mcg_system *sys, int64_t *launches);
/* identity of the specialised module of one colour (hash of the generated lattice prologue and the kernel headers = the file
 * name of its cubin in the cache): measurements taken under a profiler are keyed by it (profiles/traffic.json) */
MCG_API int mcg_jit_module_key(mcg_system *sys, int colour, uint64_t *key);
MCG_API int mcg_profile_passes(mcg_system *sys, int on);
MCG_API int mcg_profile_read(mcg_system *sys, double *total_ms, int64_t *nlaunches);

/* ---- the whole MCMainFunction loop on the resident system ----
 * thermalise nthermal intervals, then nsweep x (ninterval updates + measurement).
 * Metropolis: an interval of ninterval single-site attempts (heisenbergLib.c:614-620) = floor(ninterval/N) whole colour sweeps
 * plus, for a remainder, one sweep in which every site attempts with probability (ninterval mod N)/N.
 * Wolff: ninterval cluster updates per interval.
 * frames: NULL or [nReplica][spinFrame][N*3] (Ising [N]) host buffer (slot 27 / 10). */
MCG_API int mcg_run(mcg_system *sys, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, int spinFrame,
            double *frames);

/* ---- one-shot legacy entry points: one call = one MCMainFunction call ---- */
MCG_API int mcg_run_on(const mcg_tables *t, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, double flunc,
               double h, int spinFrame, uint64_t seed, int precision, double out27[27], double *frames,
               double *groupOut);
MCG_API int mcg_run_ising(const mcg_tables *t, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, double h,
                  int spinFrame, uint64_t seed, int precision, double out10[10], double *frames);

/* ---- parallel tempering (new capability): replica ladder, temperature-LABEL swaps ----
 * Configurations stay where they are; a swap exchanges the (beta, field) labels of two replicas.
 * Per swap step the host layer allgathers mcg_pt_state() of all ranks (2 doubles per replica - the
 * only data crossing NVLink), every rank calls mcg_pt_decide() (pure host arithmetic, same Philox
 * draw everywhere) and applies its share with mcg_pt_set_labels().  Accumulators are per label
 * (mcg_pt_configure), mcg_results(sys, label) reads them; sum them over ranks at the end. */
MCG_API int mcg_pt_configure(mcg_system *sys, int nLabels);
MCG_API int mcg_pt_state(mcg_system *sys, double *state /*[nReplica][2] = E0, M_axis*/);
MCG_API int mcg_pt_set_labels(mcg_system *sys, const int32_t *label, const double *beta, const double *field);
MCG_API int mcg_pt_decide(int n, const double *beta, const double *field, const double *E0, const double *M, int32_t *holder,
                          int parity, uint64_t seed, uint64_t step, int32_t *accepted);
/* raw accumulator row of a label (NACC doubles, plain sums over measured sweeps) for cross-rank reduction,
 * and its inverse; n_acc = row length. */
MCG_API int mcg_acc_get(mcg_system *sys, int label, double *row, int *n_acc);
MCG_API int mcg_acc_set(mcg_system *sys, int label, const double *row);


/* ---- parallel tempering driven by the library: one process per GPU, NCCL allgather per swap step ----
 * SURVEY 8e / north_star: "only per-replica energies cross NVLink through an NCCL allgather at each swap step".
 * mcg_comm_unique_id (rank 0) produces the MCG_COMM_ID_BYTES-byte communicator id that the host layer hands to every
 * rank (any channel: a socket, a file, MPI); mcg_pt_setup joins the communicator and installs the ladder;
 * mcg_pt_run enqueues the whole loop - measured sweeps, pack kernel, ncclAllGather of 3 doubles per replica, decide-and-
 * relabel kernel - on the system's stream without host synchronisation; mcg_pt_reduce (collective) sums the per-label
 * accumulators over ranks with ncclAllReduce; mcg_pt_results reads a label's result tuple from those sums.
 * libnccl.so.2 is dlopen'ed on first use (MCG_NCCL_LIB overrides the path); without it every call that needs a
 * communicator fails with MCG_ERR_NCCL.  world = 1 needs no NCCL. */
#define MCG_COMM_ID_BYTES 128
MCG_API int mcg_comm_unique_id(char *id, int len);
MCG_API int mcg_pt_setup(mcg_system *sys, int rank, int world, const char *id, int nLabels, const double *beta, const double *field);
MCG_API int mcg_pt_run(mcg_system *sys, int64_t nthermal, int64_t nsweep, int sweeps_per_swap, double *elapsed_ms);
MCG_API int mcg_pt_stats(mcg_system *sys, int64_t *attempts, int64_t *accepts, int32_t *holder);
MCG_API int mcg_pt_reduce(mcg_system *sys);
MCG_API int mcg_pt_results(mcg_system *sys, int label, double *out, double *groupOut);

#ifdef __cplusplus
}
#endif
#endif /* MCSOLVER_B200_H */
