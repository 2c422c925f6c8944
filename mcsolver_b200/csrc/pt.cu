// Parallel tempering over a replica ladder (new capability, no reference counterpart).
//
// Configurations never move: a swap exchanges the (beta, H) LABELS of two replicas.  Per swap
// step every rank contributes, per local replica, the bond+anisotropy energy E0 and the field-axis
// magnetisation M of its current configuration (two doubles; the only data that crosses NVLink, by
// an allgather done by the host layer), then every rank evaluates the same Philox-driven decisions
// (mcg_pt_decide is pure host arithmetic) and applies the resulting labels to its own replicas.
// Measurement accumulators live per label, so <O>(T_k) is collected correctly while labels wander;
// the per-label accumulators of all ranks are summed at the end (they are plain sums over sweeps).
//
// Two drivers share the bookkeeping:
//  * host-driven (mcg_pt_state / mcg_pt_decide / mcg_pt_set_labels): the caller brings its own allgather - used by the
//    CPU multi-process tests (gloo) and for debugging;
//  * in-library (mcg_pt_setup / mcg_pt_run / mcg_pt_reduce): the whole loop is enqueued on the system's stream - per swap
//    step one pack kernel, ONE ncclAllGather of three doubles per replica (NCCL is dlopen'ed like NVRTC: no link-time
//    dependency, no PyTorch), one decide-and-relabel kernel that every rank runs on the same gathered data with the same
//    Philox draw.  No host synchronisation between the first sweep and the last.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "system.hpp"
#include "nccl_dl.hpp"

using namespace mcg;

namespace mcg {

constexpr int PT_NSTATE = 3;   // per replica across NVLink: E0, M along the field axis, |M| of the last measured sweep

struct PtState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, n = 0, lo = 0;
    double *d_lbeta = nullptr, *d_lfield = nullptr;      // [n] the fixed ladder
    int32_t *d_holder = nullptr;                         // [n] global replica carrying label k (same on every rank)
    double *d_send = nullptr, *d_all = nullptr;          // [R][3], [n][3]
    unsigned long long *d_att = nullptr, *d_accn = nullptr;   // [n] swap attempts / acceptances of pair (k, k+1)
    double *d_accRed = nullptr, *d_gaccRed = nullptr;    // per-label accumulators summed over ranks (mcg_pt_reduce)
    uint64_t step = 0;
    bool reduced = false;
};

void pt_destroy(PtState *p) {
    if (!p) return;
    if (p->comm && nccl_api().ok) nccl_api().commDestroy(p->comm);
    for (void *b : {(void *)p->d_lbeta, (void *)p->d_lfield, (void *)p->d_holder, (void *)p->d_send, (void *)p->d_all, (void *)p->d_att,
                    (void *)p->d_accn, (void *)p->d_accRed, (void *)p->d_gaccRed})
        pool_free(b);
    delete p;
}

// state of the local replicas for the exchange: E0 = energy without the field term in table units, M along the field
// axis, and the label's last |M| (it travels with the label so that the lag-1 product of heisenbergLib.c:834 is the
// series AT that temperature wherever the label sits)
static __global__ void k_pt_pack(int R, int ax, const double *__restrict__ last, const double *__restrict__ beta, const double *__restrict__ field,
                                 const int32_t *__restrict__ slot, const double *__restrict__ acc, double *send) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double b = beta[r], h = field[r], M = last[4 * r + ax];
    send[PT_NSTATE * r] = last[4 * r] / b + h * M;
    send[PT_NSTATE * r + 1] = M;
    send[PT_NSTATE * r + 2] = acc[(size_t)slot[r] * NACC + ACC_MTMP];
}

// same arithmetic as mcg_pt_decide below, one thread per attempted pair; then every local replica picks up its label
__host__ __device__ __forceinline__ bool pt_accept(const double *beta, const double *field, const double *st, int sw, int k, int a, int b, const RngKey &key,
                                                   uint64_t step) {
    const double Ea = st[sw * a], Ma = st[sw * a + 1], Eb = st[sw * b], Mb = st[sw * b + 1];
    const double Hk_a = Ea - field[k] * Ma, Hk_b = Eb - field[k] * Mb;
    const double Hk1_a = Ea - field[k + 1] * Ma, Hk1_b = Eb - field[k + 1] * Mb;
    const double delta = (beta[k] * Hk_b + beta[k + 1] * Hk1_a) - (beta[k] * Hk_a + beta[k + 1] * Hk1_b);
    uint32_t w[4];
    rng4(key, 0u, STREAM_PT, 0u, step, (uint32_t)k, w);
    return delta <= 0.0 || exp(-delta) > u01<double>(w[0]);
}
static __global__ void __launch_bounds__(256) k_pt_decide(int n, int R, int lo, const double *__restrict__ lbeta, const double *__restrict__ lfield,
                                                          const double *__restrict__ all, int32_t *holder, int parity, RngKey key, uint64_t step,
                                                          unsigned long long *att, unsigned long long *accn, int32_t *slot, double *beta,
                                                          double *field, double *acc) {
    extern __shared__ int32_t prev[];   // holder before this step
    for (int k = threadIdx.x; k < n; k += blockDim.x) prev[k] = holder[k];
    __syncthreads();
    for (int k = parity + 2 * threadIdx.x; k + 1 < n; k += 2 * blockDim.x) {
        const int a = prev[k], b = prev[k + 1];
        const bool ok = pt_accept(lbeta, lfield, all, PT_NSTATE, k, a, b, key, step);
        if (ok) { holder[k] = b; holder[k + 1] = a; }
        att[k] += 1ull;
        accn[k] += ok ? 1ull : 0ull;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int g = holder[k];
        if (g < lo || g >= lo + R) continue;
        const int r = g - lo;
        slot[r] = k; beta[r] = lbeta[k]; field[r] = lfield[k];
        acc[(size_t)k * NACC + ACC_MTMP] = all[PT_NSTATE * prev[k] + 2];   // label k's previous |M|, from whoever held it
    }
}

// accumulator rows for the cross-rank sum: the non-additive state slots count only on the rank that holds the label
static __global__ void k_pt_mask_rows(int n, int R, int lo, const int32_t *__restrict__ holder, const double *__restrict__ acc, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * NACC) return;
    const int k = i / NACC, c = i - k * NACC;
    const bool mine = holder[k] >= lo && holder[k] < lo + R;
    out[i] = ((c == ACC_MTMP || c == ACC_LASTE) && !mine) ? 0.0 : acc[i];
}

template <typename F> int pt_guarded(F &&f) {
    try { f(); return MCG_OK; }
    catch (const Error &e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception &e) { set_last_error(e.what()); return MCG_ERR_ARG; }
}
}  // namespace mcg

extern "C" {

// nLabels accumulator slots (the whole ladder) instead of one per local replica; clears them.
MCG_API int mcg_pt_configure(mcg_system *sys, int nLabels) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys, "system is NULL");
        MCG_REQUIRE(nLabels >= sys->R, "nLabels must be >= the number of local replicas");
        MCG_CUDA(cudaSetDevice(sys->device));
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        pool_free(sys->d_acc);
        sys->d_acc = (double *)pool_alloc(sizeof(double) * (size_t)nLabels * NACC);
        MCG_CUDA(cudaMemset(sys->d_acc, 0, sizeof(double) * (size_t)nLabels * NACC));
        if (sys->d_gacc) {   // per-label group accumulators follow
            size_t n = (size_t)(sys->nG + 2) * (sys->nG + 1);
            pool_free(sys->d_gacc);
            sys->d_gacc = (double *)pool_alloc(sizeof(double) * nLabels * n);
            MCG_CUDA(cudaMemset(sys->d_gacc, 0, sizeof(double) * nLabels * n));
        }
        sys->nLabel = nLabels;
        MCG_CUDA(cudaDeviceSynchronize());   // the clears ran on the legacy stream: complete before sys->stream touches them
    });
}

// state[r] = {E0, M_axis}: E0 = energy without the field term in table units (divide the reduced
// energy of the last measured sweep by beta and add H*M back), M_axis = total spin along the field axis.
MCG_API int mcg_pt_state(mcg_system *sys, double *state /*[R][2]*/) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && state, "NULL argument");
        MCG_CUDA(cudaSetDevice(sys->device));
        std::vector<double> last(4 * (size_t)sys->R);
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        MCG_CUDA(cudaMemcpy(last.data(), sys->d_last, sizeof(double) * last.size(), cudaMemcpyDeviceToHost));
        int ax = sys->model == MCG_HEISENBERG ? 3 : 1;   // z for O(3), x for O(2) and Ising (component 0)
        for (int r = 0; r < sys->R; r++) {
            double b = sys->beta_host[r], h = sys->field_host[r], M = last[4 * r + ax];
            state[2 * r] = last[4 * r] / b + h * M;
            state[2 * r + 1] = M;
        }
    });
}

// replica r now carries ladder label label[r] with parameters (beta[r], field[r])
MCG_API int mcg_pt_set_labels(mcg_system *sys, const int32_t *label, const double *beta, const double *field) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && label && beta && field, "NULL argument");
        MCG_CUDA(cudaSetDevice(sys->device));
        for (int r = 0; r < sys->R; r++) MCG_REQUIRE(label[r] >= 0 && label[r] < sys->nLabel, "label out of range (mcg_pt_configure first)");
        sys->slot_host.assign(label, label + sys->R);
        sys->beta_host.assign(beta, beta + sys->R);
        sys->field_host.assign(field, field + sys->R);
        MCG_CUDA(cudaMemcpyAsync(sys->d_slot, sys->slot_host.data(), sizeof(int32_t) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaMemcpyAsync(sys->d_beta, sys->beta_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaMemcpyAsync(sys->d_field, sys->field_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
    });
}

// Pure host function (no CUDA): one exchange step over the WHOLE ladder of n labels.
//   beta[k], field[k]    parameters of label k (fixed ladder)
//   holder[k]            global replica index currently carrying label k   (in/out)
//   E0[g], M[g]          state of global replica g (allgathered)
// Pairs (k, k+1) with k % 2 == parity are attempted; acceptance min(1, exp(-Delta)),
//   Delta = (beta_k - beta_k+1) * (H(X_b;.) - H(X_a;.)) generalised to differing fields:
//   Delta = [beta_k Hk(X_b) + beta_k1 Hk1(X_a)] - [beta_k Hk(X_a) + beta_k1 Hk1(X_b)],  Hk(X) = E0(X) - field_k M(X).
// The uniform of pair k at `step` is Philox(seed; STREAM_PT, step, k): every rank draws the same.
MCG_API int mcg_pt_decide(int n, const double *beta, const double *field, const double *E0, const double *M, int32_t *holder,
                          int parity, uint64_t seed, uint64_t step, int32_t *accepted /*[n] or NULL*/) {
    return pt_guarded([&] {
        MCG_REQUIRE(n >= 1 && beta && field && E0 && M && holder, "NULL/invalid argument");
        RngKey key = make_rng_key(seed);
        for (int k = (parity & 1); k + 1 < n; k += 2) {
            int a = holder[k], b = holder[k + 1];
            double Hk_a = E0[a] - field[k] * M[a], Hk_b = E0[b] - field[k] * M[b];
            double Hk1_a = E0[a] - field[k + 1] * M[a], Hk1_b = E0[b] - field[k + 1] * M[b];
            double delta = (beta[k] * Hk_b + beta[k + 1] * Hk1_a) - (beta[k] * Hk_a + beta[k + 1] * Hk1_b);
            uint32_t w[4];
            rng4(key, 0u, STREAM_PT, 0u, step, (uint32_t)k, w);
            bool acc = delta <= 0.0 || std::exp(-delta) > u01<double>(w[0]);   // k_pt_decide: the same test on the device
            if (acc) { holder[k] = b; holder[k + 1] = a; }
            if (accepted) accepted[k] = acc ? 1 : 0;
        }
    });
}


// ---------------------------------------------------------------------------------------------
// in-library driver
// ---------------------------------------------------------------------------------------------
MCG_API int mcg_comm_unique_id(char *id, int len) {
    return pt_guarded([&] {
        MCG_REQUIRE(id && len >= (int)sizeof(ncclUniqueId), "id buffer must hold MCG_COMM_ID_BYTES bytes");
        NcclApi &api = nccl_api();
        if (!api.ok) throw Error(MCG_ERR_NCCL, api.why);
        ncclUniqueId u;
        MCG_NCCL(api.getUniqueId(&u));
        std::memcpy(id, &u, sizeof u);
    });
}

// ladder of nLabels = world * nReplica labels; this rank's replicas are the global replicas [rank*R, rank*R + R) and start
// with the labels of the same index.  world > 1 joins the NCCL communicator described by id (mcg_comm_unique_id on rank 0).
MCG_API int mcg_pt_setup(mcg_system *sys, int rank, int world, const char *id, int nLabels, const double *beta, const double *field) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && beta && field, "NULL argument");
        MCG_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
        MCG_REQUIRE(nLabels == world * sys->R, "the ladder must have world * nReplica labels (equal allgather blocks)");
        MCG_REQUIRE((int)sys->replica0 == rank * sys->R, "replica_offset must be rank * nReplica");
        MCG_REQUIRE(nLabels * (int)sizeof(int32_t) <= 48 * 1024, "ladder too long for the decide kernel");
        MCG_CUDA(cudaSetDevice(sys->device));
        if (int rc = mcg_pt_configure(sys, nLabels)) throw Error(rc, mcg_last_error());
        if (sys->pt) { pt_destroy(sys->pt); sys->pt = nullptr; }
        std::unique_ptr<PtState> p(new PtState());
        p->rank = rank; p->world = world; p->n = nLabels; p->lo = rank * sys->R;
        const size_t n = nLabels;
        p->d_lbeta = (double *)pool_alloc(n * sizeof(double));
        p->d_lfield = (double *)pool_alloc(n * sizeof(double));
        p->d_holder = (int32_t *)pool_alloc(n * sizeof(int32_t));
        p->d_send = (double *)pool_alloc((size_t)sys->R * PT_NSTATE * sizeof(double));
        p->d_all = (double *)pool_alloc(n * PT_NSTATE * sizeof(double));
        p->d_att = (unsigned long long *)pool_alloc(n * sizeof(unsigned long long));
        p->d_accn = (unsigned long long *)pool_alloc(n * sizeof(unsigned long long));
        std::vector<int32_t> ident(n);
        for (size_t k = 0; k < n; k++) ident[k] = (int32_t)k;
        cudaStream_t st = sys->stream;
        MCG_CUDA(cudaMemcpyAsync(p->d_lbeta, beta, n * sizeof(double), cudaMemcpyHostToDevice, st));
        MCG_CUDA(cudaMemcpyAsync(p->d_lfield, field, n * sizeof(double), cudaMemcpyHostToDevice, st));
        MCG_CUDA(cudaMemcpyAsync(p->d_holder, ident.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        MCG_CUDA(cudaMemsetAsync(p->d_att, 0, n * sizeof(unsigned long long), st));
        MCG_CUDA(cudaMemsetAsync(p->d_accn, 0, n * sizeof(unsigned long long), st));
        MCG_CUDA(cudaStreamSynchronize(st));
        std::vector<int32_t> lab(sys->R);
        for (int r = 0; r < sys->R; r++) lab[r] = p->lo + r;
        if (int rc = mcg_pt_set_labels(sys, lab.data(), beta + p->lo, field + p->lo)) throw Error(rc, mcg_last_error());
        if (world > 1) {
            MCG_REQUIRE(id, "world > 1 needs the communicator id");
            NcclApi &api = nccl_api();
            if (!api.ok) throw Error(MCG_ERR_NCCL, api.why);
            ncclUniqueId u;
            std::memcpy(&u, id, sizeof u);
            MCG_NCCL(api.commInitRank(&p->comm, world, u, rank));
        }
        sys->pt = p.release();
    });
}

// nthermal + nsweep Metropolis sweeps, every one measured (the exchange needs E and M), an exchange step after every
// sweeps_per_swap sweeps; the per-label accumulators are cleared after thermalisation.  Everything is enqueued on the
// system's stream; the host waits once, at the end.  elapsed_ms (may be NULL): device time between CUDA events.
MCG_API int mcg_pt_run(mcg_system *sys, int64_t nthermal, int64_t nsweep, int sweeps_per_swap, double *elapsed_ms) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && sys->pt, "mcg_pt_setup first");
        MCG_REQUIRE(nthermal >= 0 && nsweep >= 0 && sweeps_per_swap >= 1, "need nthermal, nsweep >= 0 and sweeps_per_swap >= 1");
        MCG_CUDA(cudaSetDevice(sys->device));
        PtState *p = sys->pt;
        NcclApi &api = nccl_api();
        cudaStream_t st = sys->stream;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (elapsed_ms) {
            MCG_CUDA(cudaEventCreate(&e0));
            MCG_CUDA(cudaEventCreate(&e1));
            MCG_CUDA(cudaStreamSynchronize(st));
            MCG_CUDA(cudaEventRecord(e0, st));
        }
        const int ax = sys->model == MCG_HEISENBERG ? 3 : 1;   // d_last = {E, Mx, My, Mz}: z for O(3), x for O(2) and Ising
        const RngKey key = make_rng_key(sys->seed);
        const int64_t total = nthermal + nsweep;
        int64_t done = 0;
        bool cleared = nthermal == 0;
        p->reduced = false;
        while (done < total) {
            const int64_t limit = done < nthermal ? nthermal : total;
            const int64_t n = std::min<int64_t>(sweeps_per_swap, limit - done);
            for (int64_t i = 0; i < n; i++) measured_sweep(sys, 1.0);
            done += n;
            if (!cleared && done >= nthermal) { reset_measurements_async(sys); cleared = true; }
            k_pt_pack<<<(sys->R + 63) / 64, 64, 0, st>>>(sys->R, ax, sys->d_last, sys->d_beta, sys->d_field, sys->d_slot, sys->d_acc, p->d_send);
            if (p->world > 1) MCG_NCCL(api.allGather(p->d_send, p->d_all, (size_t)sys->R * PT_NSTATE, ncclDouble, p->comm, st));
            else MCG_CUDA(cudaMemcpyAsync(p->d_all, p->d_send, sizeof(double) * sys->R * PT_NSTATE, cudaMemcpyDeviceToDevice, st));
            k_pt_decide<<<1, 256, p->n * sizeof(int32_t), st>>>(p->n, sys->R, p->lo, p->d_lbeta, p->d_lfield, p->d_all, p->d_holder, (int)(p->step & 1), key,
                                                              p->step, p->d_att, p->d_accn, sys->d_slot, sys->d_beta, sys->d_field, sys->d_acc);
            p->step++;
            sys->launches += 2;
            MCG_CUDA(cudaGetLastError());
        }
        if (elapsed_ms) MCG_CUDA(cudaEventRecord(e1, st));
        // host mirrors of the labels follow the device
        MCG_CUDA(cudaMemcpyAsync(sys->slot_host.data(), sys->d_slot, sizeof(int32_t) * sys->R, cudaMemcpyDeviceToHost, st));
        MCG_CUDA(cudaMemcpyAsync(sys->beta_host.data(), sys->d_beta, sizeof(double) * sys->R, cudaMemcpyDeviceToHost, st));
        MCG_CUDA(cudaMemcpyAsync(sys->field_host.data(), sys->d_field, sizeof(double) * sys->R, cudaMemcpyDeviceToHost, st));
        MCG_CUDA(cudaStreamSynchronize(st));
        if (elapsed_ms) {
            float ms = 0;
            MCG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            *elapsed_ms = ms;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    });
}

// swap statistics of pair (k, k+1) and the current holder of every label; arrays of nLabels entries (may be NULL)
MCG_API int mcg_pt_stats(mcg_system *sys, int64_t *attempts, int64_t *accepts, int32_t *holder) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && sys->pt, "mcg_pt_setup first");
        MCG_CUDA(cudaSetDevice(sys->device));
        PtState *p = sys->pt;
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        static_assert(sizeof(int64_t) == sizeof(unsigned long long), "counter width");
        if (attempts) MCG_CUDA(cudaMemcpy(attempts, p->d_att, sizeof(int64_t) * p->n, cudaMemcpyDeviceToHost));
        if (accepts) MCG_CUDA(cudaMemcpy(accepts, p->d_accn, sizeof(int64_t) * p->n, cudaMemcpyDeviceToHost));
        if (holder) MCG_CUDA(cudaMemcpy(holder, p->d_holder, sizeof(int32_t) * p->n, cudaMemcpyDeviceToHost));
    });
}

// COLLECTIVE: sums the per-label accumulators over the ranks (ncclAllReduce; plain copy for one rank) into a separate
// buffer that mcg_pt_results reads; this rank's own partial sums stay untouched, so the run can be continued.
MCG_API int mcg_pt_reduce(mcg_system *sys) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && sys->pt, "mcg_pt_setup first");
        MCG_CUDA(cudaSetDevice(sys->device));
        PtState *p = sys->pt;
        NcclApi &api = nccl_api();
        cudaStream_t st = sys->stream;
        const size_t na = (size_t)p->n * NACC, ng = sys->d_gacc ? (size_t)p->n * (sys->nG + 2) * (sys->nG + 1) : 0;
        if (!p->d_accRed) p->d_accRed = (double *)pool_alloc(na * sizeof(double));
        if (ng && !p->d_gaccRed) p->d_gaccRed = (double *)pool_alloc(ng * sizeof(double));
        k_pt_mask_rows<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(p->n, sys->R, p->lo, p->d_holder, sys->d_acc, p->d_accRed);
        MCG_CUDA(cudaGetLastError());
        if (ng) MCG_CUDA(cudaMemcpyAsync(p->d_gaccRed, sys->d_gacc, ng * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (p->world > 1) {
            MCG_NCCL(api.allReduce(p->d_accRed, p->d_accRed, na, ncclDouble, ncclSum, p->comm, st));
            if (ng) MCG_NCCL(api.allReduce(p->d_gaccRed, p->d_gaccRed, ng, ncclDouble, ncclSum, p->comm, st));
        }
        MCG_CUDA(cudaStreamSynchronize(st));
        p->reduced = true;
    });
}

// result tuple of ladder label `label` from the cross-rank sums (mcg_pt_reduce first); layout as mcg_results
MCG_API int mcg_pt_results(mcg_system *sys, int label, double *out, double *groupOut) {
    return pt_guarded([&] {
        MCG_REQUIRE(sys && sys->pt && sys->pt->reduced, "mcg_pt_reduce first");
        MCG_CUDA(cudaSetDevice(sys->device));
        results_from(sys, sys->pt->d_accRed, sys->pt->d_gaccRed, label, out, groupOut);
    });
}

}  // extern "C"
