// Parallel tempering over a replica ladder (new capability, no reference counterpart).
//
// Configurations never move: a swap exchanges the (beta, H) LABELS of two replicas.  Per swap
// step every rank contributes, per local replica, the bond+anisotropy energy E0 and the field-axis
// magnetisation M of its current configuration (two doubles; the only data that crosses NVLink, by
// an allgather done by the host layer), then every rank evaluates the same Philox-driven decisions
// (mcg_pt_decide is pure host arithmetic) and applies the resulting labels to its own replicas.
// Measurement accumulators live per label, so <O>(T_k) is collected correctly while labels wander;
// the per-label accumulators of all ranks are summed at the end (they are plain sums over sweeps).
#include <cmath>
#include <cstring>
#include <vector>

#include "system.hpp"

using namespace mcg;

namespace mcg {
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
            bool acc = delta <= 0.0 || std::exp(-delta) > u01<double>(w[0]);
            if (acc) { holder[k] = b; holder[k + 1] = a; }
            if (accepted) accepted[k] = acc ? 1 : 0;
        }
    });
}

}  // extern "C"
