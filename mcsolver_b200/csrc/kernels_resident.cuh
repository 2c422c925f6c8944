// Resident engine for small lattices: the WHOLE MCMainFunction loop in one kernel, one thread block per replica.
//
// What it replaces: heisenbergLib.c:614-620, 661-831 (xyLib.c:556-760, isingLib.c:352-431) run thermalisation and
// nsweep x (updates + measurement) as one sequential loop per (T,H) point.  The reference's own workloads are small
// (samples/: 16x16 ... 32x32x2 sites, 10^5 sweeps): on a GPU such a lattice occupies a fraction of one SM and a
// kernel launch per colour pass / Wolff phase / measurement costs more than the work it starts.  Here a block owns
// a replica: colour passes, union-find phases and measurement reductions are separated by __syncthreads() instead
// of launches, the spins stay in L1/L2, the union-find forest and the raw sums in shared memory (no atomics on the sums:
// the block owns them), and the host sees one launch per 65536 sweeps.  Replicas ((T,H) points)
// run on different SMs.  The per-site arithmetic is the SAME device code the per-phase kernels call
// (metro_site, measure_site, wolff_*_site, rg_*, finalize_replica), with the same Philox counters, so the trajectory
// is identical to the launch-per-phase path; the reductions differ only in summation order.
#pragma once
#include <cooperative_groups.h>

#include "kernels_extra.cuh"
#include "kernels_wolff.cuh"

namespace mcg {

constexpr int RES_THREADS = 512;

struct ResidentPlan {
    int algorithm, model;
    long long thermal;        // updates before the first measurement (Metropolis: sweeps; Wolff: cluster steps)
    long long perSweep;       // updates before every measurement
    long long nsweep;         // measured sweeps in this launch
    long long i0;             // index of the first measured sweep of this launch (frame schedule)
    double pAtt;
    int C;
    const int *colourStart;   // device [C+1], storage positions
    unsigned long long sweep0, step0, meas0;
    int needResidual;
    int spinFrame;
    long long per;
    double *frames;           // device [R][spinFrame][N*(1|3)]
    const int32_t *mi, *mj, *pairs, *tri;
    int nLat, nTri, nG, maxG;
    const int32_t *groups;
    double *gsum, *gacc;
    double rg_ci, rg_cj, rg_cij;
    const double *signS;
    double *acc, *last;
    const int32_t *slot;
};

// dynamic shared memory of a Wolff run: union-find forest [N] int32 + projections [N] real
static inline size_t resident_wolff_smem(int N, size_t realSize) { return (((size_t)N * 4 + 15) & ~(size_t)15) + (size_t)N * realSize; }

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(RES_THREADS, 1) k_resident(GenArgs a, RgArgs g, WolffArgs w, const int32_t *pos_of, ResidentPlan P) {
    __shared__ double smem[NSUM * 32];
    __shared__ double ssum[NSUM], srs[NRS], swres[2];   // this replica's raw sums: one block owns them, no atomics
    __shared__ SeedShared<NC, real> sh;
    __shared__ int shSeedPos;
    extern __shared__ __align__(16) unsigned char res_dyn[];
    const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int N = a.N;
    real *sp = (real *)a.spin + (size_t)r * NC * N;
    const real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
    const TableTopo<NC, real> topo{a, pos_of};
    unsigned long long sweep = P.sweep0, step = P.step0, meas = P.meas0;
    unsigned long long nAtt = 0, nAcc = 0, nClu = 0;
    int32_t *parent = reinterpret_cast<int32_t *>(res_dyn);
    real *proj = reinterpret_cast<real *>(res_dyn + (((size_t)N * 4 + 15) & ~(size_t)15));
    if (tid < NSUM) ssum[tid] = 0.0;
    if (tid < NRS) srs[tid] = 0.0;
    __syncthreads();

    auto metro_sweep = [&](double pAttThis) {
        int att = 0, acc = 0;
        for (int c = 0; c < P.C; c++) {
            const int cb = P.colourStart[c], ce = P.colourStart[c + 1];
            for (int p = cb + tid; p < ce; p += nt) metro_site<NC, real, FULLJ>(a, sp, r, p, sweep, (real)pAttThis, beta, hf, att, acc);
            __syncthreads();
        }
        sweep++;
        nAtt += att; nAcc += acc;
    };

    // one cluster update; forest and projections live in shared memory and are rebuilt every step
    auto wolff_step = [&]() {
        WolffArgs ws = w;
        ws.step = step;
        real n[3], uAcc; int seed;
        wolff_seed_block<NC, real>(ws, r, sh, n, seed, uAcc);
        if (tid == 0) shSeedPos = pos_of[seed];
        for (int p = tid; p < N; p += nt) wolff_init_site<NC, real>(sp, N, p, n, parent, proj);
        __syncthreads();
        for (int p = tid; p < N; p += nt) wolff_bonds_site<NC, real, FULLJ>(topo, ws, r, p, n, sp, proj, parent);
        __syncthreads();
        const int seedPos = shSeedPos;
        const real zero3[3] = {0, 0, 0};
        if (P.needResidual) {
            for (int p = tid; p < N; p += nt) wolff_flatten_site(parent, p);
            __syncthreads();
            const int root = parent[seedPos];
            double v[2] = {0.0, 0.0};
            for (int p = tid; p < N; p += nt) wolff_residual_site<NC, real, FULLJ>(topo, ws, r, p, n, root, v, sp, proj, parent);
            block_reduce_to<2>(v, swres, smem);
            const double res = swres[0], csize = swres[1];
            const bool accept = res <= 0.0 || r_exp<real>((real)-res) > uAcc;     // heisenbergLib.c:423 / isingLib.c:225
            for (int p = tid; p < N; p += nt)
                wolff_flip_site<NC, real, true>(topo, ws, r, p, n, zero3, seedPos, accept, parent[p] == root, csize, sp, proj, nullptr, nullptr);
        } else {
            const int rootSeed = uf_find(parent, seedPos);   // roots are stable once every bond has been united
            for (int p = tid; p < N; p += nt) {
                const bool in = uf_find(parent, p) == rootSeed;
                nClu += in ? 1 : 0;
                wolff_flip_site<NC, real, false>(topo, ws, r, p, n, zero3, seedPos, true, in, 0.0, sp, proj, nullptr, nullptr);
            }
        }
        __syncthreads();
        step++;
    };

    auto measure = [&]() {
        {   // site sums, pair correlation and solid angles in ONE block reduction (SUM_* order)
            double v[NSUM];
#pragma unroll
            for (int i = 0; i < NSUM; i++) v[i] = 0.0;
            double (&v10)[10] = reinterpret_cast<double (&)[10]>(v);
            for (int p = tid; p < N; p += nt) measure_site<NC, real, FULLJ>(a, sp, p, beta, hf, P.mi, P.mj, v10, nullptr, nullptr);
            for (int j = tid; j < P.nLat; j += nt) v[SUM_SIJ] += pair_term<NC, real>(sp, N, P.pairs, j);
            if constexpr (NC == 3)
                for (int t = tid; t < P.nTri; t += nt) v[SUM_AREA] += topo_term<real>(a, sp, P.tri, t);
            block_reduce_to<NSUM>(v, ssum, smem);
        }
        const bool groups = P.nG > 0 && P.model != MCG_ISING;
        if (g.nR > 0) {
            RgArgs gg = g;
            gg.meas = meas;
            for (int row = tid; row < g.nR; row += nt) rg_majority_row<NC, real>(a, gg, P.signS, r, row);
            __syncthreads();
            double v[NRS];
#pragma unroll
            for (int i = 0; i < NRS; i++) v[i] = 0.0;
            const int nidx = g.nR > g.nLat ? g.nR : g.nLat;
            for (int t = tid; t < nidx; t += nt) rg_sums_index<NC, real, FULLJ>(a, gg, r, t, v);
            block_reduce_to<NRS>(v, srs, smem);
        }
        if (groups)
            for (int gI = 0; gI < P.nG; gI++) {   // thread 0 stores the totals to global memory and is the one to read them back
                double v[3] = {0.0, 0.0, 0.0};
                for (int k = tid; k < P.maxG; k += nt) group_member<NC, real>(sp, N, P.groups, (size_t)gI * P.maxG + k, v);
                block_reduce_to<3>(v, P.gsum + ((size_t)r * (P.nG + 1) + gI) * 3, smem);
            }
        if (tid == 0) {
            const int n1 = P.nG + 1;
            double *A = P.acc + (size_t)P.slot[r] * NACC;
            if (g.nR > 0 || groups)
                extra_finalize_replica<false>(P.model, P.nLat, g.nR, P.rg_ci, P.rg_cj, P.rg_cij, P.nG, ssum, srs,
                                              P.gsum ? P.gsum + (size_t)r * n1 * 3 : nullptr, A,
                                              P.gacc ? P.gacc + (size_t)P.slot[r] * (n1 + 1) * n1 : nullptr);
            finalize_replica<false>(P.model, N, P.nLat, ssum, A, P.last + 4 * r);
        }
        if (g.nR > 0 || P.nG > 0) meas++;
        __syncthreads();
    };

    // u: index of the update inside its measurement interval - only the LAST sweep of an interval carries the attempt probability
    // P.pAtt (the remainder of ninterval / N), the others are whole sweeps
    auto update = [&](long long u) {
        if (P.algorithm == MCG_METROPOLIS) metro_sweep(u == P.perSweep - 1 ? P.pAtt : 1.0);
        else wolff_step();
    };

    for (long long u = 0; u < P.thermal; u++) update(P.perSweep > 0 ? u % P.perSweep : 0);
    long long iFrame = P.spinFrame > 0 ? (P.i0 + P.per - 1) / P.per : 0;
    if (iFrame > P.spinFrame) iFrame = P.spinFrame;
    const size_t fsz = (size_t)N * (NC == 1 ? 1 : 3);
    for (long long i = 0; i < P.nsweep; i++) {
        for (long long u = 0; u < P.perSweep; u++) update(u);
        if (P.spinFrame > 0 && (P.i0 + i) % P.per == 0 && iFrame < P.spinFrame) {   // heisenbergLib.c:664-675, capped
            double *dst = P.frames + ((size_t)r * P.spinFrame + iFrame) * fsz;
            for (int p = tid; p < N; p += nt) {
                const size_t id = (size_t)a.site_of[p];
                if (NC == 1) dst[id] = sp[p];
                else { dst[3 * id] = sp[p]; dst[3 * id + 1] = sp[N + p]; dst[3 * id + 2] = NC == 3 ? (double)sp[2 * N + p] : 0.0; }
            }
            iFrame++;
        }
        measure();
    }
    if (nAtt) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, nAtt);
    if (nAcc) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, nAcc);
    if (nClu) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_CLUSTER, nClu);
}

// ---------------------------------------------------------------------------------------------
// Mid-size lattices: the same loop as a COOPERATIVE kernel.  B blocks share a replica (all R*B blocks are co-resident,
// one per SM), phases are separated by grid-wide barriers instead of launches, the union-find forest and the raw sums
// stay in global memory exactly as in the per-phase kernels (atomics; read back through L2 after the barrier).
// Worth it while a pass is shorter than a launch (N up to ~10^5 sites): a grid barrier costs ~2 us, a launch ~5 us
// of host time, and a measured sweep needs 7-9 launches.
// ---------------------------------------------------------------------------------------------
struct CoopPlan {
    int B;                    // blocks per replica
    int primed;               // Wolff buffers of step0 prepared by an earlier flip pass
    int32_t *parentBase;      // [2][R][N]
    char *projBase;           // [2][R][N] real
    double *sums, *rsums;     // [R][NSUM], [R][NRS]
};

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(RES_THREADS, 1) k_resident_coop(GenArgs a, RgArgs g, WolffArgs w, const int32_t *pos_of, ResidentPlan P, CoopPlan Q) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double smem[NSUM * 32];
    __shared__ SeedShared<NC, real> sh, shNext;
    __shared__ int shSeedPos;
    const int r = blockIdx.x / Q.B, bq = blockIdx.x - r * Q.B;
    const int nt = blockDim.x * Q.B, tid = bq * blockDim.x + threadIdx.x;   // thread index / count within the replica
    const int N = a.N;
    real *sp = (real *)a.spin + (size_t)r * NC * N;
    const real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
    const TableTopo<NC, real> topo{a, pos_of};
    unsigned long long sweep = P.sweep0, step = P.step0, meas = P.meas0;
    unsigned long long nAtt = 0, nAcc = 0, nClu = 0;
    bool primed = Q.primed != 0;
    const size_t RN = (size_t)w.R * N;
    const bool leader = bq == 0 && threadIdx.x == 0;

    auto metro_sweep = [&](double pAttThis) {
        int att = 0, acc = 0;
        for (int c = 0; c < P.C; c++) {
            const int cb = P.colourStart[c], ce = P.colourStart[c + 1];
            for (int p = cb + tid; p < ce; p += nt) metro_site<NC, real, FULLJ>(a, sp, r, p, sweep, (real)pAttThis, beta, hf, att, acc);
            grid.sync();
        }
        sweep++;
        nAtt += att; nAcc += acc;
    };

    auto wolff_step = [&]() {
        WolffArgs ws = w;
        ws.step = step;
        const int b = (int)(step & 1);
        ws.parent = Q.parentBase + (size_t)b * RN; ws.parentNext = Q.parentBase + (size_t)(1 - b) * RN;
        ws.proj = Q.projBase + (size_t)b * RN * sizeof(real); ws.projNext = Q.projBase + (size_t)(1 - b) * RN * sizeof(real);
        int32_t *parent = ws.parent + (size_t)r * N, *parentNext = ws.parentNext + (size_t)r * N;
        real *proj = (real *)ws.proj + (size_t)r * N, *projNext = (real *)ws.projNext + (size_t)r * N;
        real n[3], uAcc; int seed;
        wolff_seed_block<NC, real>(ws, r, sh, n, seed, uAcc);
        real n2[3] = {0, 0, 0};
        if (NC > 1) {
            WolffArgs wn = ws;
            wn.step = step + 1;
            real u2; int seed2;
            wolff_seed_block<NC, real>(wn, r, shNext, n2, seed2, u2);
        }
        if (threadIdx.x == 0) shSeedPos = pos_of[seed];
        if (leader) { ws.wres[2 * r] = 0.0; ws.wres[2 * r + 1] = 0.0; }
        if (!primed) {
            for (int p = tid; p < N; p += nt) wolff_init_site<NC, real>(sp, N, p, n, parent, proj);
            primed = true;
        }
        grid.sync();
        for (int p = tid; p < N; p += nt) wolff_bonds_site<NC, real, FULLJ>(topo, ws, r, p, n, sp, proj, parent);
        grid.sync();
        const int seedPos = shSeedPos;
        if (P.needResidual) {
            for (int p = tid; p < N; p += nt) wolff_flatten_site(parent, p);
            grid.sync();
            const int root = __ldcg(parent + seedPos);
            double v[2] = {0.0, 0.0};
            for (int p = tid; p < N; p += nt) wolff_residual_site<NC, real, FULLJ>(topo, ws, r, p, n, root, v, sp, proj, parent);
            block_accumulate<2>(v, ws.wres + 2 * r, smem);
            grid.sync();
            const double res = __ldcg(ws.wres + 2 * r), csize = __ldcg(ws.wres + 2 * r + 1);
            const bool accept = res <= 0.0 || r_exp<real>((real)-res) > uAcc;     // heisenbergLib.c:423 / isingLib.c:225
            for (int p = tid; p < N; p += nt)
                wolff_flip_site<NC, real, true>(topo, ws, r, p, n, n2, seedPos, accept, parent[p] == root, csize, sp, proj, parentNext, projNext);
        } else {
            const int rootSeed = uf_find(parent, seedPos);   // roots are stable once every bond has been united
            for (int p = tid; p < N; p += nt) {
                const bool in = uf_find(parent, p) == rootSeed;
                nClu += in ? 1 : 0;
                wolff_flip_site<NC, real, false>(topo, ws, r, p, n, n2, seedPos, true, in, 0.0, sp, proj, parentNext, projNext);
            }
        }
        grid.sync();
        step++;
    };

    auto measure = [&]() {
        double *sums = Q.sums + (size_t)r * NSUM, *rsums = Q.rsums + (size_t)r * NRS;
        {
            double v[NSUM];
#pragma unroll
            for (int i = 0; i < NSUM; i++) v[i] = 0.0;
            double (&v10)[10] = reinterpret_cast<double (&)[10]>(v);
            for (int p = tid; p < N; p += nt) measure_site<NC, real, FULLJ>(a, sp, p, beta, hf, P.mi, P.mj, v10, nullptr, nullptr);
            for (int j = tid; j < P.nLat; j += nt) v[SUM_SIJ] += pair_term<NC, real>(sp, N, P.pairs, j);
            if constexpr (NC == 3)
                for (int t = tid; t < P.nTri; t += nt) v[SUM_AREA] += topo_term<real>(a, sp, P.tri, t);
            block_accumulate<NSUM>(v, sums, smem);
        }
        const bool groups = P.nG > 0 && P.model != MCG_ISING;
        if (groups)
            for (int gI = 0; gI < P.nG; gI++) {
                double v[3] = {0.0, 0.0, 0.0};
                for (int k = tid; k < P.maxG; k += nt) group_member<NC, real>(sp, N, P.groups, (size_t)gI * P.maxG + k, v);
                block_accumulate<3>(v, P.gsum + ((size_t)r * (P.nG + 1) + gI) * 3, smem);
            }
        if (g.nR > 0) {
            RgArgs gg = g;
            gg.meas = meas;
            for (int row = tid; row < g.nR; row += nt) rg_majority_row<NC, real>(a, gg, P.signS, r, row);
            grid.sync();
            double v[NRS];
#pragma unroll
            for (int i = 0; i < NRS; i++) v[i] = 0.0;
            const int nidx = g.nR > g.nLat ? g.nR : g.nLat;
            for (int t = tid; t < nidx; t += nt) rg_sums_index<NC, real, FULLJ>(a, gg, r, t, v);
            block_accumulate<NRS>(v, rsums, smem);
        }
        grid.sync();
        if (leader) {
            const int n1 = P.nG + 1;
            double *A = P.acc + (size_t)P.slot[r] * NACC;
            if (g.nR > 0 || groups)
                extra_finalize_replica<true>(P.model, P.nLat, g.nR, P.rg_ci, P.rg_cj, P.rg_cij, P.nG, sums, rsums,
                                             P.gsum ? P.gsum + (size_t)r * n1 * 3 : nullptr, A,
                                             P.gacc ? P.gacc + (size_t)P.slot[r] * (n1 + 1) * n1 : nullptr);
            finalize_replica<true>(P.model, N, P.nLat, sums, A, P.last + 4 * r);
        }
        if (g.nR > 0 || P.nG > 0) meas++;
        grid.sync();   // the cleared sums must be in place before the next measurement adds to them
    };

    // u: index of the update inside its measurement interval - only the LAST sweep of an interval carries the attempt probability
    // P.pAtt (the remainder of ninterval / N), the others are whole sweeps
    auto update = [&](long long u) {
        if (P.algorithm == MCG_METROPOLIS) metro_sweep(u == P.perSweep - 1 ? P.pAtt : 1.0);
        else wolff_step();
    };

    for (long long u = 0; u < P.thermal; u++) update(P.perSweep > 0 ? u % P.perSweep : 0);
    long long iFrame = P.spinFrame > 0 ? (P.i0 + P.per - 1) / P.per : 0;
    if (iFrame > P.spinFrame) iFrame = P.spinFrame;
    const size_t fsz = (size_t)N * (NC == 1 ? 1 : 3);
    for (long long i = 0; i < P.nsweep; i++) {
        for (long long u = 0; u < P.perSweep; u++) update(u);
        if (P.spinFrame > 0 && (P.i0 + i) % P.per == 0 && iFrame < P.spinFrame) {   // heisenbergLib.c:664-675, capped
            double *dst = P.frames + ((size_t)r * P.spinFrame + iFrame) * fsz;
            for (int p = tid; p < N; p += nt) {
                const size_t id = (size_t)a.site_of[p];
                if (NC == 1) dst[id] = sp[p];
                else { dst[3 * id] = sp[p]; dst[3 * id + 1] = sp[N + p]; dst[3 * id + 2] = NC == 3 ? (double)sp[2 * N + p] : 0.0; }
            }
            iFrame++;
        }
        measure();
    }
    if (nAtt) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, nAtt);
    if (nAcc) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, nAcc);
    if (nClu) atomicAdd(a.cnt + (size_t)r * NCNT + CNT_CLUSTER, nClu);
}

}  // namespace mcg
