// Per-sweep block-spin ("renormalised lattice") and orbital-group statistics of the table path.
//   block-spin: heisenbergLib.c:255-286, 748-803 / xyLib.c:206-237, 675-732 / isingLib.c:133-163, 388-391
//   groups:     heisenbergLib.c:806-830 / xyLib.c:734-757
// These fill result-tuple slots 11-19 and 28 (Ising: 6, 7).  They are cheap gathers over nR = N/8
// (N/4 in 2D) chosen sites and over the group member lists; fp64 throughout.
#pragma once
#include "kernels_generic.cuh"

namespace mcg {

enum : uint32_t { STREAM_RG = 5 };
enum { RS_I = 0, RS_J = 3, RS_IJ = 6, RS_E = 7, NRS = 8 };

struct RgArgs {
    int nR, nC, nLat;
    const int32_t *rPos;      // [nR] storage position of the chosen site
    const int32_t *rCl;       // [nR][nC] storage positions of the cluster members
    const int32_t *rNbrRow;   // [nR][maxL] row of the k-th doubled-bond neighbour (-1: none)
    const int32_t *rNl;       // [nR] nlink of the chosen site
    const int32_t *pairRowI, *pairRowJ;   // [nLat] row of pair member or -1
    double *ms;               // [R][nR][3] majority spins
    double *rsums;            // [R][NRS]
    uint64_t meas;            // measurement index (tie-break stream)
};

// getMajoritySpin: normalise(sum of the cluster) * S (signed S of the chosen site);
// Ising: sign of the sum times |s_o|, ties broken at random (rand() in the reference, Philox here)
template <int NC, typename real>
__device__ __forceinline__ void rg_majority_row(const GenArgs &a, const RgArgs &g, const double *__restrict__ signS, int r, int row) {
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    double s[3] = {0, 0, 0};
    for (int q = 0; q < g.nC; q++) {
        int p = g.rCl[(size_t)row * g.nC + q];
        s[0] += sp[p];
        if (NC >= 2) s[1] += sp[a.N + p];
        if (NC == 3) s[2] += sp[2 * a.N + p];
    }
    double *out = g.ms + ((size_t)r * g.nR + row) * 3;
    int po = g.rPos[row];
    if (NC == 1) {
        double mag = fabs((double)sp[po]);
        double v;
        if (s[0] > 0) v = mag;
        else if (s[0] < 0) v = -mag;
        else {
            uint32_t w[4];
            rng4(a.key, a.replica0 + r, STREAM_RG, 0, g.meas, (uint32_t)a.site_of[po], w);
            v = u01<double>(w[0]) > 0.5 ? mag : -mag;
        }
        out[0] = v; out[1] = 0; out[2] = 0;
        return;
    }
    double len = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    if (!(len < 1e-5)) { s[0] /= len; s[1] /= len; s[2] /= len; }
    double S = signS[po];
    out[0] = s[0] * S; out[1] = s[1] * S; out[2] = s[2] * S;
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_rg_majority(GenArgs a, RgArgs g, const double *__restrict__ signS) {
    int r = blockIdx.y;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < g.nR) rg_majority_row<NC, real>(a, g, signS, r, row);
}

// coarse-lattice energy (index t = chosen site) and coarse pair statistics (index t = pair), added into v[NRS]
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ void rg_sums_index(const GenArgs &a, const RgArgs &g, int r, int t, double (&v)[NRS]) {
    constexpr int JW = NC == 1 ? 1 : 9;
    const double *ms = g.ms + (size_t)r * g.nR * 3;
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    const double beta = a.beta[r], hf = a.beta[r] * a.field[r];
    if (t < g.nR) {
        int p = g.rPos[t];
        const real *Jtab = (const real *)a.Jtab;
        double m[3] = {ms[3 * t], ms[3 * t + 1], ms[3 * t + 2]};
        double corr = 0;
        for (int k = 0; k < g.rNl[t]; k++) {
            int nr = g.rNbrRow[(size_t)t * a.maxL + k];
            if (nr < 0) continue;
            const real *J = Jtab + (size_t)a.jtype[(size_t)k * a.N + p] * JW;   // J of the ORIGINAL link slot (getCorrEnergy_rnorm)
            double n[3] = {ms[3 * nr], ms[3 * nr + 1], ms[3 * nr + 2]};
            if (NC == 1) corr += (double)J[0] * m[0] * n[0];
            else if (NC == 2) {
                corr += m[0] * n[0] * (double)J[0] + m[1] * n[1] * (double)J[1];
                if (FULLJ) corr += m[0] * n[1] * (double)J[3] + m[1] * n[0] * (double)J[6];
            } else {
                corr += m[0] * n[0] * (double)J[0] + m[1] * n[1] * (double)J[1] + m[2] * n[2] * (double)J[2];
                if (FULLJ)
                    corr += m[0] * n[1] * (double)J[3] + m[0] * n[2] * (double)J[4] + m[1] * n[2] * (double)J[5] +
                            m[1] * n[0] * (double)J[6] + m[2] * n[0] * (double)J[7] + m[2] * n[1] * (double)J[8];
            }
        }
        // on-site part with the UN-renormalised spin of the chosen site (heisenbergLib.c:799, isingLib.c:390)
        double s[3] = {(double)sp[p], NC >= 2 ? (double)sp[a.N + p] : 0.0, NC == 3 ? (double)sp[2 * a.N + p] : 0.0};
        double eo;
        if (NC == 1) eo = -hf * s[0];
        else {
            const real *D = (const real *)a.clsD + 3 * a.cls[p];
            eo = beta * ((double)D[0] * s[0] * s[0] + (double)D[1] * s[1] * s[1] + (NC == 3 ? (double)D[2] * s[2] * s[2] : 0.0)) -
                 hf * (NC == 3 ? s[2] : s[0]);
        }
        v[RS_E] += 0.5 * beta * corr + eo;
    }
    if (t < g.nLat) {
        int ri = g.pairRowI[t], rj = g.pairRowJ[t];
        if (ri >= 0) { v[RS_I] += ms[3 * ri]; v[RS_I + 1] += ms[3 * ri + 1]; v[RS_I + 2] += ms[3 * ri + 2]; }
        if (rj >= 0) { v[RS_J] += ms[3 * rj]; v[RS_J + 1] += ms[3 * rj + 1]; v[RS_J + 2] += ms[3 * rj + 2]; }
        if (ri >= 0 && rj >= 0) v[RS_IJ] += ms[3 * ri] * ms[3 * rj] + ms[3 * ri + 1] * ms[3 * rj + 1] + ms[3 * ri + 2] * ms[3 * rj + 2];
    }
}

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(256) k_rg_sums(GenArgs a, RgArgs g) {
    __shared__ double smem[NRS * 32];
    int r = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    double v[NRS];
#pragma unroll
    for (int i = 0; i < NRS; i++) v[i] = 0.0;
    rg_sums_index<NC, real, FULLJ>(a, g, r, t, v);
    block_accumulate<NRS>(v, g.rsums + (size_t)r * NRS, smem);
}

// group sums: one block per (group, replica)
template <int NC, typename real>
__device__ __forceinline__ void group_member(const real *sp, int N, const int32_t *__restrict__ groups, size_t idx, double (&v)[3]) {
    int p = groups[idx];
    if (p < 0) return;   // -1 padding only at the tail (the reference breaks at the first -1)
    v[0] += sp[p];
    if (NC >= 2) v[1] += sp[N + p];
    if (NC == 3) v[2] += sp[2 * N + p];
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_group_sums(GenArgs a, int nG, int maxG, const int32_t *__restrict__ groups, double *gsum) {
    __shared__ double smem[3 * 32];
    int gI = blockIdx.x, r = blockIdx.y;
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    double v[3] = {0, 0, 0};
    for (int k = threadIdx.x; k < maxG; k += blockDim.x) group_member<NC, real>(sp, a.N, groups, (size_t)gI * maxG + k, v);
    block_accumulate<3>(v, gsum + ((size_t)r * (nG + 1) + gI) * 3, smem);
}

// fold block-spin and group sums of replica r into its accumulators (before the raw sums are cleared); CG as in finalize_replica
template <bool CG>
__device__ __forceinline__ void extra_finalize_replica(int model, int nLat, int nR, double ci, double cj, double cij, int nG, const double *s,
                                                       double *rs, double *g, double *A, double *G) {
    if (nR > 0) {
        double q[NRS];
        for (int i = 0; i < NRS; i++) { q[i] = ld_sum<CG>(rs + i); rs[i] = 0.0; }
        const double ici = 1.0 / ci, icj = 1.0 / cj;
        for (int c = 0; c < 3; c++) {
            A[ACC_SIR + c] += fabs(q[RS_I + c] * ici);
            A[ACC_SJR + c] += fabs(q[RS_J + c] * icj);
        }
        A[ACC_SIJR] += q[RS_IJ] / cij;   // 0/0 = NaN when no pair has both members on the coarse lattice, as in the reference
        double er = q[RS_E] / nR;
        A[ACC_ER] += er;
        A[ACC_E2R] += er * er;
    }
    if (nG > 0 && model != MCG_ISING) {
        int n1 = nG + 1;
        for (int c = 0; c < 3; c++) g[3 * nG + c] = ld_sum<CG>(s + SUM_TOT + c) * (1.0 / (double)nLat);   // last "group" = totSpin/nLat
        for (int x = 0; x < n1; x++)
            for (int y = 0; y < n1; y++) {
                double d = ld_sum<CG>(g + 3 * x) * ld_sum<CG>(g + 3 * y) + ld_sum<CG>(g + 3 * x + 1) * ld_sum<CG>(g + 3 * y + 1) +
                           ld_sum<CG>(g + 3 * x + 2) * ld_sum<CG>(g + 3 * y + 2);
                G[x * n1 + y] += d;
                if (x == y) G[n1 * n1 + x] += d * d;
            }
        for (int i = 0; i < n1 * 3; i++) g[i] = 0.0;
    }
}

static __global__ void k_extra_finalize(int model, int R, int nLat, int nR, double ci, double cj, double cij, int nG, const double *sums,
                                 double *rsums, double *gsum, double *acc, double *gacc, const int32_t *slot) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int n1 = nG + 1;
    extra_finalize_replica<true>(model, nLat, nR, ci, cj, cij, nG, sums + (size_t)r * NSUM, rsums + (size_t)r * NRS,
                           gsum ? gsum + (size_t)r * n1 * 3 : nullptr, acc + (size_t)slot[r] * NACC,
                           gacc ? gacc + (size_t)slot[r] * (n1 + 1) * n1 : nullptr);
}

}  // namespace mcg
