// Table-driven ("generic") kernels: any bond graph handed over the legacy MCMainFunction boundary.
//
// Sites are stored colour-major (storage position p; reference id = site_of[p]); spins are
// structure-of-arrays planes [replica][component][N]; neighbour positions and the index of the
// (deduplicated) exchange tensor are link-major [k][N] so that a warp's loads are coalesced.
// One colour class per launch: no two sites of a class are linked, so the class is updated in
// parallel with exactly the sequential semantics of the reference's localUpdate.
#pragma once
#include "system.hpp"

namespace mcg {

template <int NC, typename real>
__device__ __forceinline__ void load_spin(const real *sp, int N, int p, real (&s)[3]) {
    s[0] = sp[p];
    s[1] = NC >= 2 ? sp[N + p] : real(0);
    s[2] = NC == 3 ? sp[2 * N + p] : real(0);
}
template <int NC, typename real>
__device__ __forceinline__ void store_spin(real *sp, int N, int p, const real (&s)[3]) {
    sp[p] = s[0];
    if (NC >= 2) sp[N + p] = s[1];
    if (NC == 3) sp[2 * N + p] = s[2];
}

// local field  H = sum_k J_k . s_nb(k)   (getCorrEnergy / getDeltaCorrEnergy inner loop,
// heisenbergLib.c:238-247, 288-297).  lowLimit: only neighbours stored below that position.
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ void local_field(const GenArgs &a, const real *sp, int p, real (&H)[3]) {
    constexpr int JW = NC == 1 ? 1 : 9;
    const real *__restrict__ Jtab = (const real *)a.Jtab;
    H[0] = H[1] = H[2] = real(0);
    for (int k = 0; k < a.maxL; k++) {
        int q = a.nbrp[(size_t)k * a.N + p];
        int jt = a.jtype[(size_t)k * a.N + p];
        real t[3];
        load_spin<NC, real>(sp, a.N, q, t);
        add_field<NC, real, FULLJ>(H, Jtab + (size_t)jt * JW, t);
    }
}

// ---------------------------------------------------------------------------------------------
// one colour class of a Metropolis sweep
// ---------------------------------------------------------------------------------------------
// one attempt at storage position p (localUpdate, heisenbergLib.c:441-473 / xyLib.c:382-409 / isingLib.c:238-254)
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ void metro_site(const GenArgs &a, real *sp, int r, int p, uint64_t sweep, real pAtt, real beta, real hf,
                                           int &attempted, int &accepted) {
    real s[3], H[3];
    load_spin<NC, real>(sp, a.N, p, s);
    local_field<NC, real, FULLJ>(a, sp, p, H);
    uint32_t w[4];
    if (NC == 1) {   // one uniform per attempt: word id & 3 of the block shared by four site ids (rng.cuh)
        IsingWords iw;
        iw.get(a.key, a.replica0 + r, sweep, (uint32_t)a.site_of[p], pAtt < real(1), w[2], w[3]);
    } else rng4(a.key, a.replica0 + r, STREAM_METRO, 0, sweep, (uint32_t)a.site_of[p], w);
    if (!(pAtt < real(1)) || u01<real>(w[3]) < pAtt) {
        attempted++;
        if (NC == 1) {
            // isingLib.c:242-252: corr = 2*(sum J s_i s_j - h s_i); flip if corr>=0 or exp(corr)>u
            real corr = real(2) * (beta * s[0] * H[0] - hf * s[0]);
            if (metro_accept<real>(corr, w[2])) {
                sp[p] = -s[0];
                accepted++;
            }
        } else {
            int c = a.cls[p];
            const real *D = (const real *)a.clsD + 3 * c;
            real n[3];
            random_dir<NC, real>(w[0], w[1], n);
            // heisenbergLib.c:451-456: transSpin = -2 (s.n) n ; dE = trans.J.s_nb + onsite difference
            real sn = s[0] * n[0] + s[1] * n[1] + (NC == 3 ? s[2] * n[2] : real(0));
            real s1n = real(-2) * sn;
            real tr[3] = {n[0] * s1n, n[1] * s1n, NC == 3 ? n[2] * s1n : real(0)};
            real dE = tr[0] * H[0] + tr[1] * H[1] + (NC == 3 ? tr[2] * H[2] : real(0));
            real t[3] = {s[0] + tr[0], s[1] + tr[1], s[2] + tr[2]};
            real dOn = D[0] * (t[0] * t[0] - s[0] * s[0]) + D[1] * (t[1] * t[1] - s[1] * s[1]);
            if (NC == 3) dOn += D[2] * (t[2] * t[2] - s[2] * s[2]);
            dE = beta * (dE + dOn) - hf * (NC == 3 ? tr[2] : tr[0]);
            if (metro_accept<real>(-dE, w[2])) {   // heisenbergLib.c:461
                if (sizeof(real) == 4) {
                    // fp32 state: pin |s| = S so rounding cannot random-walk the spin length
                    real S = ((const real *)a.clsS)[c];
                    real f = S * r_rsqrt<real>(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
                    t[0] *= f; t[1] *= f; t[2] *= f;
                }
                store_spin<NC, real>(sp, a.N, p, t);
                accepted++;
            }
        }
    }
}

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(256) k_metro_generic(GenArgs a, int cbeg, int cend, uint64_t sweep, real pAtt) {
    int r = blockIdx.y;
    int p = cbeg + blockIdx.x * blockDim.x + threadIdx.x;
    int attempted = 0, accepted = 0;
    if (p < cend) {
        real *sp = (real *)a.spin + (size_t)r * NC * a.N;
        real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
        metro_site<NC, real, FULLJ>(a, sp, r, p, sweep, pAtt, beta, hf, attempted, accepted);
    }
    int natt = __syncthreads_count(attempted), nacc = __syncthreads_count(accepted);
    if (threadIdx.x == 0) {
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, (unsigned long long)natt);
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, (unsigned long long)nacc);
    }
}

// ---------------------------------------------------------------------------------------------
// measurement pass: total spin, energy, pair-weighted sums (fp64 accumulation)
//   sums[r][SUM_TOT..] += s ; sums[r][SUM_E] += e_bond/2 + e_onsite ; SUM_SI += mi*s ; SUM_SJ += mj*s
// optional per-site energies (parity hook mcg_energy) for replica `siteRep`
// ---------------------------------------------------------------------------------------------
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ void measure_site(const GenArgs &a, const real *sp, int p, real beta, real hf, const int32_t *__restrict__ mi,
                                             const int32_t *__restrict__ mj, double (&v)[10], double *ebond_site, double *eons_site) {
    real s[3], H[3];
    load_spin<NC, real>(sp, a.N, p, s);
    local_field<NC, real, FULLJ>(a, sp, p, H);
    int c = a.cls[p];
    real eb = beta * (s[0] * H[0] + (NC >= 2 ? s[1] * H[1] : real(0)) + (NC == 3 ? s[2] * H[2] : real(0)));
    real eo = onsite_energy<NC, real>(s, (const real *)a.clsD + 3 * c, beta, hf);
    v[0] += s[0]; v[1] += s[1]; v[2] += s[2];
    v[3] += 0.5 * (double)eb + (double)eo;
    double wi = mi[p], wj = mj[p];
    v[4] += wi * s[0]; v[5] += wi * s[1]; v[6] += wi * s[2];
    v[7] += wj * s[0]; v[8] += wj * s[1]; v[9] += wj * s[2];
    if (ebond_site) {
        ebond_site[a.site_of[p]] = (double)eb;
        eons_site[a.site_of[p]] = (double)eo;
    }
}

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(256) k_measure_generic(GenArgs a, const int32_t *__restrict__ mi,
                                                         const int32_t *__restrict__ mj, double *sums, int siteRep,
                                                         double *ebond_site, double *eons_site) {
    __shared__ double smem[10 * 32];
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    double v[10];
#pragma unroll
    for (int i = 0; i < 10; i++) v[i] = 0.0;
    if (p < a.N) {
        const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
        real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
        bool site = ebond_site && r == siteRep;
        measure_site<NC, real, FULLJ>(a, sp, p, beta, hf, mi, mj, v, site ? ebond_site : nullptr, site ? eons_site : nullptr);
    }
    block_accumulate<10>(v, sums + (size_t)r * NSUM, smem);
}

// s_i . s_j of one correlated pair   (heisenbergLib.c:701)
template <int NC, typename real>
__device__ __forceinline__ double pair_term(const real *sp, int N, const int32_t *__restrict__ pairs, int j) {
    real a[3], b[3];
    load_spin<NC, real>(sp, N, pairs[2 * j], a);
    load_spin<NC, real>(sp, N, pairs[2 * j + 1], b);
    return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2];
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_pairs_generic(int N, int nLat, const int32_t *__restrict__ pairs, const void *spin,
                                                       double *sums) {
    __shared__ double smem[32];
    int r = blockIdx.y;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (j < nLat) v[0] = pair_term<NC, real>((const real *)spin + (size_t)r * NC * N, N, pairs, j);
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_SIJ, smem);
}

// signed solid angle of one triangle - calcSignedArea heisenbergLib.c:114-127 (plain atan, +-PI guard)
__device__ __forceinline__ double signed_area(const double (&s1)[3], const double (&s2)[3], const double (&s3)[3], double l1,
                                              double l2, double l3) {
    double s1s2 = (s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2]) / l1 / l2;
    double s2s3 = (s2[0] * s3[0] + s2[1] * s3[1] + s2[2] * s3[2]) / l2 / l3;
    double s3s1 = (s3[0] * s1[0] + s3[1] * s1[1] + s3[2] * s1[2]) / l3 / l1;
    double cx = s2[1] * s3[2] - s2[2] * s3[1], cy = s2[2] * s3[0] - s2[0] * s3[2], cz = s2[0] * s3[1] - s2[1] * s3[0];
    double re = 1 + s1s2 + s2s3 + s3s1;
    double im = (s1[0] * cx + s1[1] * cy + s1[2] * cz) / l1 / l2 / l3;
    if (fabs(re) < 1e-6) return im > 0 ? MCG_REF_PI : -MCG_REF_PI;
    return 2 * atan(im / re);
}

template <typename real>
__device__ __forceinline__ double topo_term(const GenArgs &a, const real *sp, const int32_t *__restrict__ tri, int t) {
    double s[3][3], l[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        int p = tri[3 * t + q];
        s[q][0] = sp[p]; s[q][1] = sp[a.N + p]; s[q][2] = sp[2 * a.N + p];
        l[q] = (double)((const real *)a.clsS)[a.cls[p]];   // spin.len = |S| set once (heisenbergLib.c:165)
    }
    return signed_area(s[0], s[1], s[2], l[0], l[1], l[2]);
}

template <typename real>
__global__ void __launch_bounds__(256) k_topo_generic(GenArgs a, int nTri, const int32_t *__restrict__ tri, double *sums) {
    __shared__ double smem[32];
    int r = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (t < nTri) v[0] = topo_term<real>(a, (const real *)a.spin + (size_t)r * 3 * a.N, tri, t);
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_AREA, smem);
}

// ---------------------------------------------------------------------------------------------
// end-of-sweep bookkeeping: one thread per replica folds the raw sums into the running
// accumulators with the reference's (non-linear) definitions, then clears the sums.
//   O(n):  heisenbergLib.c:677-744 / xyLib.c:605-673      Ising: isingLib.c:395-431
// All projections are linear in the sums:  sum_j n.s_j = n.(sum_j s_j).
// ---------------------------------------------------------------------------------------------
// sg: the replica's raw sums, A: its accumulator row.  CG: the sums were built with atomics by other blocks -> read them
// through L2; the resident kernel hands over sums its own thread 0 wrote (plain loads, possibly shared memory).
template <bool CG> __device__ __forceinline__ double ld_sum(const double *p) { return CG ? __ldcg(p) : *p; }
template <bool CG>
__device__ __forceinline__ void finalize_replica(int model, int N, int nLat, double *sg, double *A, double *last) {
    double s[NSUM];
    for (int i = 0; i < NSUM; i++) { s[i] = ld_sum<CG>(sg + i); sg[i] = 0.0; }
    last[0] = s[SUM_E]; last[1] = s[SUM_TOT]; last[2] = s[SUM_TOT + 1]; last[3] = s[SUM_TOT + 2];
    // one reciprocal per divisor instead of ~25 dependent fp64 divisions (this runs on one thread per replica;
    // x * (1/n) is within one ulp of the reference's x / n)
    const double inl = 1.0 / (double)nLat;
    double E = s[SUM_E];
    double e_avg = E / N;
    A[ACC_E] += e_avg;
    A[ACC_E2] += e_avg * e_avg;
    A[ACC_LASTE] = E;
    double M;
    if (model == MCG_ISING) {
        double si = s[SUM_SI], sj = s[SUM_SJ];
        M = si * inl;                                  // signed (isingLib.c:406)
        A[ACC_SI] += fabs(si) * inl;
        A[ACC_SJ] += fabs(sj) * inl;
        A[ACC_SIJ] += s[SUM_SIJ] * inl;
        A[ACC_STOT] += s[SUM_TOT];
    } else {
        int ax = model == MCG_HEISENBERG ? 2 : 0;      // field axis: z (Heisenberg) / x (XY)
        double t[3] = {s[SUM_TOT], s[SUM_TOT + 1], s[SUM_TOT + 2]};
        double len = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        double d[3] = {t[0], t[1], t[2]};
        if (!(len < 1e-5)) { double il = 1.0 / len; d[0] *= il; d[1] *= il; d[2] *= il; }   // normalize(): heisenbergLib.c:19-25
        const double *si = s + SUM_SI, *sj = s + SUM_SJ;
        A[ACC_SIZ] += (d[0] * si[0] + d[1] * si[1] + d[2] * si[2]) * inl;
        A[ACC_SJZ] += (d[0] * sj[0] + d[1] * sj[1] + d[2] * sj[2]) * inl;
        A[ACC_STZ] += (d[0] * t[0] + d[1] * t[1] + d[2] * t[2]) * inl;
        A[ACC_SIH] += si[ax] * inl;
        A[ACC_SJH] += sj[ax] * inl;
        A[ACC_STH] += t[ax] * inl;
        if (model == MCG_HEISENBERG) M = len * inl;                                      // heisenbergLib.c:726
        else M = sqrt(si[0] * si[0] + si[1] * si[1]) * inl;                               // xyLib.c:654
        for (int c = 0; c < 3; c++) {
            A[ACC_SI + c] += fabs(si[c] * inl);
            A[ACC_SJ + c] += fabs(sj[c] * inl);
        }
        A[ACC_SIJ] += s[SUM_SIJ] * inl;
        A[ACC_Q] += s[SUM_AREA] * (1.0 / (4.0 * MCG_REF_PI));
    }
    A[ACC_M2] += M * M;
    A[ACC_M4] += M * M * M * M;
    A[ACC_MTOT] += M;
    A[ACC_MDOTM] += A[ACC_MTMP] * M;
    A[ACC_MTMP] = M;
    A[ACC_NMEAS] += 1.0;
}

static __global__ void k_finalize_sweep(int model, int R, int N, int nLat, double *sums, double *acc, const int32_t *slot, double *last) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    finalize_replica<true>(model, N, nLat, sums + (size_t)r * NSUM, acc + (size_t)slot[r] * NACC, last + 4 * r);
}

// ---------------------------------------------------------------------------------------------
// initial state - establishLattice heisenbergLib.c:157-172: normalise((S,0,0)+flunc*n)*|S|
// ---------------------------------------------------------------------------------------------
template <int NC, typename real>
__global__ void __launch_bounds__(256) k_init_generic(GenArgs a, const double *__restrict__ signS, double flunc) {
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    double S = signS[p];
    if (NC == 1) { sp[p] = (real)S; return; }
    uint32_t w[4];
    rng4(a.key, a.replica0 + r, STREAM_INIT, 0, 0, (uint32_t)a.site_of[p], w);
    double n[3];
    random_dir<NC, double>(w[0], w[1], n);
    if (sizeof(real) == 4) {   // fp32 engines draw the direction with the fp32 uniforms
        float nf[3];
        random_dir<NC, float>(w[0], w[1], nf);
        n[0] = nf[0]; n[1] = nf[1]; n[2] = nf[2];
    }
    double v[3] = {S + flunc * n[0], flunc * n[1], NC == 3 ? flunc * n[2] : 0.0};
    double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(len < 1e-5)) { v[0] /= len; v[1] /= len; v[2] /= len; }
    double aS = fabs(S);
    real t[3] = {(real)(v[0] * aS), (real)(v[1] * aS), (real)(v[2] * aS)};
    store_spin<NC, real>(sp, a.N, p, t);
}

// frame capture: storage order, `real` -> reference order, double [N][3] (Ising [N])
template <int NC, typename real>
__global__ void __launch_bounds__(256) k_gather_frame(GenArgs a, int r, double *out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    int i = a.site_of[p];
    if (NC == 1) { out[i] = sp[p]; return; }
    out[3 * (size_t)i] = sp[p];
    out[3 * (size_t)i + 1] = sp[a.N + p];
    out[3 * (size_t)i + 2] = NC == 3 ? (double)sp[2 * a.N + p] : 0.0;
}
template <int NC, typename real>
__global__ void __launch_bounds__(256) k_scatter_frame(GenArgs a, int r, const double *in) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    int i = a.site_of[p];
    if (NC == 1) { sp[p] = (real)in[i]; return; }
    sp[p] = (real)in[3 * (size_t)i];
    sp[a.N + p] = (real)in[3 * (size_t)i + 1];
    if (NC == 3) sp[2 * a.N + p] = (real)in[3 * (size_t)i + 2];
}

}  // namespace mcg
