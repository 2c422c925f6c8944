// Wolff single-cluster update as bond activation + lock-free union-find labelling.
//
// Reference: blockUpdate/expandBlock - heisenbergLib.c:310-439, xyLib.c:256-380, isingLib.c:165-236.
// The reference grows ONE cluster from a random seed by a sequential FIFO and pays O(N) per step
// anyway (flag reset, two N-pointer mallocs, full energy recompute).  Here every bond of the lattice
// is activated independently with the reference's probability 1-exp(min(0,corr)) (each bond owns one
// Philox word keyed by its lower-id endpoint and that endpoint's link slot), active bonds are merged
// with an atomicCAS union-find, and only the cluster containing the seed is reflected: the seed's
// cluster of the bond-percolation configuration has exactly the distribution of the FIFO-grown one
// (a bond is tested at most once in either formulation).  The cluster-wide residual energy (the parts
// of J not proportional to n n^T, single-ion anisotropy, field) is reduced over the selected cluster
// and the reflection accepted with min(1,exp(-res)) as in heisenbergLib.c:403-423 - evaluated with the
// FULL move (the reference reads the half move; oracle flag wolffHalfMove=0).
#pragma once
#include "kernels_generic.cuh"

namespace mcg {

struct WolffArgs {
    int32_t *parent;      // [R][N]
    void *proj;           // [R][N] real: a_p = -(s_p . n)
    double *wres;         // [R][2]: residual energy, cluster size
    const int32_t *pos_of;
    uint64_t step;
};

template <int NC, typename real>
__device__ __forceinline__ void wolff_seed(const GenArgs &a, int r, uint64_t step, real (&n)[3], int &seedSite, real &uAcc) {
    uint32_t w[4];
    rng4(a.key, a.replica0 + r, STREAM_WSEED, 0, step, 0u, w);
    seedSite = (int)(((uint64_t)w[3] * (uint64_t)a.N) >> 32);
    uAcc = u01<real>(w[2]);
    if (NC == 1) { n[0] = 1; n[1] = n[2] = 0; }
    else random_dir<NC, real>(w[0], w[1], n);
}

__device__ __forceinline__ int uf_find(int32_t *parent, int x) {
    for (;;) {
        int p = parent[x];
        if (p == x) return x;
        int gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving; racy writes only ever point to an ancestor
        x = p;
    }
}
__device__ __forceinline__ void uf_unite(int32_t *parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        if (atomicCAS(parent + a, a, b) == a) return;   // hook the larger root under the smaller
    }
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_wolff_init(GenArgs a, WolffArgs w) {
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { w.wres[2 * r] = 0.0; w.wres[2 * r + 1] = 0.0; }
    if (p >= a.N) return;
    w.parent[(size_t)r * a.N + p] = p;
    if (NC > 1) {
        real n[3], u; int seed;
        wolff_seed<NC, real>(a, r, w.step, n, seed, u);
        const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
        real s[3];
        load_spin<NC, real>(sp, a.N, p, s);
        ((real *)w.proj)[(size_t)r * a.N + p] = -(s[0] * n[0] + s[1] * n[1] + s[2] * n[2]);   // sDotN, heisenbergLib.c:324
    }
}

template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(256) k_wolff_bonds(GenArgs a, WolffArgs w) {
    constexpr int JW = NC == 1 ? 1 : 9;
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    const real *Jtab = (const real *)a.Jtab;
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    const real *proj = (const real *)w.proj + (size_t)r * a.N;
    int32_t *parent = w.parent + (size_t)r * a.N;
    real beta = (real)a.beta[r];
    real n[3], uAcc; int seed;
    wolff_seed<NC, real>(a, r, w.step, n, seed, uAcc);
    int ip = a.site_of[p];
    real ap = NC == 1 ? sp[p] : proj[p];
    uint32_t wd[4];
    int have = -1;
    for (int k = 0; k < a.maxL; k++) {
        int q = a.nbrp[(size_t)k * a.N + p];
        if (q == p) continue;                         // padding / self image
        if (a.site_of[q] < ip) continue;              // bond owned by the lower reference id
        int jt = a.jtype[(size_t)k * a.N + p];
        const real *J = Jtab + (size_t)jt * JW;
        real corr;
        if (NC == 1) corr = real(2) * beta * J[0] * ap * sp[q];                       // isingLib.c:183-185
        else corr = real(2) * ap * proj[q] * beta * quad_form<NC, real, FULLJ>(J, n, n);   // heisenbergLib.c:355
        if (corr < real(0)) {
            if (have != (k >> 2)) { rng4(a.key, a.replica0 + r, STREAM_WBOND, (uint32_t)(k >> 2), w.step, (uint32_t)ip, wd); have = k >> 2; }
            if ((real(1) - r_exp<real>(corr)) > u01<real>(wd[k & 3])) uf_unite(parent, p, q);
        }
    }
}

__global__ void __launch_bounds__(256) k_wolff_flatten(int N, int32_t *parentAll) {
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    int32_t *parent = parentAll + (size_t)r * N;
    int x = p;
    while (parent[x] != x) x = parent[x];
    parent[p] = x;
}

// residual energy of reflecting the seed's cluster, and its size
template <int NC, typename real, bool FULLJ>
__global__ void __launch_bounds__(256) k_wolff_residual(GenArgs a, WolffArgs w) {
    constexpr int JW = NC == 1 ? 1 : 9;
    __shared__ double smem[2 * 32];
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    real n[3], uAcc; int seed;
    wolff_seed<NC, real>(a, r, w.step, n, seed, uAcc);
    const int32_t *parent = w.parent + (size_t)r * a.N;
    if (p < a.N) {
        int root = parent[w.pos_of[seed]];
        if (parent[p] == root) {
            const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
            real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
            v[1] = 1.0;
            if (NC == 1) {
                v[0] = 2.0 * (double)hf * (double)sp[p];                         // getDeltaOnsiteEnergy isingLib.c:129-131
            } else {
                const real *Jtab = (const real *)a.Jtab;
                const real *proj = (const real *)w.proj + (size_t)r * a.N;
                real s[3];
                load_spin<NC, real>(sp, a.N, p, s);
                real ap = proj[p];
                real perp_p[3] = {s[0] + ap * n[0], s[1] + ap * n[1], s[2] + ap * n[2]};
                double res = 0.0;
                for (int k = 0; k < a.maxL; k++) {
                    int q = a.nbrp[(size_t)k * a.N + p];
                    if (q == p) continue;
                    const real *J = Jtab + (size_t)a.jtype[(size_t)k * a.N + p] * JW;
                    real t[3];
                    load_spin<NC, real>(sp, a.N, q, t);
                    real aq = proj[q];
                    real perp_q[3] = {t[0] + aq * n[0], t[1] + aq * n[1], t[2] + aq * n[2]};
                    real src = ap * beta * quad_form<NC, real, FULLJ>(J, n, perp_q);          // heisenbergLib.c:407
                    res += (double)src;
                    if (parent[q] == root) res += (double)(aq * beta * quad_form<NC, real, FULLJ>(J, perp_p, n));   // :411
                    else res += (double)src;                                                   // :413
                }
                const real *D = (const real *)a.clsD + 3 * a.cls[p];
                real tr[3] = {real(2) * ap * n[0], real(2) * ap * n[1], real(2) * ap * n[2]};
                real t1[3] = {s[0] + tr[0], s[1] + tr[1], s[2] + tr[2]};
                real dOn = D[0] * (t1[0] * t1[0] - s[0] * s[0]) + D[1] * (t1[1] * t1[1] - s[1] * s[1]);
                if (NC == 3) dOn += D[2] * (t1[2] * t1[2] - s[2] * s[2]);
                res += (double)(beta * dOn - hf * (NC == 3 ? tr[2] : tr[0]));
                v[0] = res;
            }
        }
    }
    block_accumulate<2>(v, w.wres + 2 * r, smem);
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_wolff_flip(GenArgs a, WolffArgs w) {
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    real n[3], uAcc; int seed;
    wolff_seed<NC, real>(a, r, w.step, n, seed, uAcc);
    const int32_t *parent = w.parent + (size_t)r * a.N;
    int seedPos = w.pos_of[seed];
    int root = parent[seedPos];
    double res = w.wres[2 * r];
    bool accept = res <= 0.0 || r_exp<real>((real)-res) > uAcc;     // heisenbergLib.c:423 / isingLib.c:225
    if (p == seedPos) {
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_WSTEPS, 1ull);
        if (accept) {
            atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, 1ull);
            atomicAdd(a.cnt + (size_t)r * NCNT + CNT_CLUSTER, (unsigned long long)(w.wres[2 * r + 1] + 0.5));
        }
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, 1ull);
    }
    if (!accept || parent[p] != root) return;
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    if (NC == 1) { sp[p] = -sp[p]; return; }
    real ap = ((const real *)w.proj)[(size_t)r * a.N + p];
    real s[3];
    load_spin<NC, real>(sp, a.N, p, s);
    s[0] += real(2) * ap * n[0]; s[1] += real(2) * ap * n[1]; s[2] += real(2) * ap * n[2];   // heisenbergLib.c:425-426
    if (sizeof(real) == 4) {
        real S = ((const real *)a.clsS)[a.cls[p]];
        real f = S * r_rsqrt<real>(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        s[0] *= f; s[1] *= f; s[2] *= f;
    }
    store_spin<NC, real>(sp, a.N, p, s);
}

}  // namespace mcg
