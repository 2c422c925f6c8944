// Wolff single-cluster update as bond activation + lock-free union-find labelling.
//
// Reference: blockUpdate/expandBlock - heisenbergLib.c:310-439, xyLib.c:256-380, isingLib.c:165-236.
// The reference grows ONE cluster from a random seed by a sequential FIFO and pays O(N) per step
// anyway (flag reset, two N-pointer mallocs, full energy recompute).  Here every bond of the lattice
// is activated independently with the reference's probability 1-exp(min(0,corr)), active bonds are
// merged with an atomicCAS union-find, and only the cluster containing the seed is reflected: the
// seed's cluster of the bond-percolation configuration has exactly the distribution of the FIFO-grown
// one (a bond is tested at most once in either formulation).  The cluster-wide residual energy (the
// parts of J not proportional to n n^T, single-ion anisotropy, field) is reduced over the selected
// cluster and the reflection accepted with min(1,exp(-res)) as in heisenbergLib.c:403-423 - evaluated
// with the FULL move (the reference reads the half move; oracle flag wolffHalfMove=0).
//
// The uniform of a bond is one Philox word keyed by (lower reference site id, higher reference site
// id, occurrence, step, replica): independent of the storage layout and the path, so the table path,
// the structured path and the oracle's FIFO restatement select the same clusters.  "occurrence"
// separates several links between the same two sites (tables built with forceAdd: a dipole link next
// to an exchange bond, Lattice.py:298): each is an independent bond, so the pair activates with
// 1-(1-p1)(1-p2) as in the reference's sequential test (heisenbergLib.c:355-366), not with max(p1,p2).
//
// The kernels are written over a topology policy: TableTopo (neighbour tables of the legacy payload)
// or StructTopo (structured.cu: neighbours computed from the class decomposition).
#pragma once
#include "kernels_generic.cuh"

namespace mcg {

struct WolffArgs {
    int32_t *parent;      // [R][N] union-find forest of THIS step (identity on entry)
    int32_t *parentNext;  // [R][N] the other buffer: reset to identity by this step's flip kernel
    void *proj;           // [R][N] real: a_p = -(s_p . n) for this step's plane normal
    void *projNext;       // [R][N] real: the same for the next step, written by this step's flip kernel
    double *wres;         // [R][2]: residual energy, cluster size
    uint64_t step;
    int N, R;
    void *spin;           // [R][NC][N] real
    const double *beta, *field;
    unsigned long long *cnt;
    RngKey key;
    uint32_t replica0;
    // frontier/global hybrid (large lattices, wolff_launch_hybrid): per-replica state words, visit stamps and member queue.
    // mode == nullptr: the plain sequence, parent/parentNext/proj/projNext are this step's buffers.  Otherwise parent and proj
    // are the bases of [2][R][N] buffer pairs and each replica reads which half is clean from its state words.
    int32_t *mode = nullptr;      // [R][WM_N]
    uint32_t *stamp = nullptr;    // [R][N]: == tag once the site has joined this step's cluster
    int32_t *queue = nullptr;     // [R][cap] storage positions of the members, in the order they joined
    int cap = 0;
    uint32_t tag = 0;             // unique per (system, step), never 0
    int hostPrimed = 0;           // nothing but hybrid Wolff steps touched spins/buffers since the previous step
};

// per-replica state words of the hybrid
enum { WM_PENDING = 0,   // this step still needs the global passes (cluster outgrew the queue, or frontier growth was not tried)
       WM_STREAK,        // consecutive steps whose cluster outgrew the queue
       WM_CLEAN,         // which half of the parent/proj buffer pair holds the identity forest
       WM_TOGGLE,        // the global flip pass of the previous step reset the other half: switch before use
       WM_PROJTAG,       // tag of the step whose projections the clean half holds (O(2)/O(3))
       WM_NFRONT,        // steps completed by frontier growth so far (instrumentation)
       WM_N = 8 };

template <typename real> struct WolffBufs { int32_t *parent, *parentNext; real *proj, *projNext; };
template <typename real> __device__ __forceinline__ WolffBufs<real> wolff_bufs(const WolffArgs &w, int r) {
    if (w.mode) {
        const int b = w.mode[WM_N * r + WM_CLEAN];
        const size_t o0 = ((size_t)b * w.R + r) * w.N, o1 = ((size_t)(1 - b) * w.R + r) * w.N;
        return {w.parent + o0, w.parent + o1, (real *)w.proj + o0, (real *)w.proj + o1};
    }
    const size_t o = (size_t)r * w.N;
    return {w.parent + o, w.parentNext + o, (real *)w.proj + o, (real *)w.projNext + o};
}
__device__ __forceinline__ bool wolff_skip(const WolffArgs &w, int r) { return w.mode && !w.mode[WM_N * r + WM_PENDING]; }

// occ: occurrence index of the bond among the links of its site pair (0 unless a pair is linked more than once: every
// duplicate is an independent bond with its own uniform, word occ & 3 of the block with sub-stream occ >> 2)
__host__ __device__ __forceinline__ uint32_t rng_bond(const RngKey &key, uint32_t replica, uint64_t step, uint32_t lo, uint32_t hi, uint32_t occ = 0) {
    uint32_t out[4];
    philox4x32_10(lo, hi, (STREAM_WBOND << 24) | ((occ >> 2) << 16) | (uint32_t)((step >> 16) & 0xFFFFu),
                  (replica & 0xFFFFu) | ((uint32_t)(step & 0xFFFFu) << 16), key, out);
    return pick4(out, occ);
}

template <int NC, typename real>
__device__ __forceinline__ void wolff_seed(const WolffArgs &w, int r, real (&n)[3], int &seedSite, real &uAcc) {
    uint32_t ww[4];
    rng4(w.key, w.replica0 + r, STREAM_WSEED, 0, w.step, 0u, ww);
    seedSite = (int)(((uint64_t)ww[3] * (uint64_t)w.N) >> 32);
    uAcc = u01<real>(ww[2]);
    if (NC == 1) { n[0] = 1; n[1] = n[2] = 0; }
    else random_dir<NC, real>(ww[0], ww[1], n);
}

// seed site, plane normal and acceptance uniform are the same for every thread of a replica: one thread per
// block draws them (one Philox call instead of one per thread) and shares them through shared memory
template <int NC, typename real> struct SeedShared { real n[3]; real uAcc; int seed; };
template <int NC, typename real>
__device__ __forceinline__ void wolff_seed_block(const WolffArgs &w, int r, SeedShared<NC, real> &sh, real (&n)[3], int &seedSite, real &uAcc) {
    if (threadIdx.x == 0) {
        real nn[3], u; int sd;
        wolff_seed<NC, real>(w, r, nn, sd, u);
        sh.n[0] = nn[0]; sh.n[1] = nn[1]; sh.n[2] = nn[2]; sh.uAcc = u; sh.seed = sd;
    }
    __syncthreads();
    n[0] = sh.n[0]; n[1] = sh.n[1]; n[2] = sh.n[2]; uAcc = sh.uAcc; seedSite = sh.seed;
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

// ---- topology of the table path ----
template <int NC, typename real> struct TableTopo {
    GenArgs a;
    const int32_t *pos_of;
    struct Ctx { int p; };
    __device__ __forceinline__ Ctx begin(int p) const { return Ctx{p}; }
    __device__ __forceinline__ int site_id(const Ctx &c) const { return a.site_of[c.p]; }
    __device__ __forceinline__ int site_id_of(int q) const { return a.site_of[q]; }
    __device__ __forceinline__ int pos_of_site(int id) const { return pos_of[id]; }
    __device__ __forceinline__ int nlinks(const Ctx &) const { return a.maxL; }
    // k-th link: neighbour storage position and exchange tensor; false for padding / self image
    __device__ __forceinline__ bool link(const Ctx &c, int k, int &q, const real *&J) const {
        q = a.nbrp[(size_t)k * a.N + c.p];
        if (q == c.p) return false;
        J = (const real *)a.Jtab + (size_t)a.jtype[(size_t)k * a.N + c.p] * (NC == 1 ? 1 : 9);
        return true;
    }
    // bond ownership: the endpoint with the lower reference id activates the bond; returns the neighbour's id
    __device__ __forceinline__ bool owns(const Ctx &c, int k, int q, int ip, int &iq) const {
        iq = a.site_of[q];
        return iq >= ip;
    }
    __device__ __forceinline__ int nbr_id(const Ctx &, int, int q) const { return a.site_of[q]; }
    // how many earlier link slots of this site lead to the same neighbour (link lists are symmetric: the k-th link of a pair
    // has the same rank in both endpoints' lists, Lattice.py:260-262)
    __device__ __forceinline__ uint32_t occurrence(const Ctx &c, int k, int q) const {
        uint32_t n = 0;
        if (a.dupLinks)
            for (int j = 0; j < k; j++) n += a.nbrp[(size_t)j * a.N + c.p] == q;
        return n;
    }
    __device__ __forceinline__ real S(const Ctx &c) const { return ((const real *)a.clsS)[a.cls[c.p]]; }
    __device__ __forceinline__ void D(const Ctx &c, real (&d)[3]) const {
        const real *D = (const real *)a.clsD + 3 * a.cls[c.p];
        d[0] = D[0]; d[1] = D[1]; d[2] = D[2];
    }
};

// ---- per-site bodies (shared by the per-phase kernels below and by the resident kernel) ----
// All per-site bodies take replica-local pointers: sp [NC][N], proj [N], parent [N] (global memory in the per-phase
// kernels, shared memory in the resident kernel).

// a_p = -(s_p . n) (sDotN, heisenbergLib.c:324).  Explicit fused multiply-adds: the stored projections (global passes) and the
// ones the frontier kernel forms on the fly must be the same bits, whatever the compiler contracts around them.
template <typename real> __device__ __forceinline__ real wolff_proj(const real (&s)[3], const real (&n)[3]) {
    return -fma(s[2], n[2], fma(s[1], n[1], s[0] * n[0]));
}

template <int NC, typename real>
__device__ __forceinline__ void wolff_init_site(const real *sp, int N, int p, const real (&n)[3], int32_t *parent, real *proj) {
    if (parent) parent[p] = p;
    if (NC > 1) {
        real s[3];
        load_spin<NC, real>(sp, N, p, s);
        proj[p] = wolff_proj<real>(s, n);
    }
}

// is the bond between p (projection/Ising spin ap, reference id ip) and its k-th neighbour q (aq, iq) active in this step?
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ bool wolff_bond_active(const WolffArgs &w, int r, real beta, const real *J, const real (&n)[3], real ap, real aq,
                                                  int ip, int iq, uint32_t occ) {
    real corr;
    if (NC == 1) corr = real(2) * beta * J[0] * ap * aq;                           // isingLib.c:183-185
    else corr = real(2) * ap * aq * beta * quad_form<NC, real, FULLJ>(J, n, n);    // heisenbergLib.c:355
    if (!(corr < real(0))) return false;
    const uint32_t wd = rng_bond(w.key, w.replica0 + r, w.step, (uint32_t)min(ip, iq), (uint32_t)max(ip, iq), occ);
    return (real(1) - r_exp<real>(corr)) > u01<real>(wd);
}

template <int NC, typename real, bool FULLJ, typename TOPO>
__device__ __forceinline__ void wolff_bonds_site(const TOPO &topo, const WolffArgs &w, int r, int p, const real (&n)[3], const real *sp,
                                                 const real *proj, int32_t *parent) {
    real beta = (real)w.beta[r];
    auto c = topo.begin(p);
    const int ip = topo.site_id(c);
    real ap = NC == 1 ? sp[p] : proj[p];
    const int nl = topo.nlinks(c);
    for (int k = 0; k < nl; k++) {
        int q; const real *J;
        if (!topo.link(c, k, q, J)) continue;
        int iq;
        if (!topo.owns(c, k, q, ip, iq)) continue;    // every bond is activated by exactly one of its endpoints
        // the occurrence index is only looked up for bonds that can activate (it walks the link list)
        real corr;
        if (NC == 1) corr = real(2) * beta * J[0] * ap * sp[q];
        else corr = real(2) * ap * proj[q] * beta * quad_form<NC, real, FULLJ>(J, n, n);
        if (corr < real(0) &&
            wolff_bond_active<NC, real, FULLJ>(w, r, beta, J, n, ap, NC == 1 ? sp[q] : proj[q], ip, iq, topo.occurrence(c, k, q)))
            uf_unite(parent, p, q);
    }
}

__device__ __forceinline__ void wolff_flatten_site(int32_t *parent, int p) {
    int x = p;
    while (parent[x] != x) x = parent[x];
    parent[p] = x;
}

// residual energy of reflecting the seed's cluster (v[0]) and its size (v[1]) - the contribution of member p.
// projOf(q) = a_q, member(q) = q belongs to the cluster.
template <int NC, typename real, bool FULLJ, typename TOPO, typename PROJ, typename MEMB>
__device__ __forceinline__ void wolff_residual_member(const TOPO &topo, const WolffArgs &w, int r, int p, const real (&n)[3], double (&v)[2],
                                                      const real *sp, PROJ &&projOf, MEMB &&member) {
    real beta = (real)w.beta[r], hf = (real)(w.beta[r] * w.field[r]);
    v[1] += 1.0;
    if (NC == 1) {
        v[0] += 2.0 * (double)hf * (double)sp[p];                         // getDeltaOnsiteEnergy isingLib.c:129-131
        return;
    }
    auto c = topo.begin(p);
    real s[3];
    load_spin<NC, real>(sp, w.N, p, s);
    real ap = projOf(p);
    real perp_p[3] = {s[0] + ap * n[0], s[1] + ap * n[1], s[2] + ap * n[2]};
    double res = 0.0;
    const int nl = topo.nlinks(c);
    for (int k = 0; k < nl; k++) {
        int q; const real *J;
        if (!topo.link(c, k, q, J)) continue;
        real t[3];
        load_spin<NC, real>(sp, w.N, q, t);
        real aq = projOf(q);
        real perp_q[3] = {t[0] + aq * n[0], t[1] + aq * n[1], t[2] + aq * n[2]};
        real src = ap * beta * quad_form<NC, real, FULLJ>(J, n, perp_q);          // heisenbergLib.c:407
        res += (double)src;
        if (member(q)) res += (double)(aq * beta * quad_form<NC, real, FULLJ>(J, perp_p, n));   // :411
        else res += (double)src;                                                   // :413
    }
    real D[3];
    topo.D(c, D);
    real tr[3] = {real(2) * ap * n[0], real(2) * ap * n[1], real(2) * ap * n[2]};
    real t1[3] = {s[0] + tr[0], s[1] + tr[1], s[2] + tr[2]};
    real dOn = D[0] * (t1[0] * t1[0] - s[0] * s[0]) + D[1] * (t1[1] * t1[1] - s[1] * s[1]);
    if (NC == 3) dOn += D[2] * (t1[2] * t1[2] - s[2] * s[2]);
    res += (double)(beta * dOn - hf * (NC == 3 ? tr[2] : tr[0]));
    v[0] += res;
}
// parent[] flattened, root = parent[seed]
template <int NC, typename real, bool FULLJ, typename TOPO>
__device__ __forceinline__ void wolff_residual_site(const TOPO &topo, const WolffArgs &w, int r, int p, const real (&n)[3], int root, double (&v)[2],
                                                    const real *sp, const real *proj, const int32_t *parent) {
    if (parent[p] != root) return;
    wolff_residual_member<NC, real, FULLJ, TOPO>(topo, w, r, p, n, v, sp, [&](int q) { return proj[q]; }, [&](int q) { return parent[q] == root; });
}

// the reflection itself (heisenbergLib.c:425-426); fp32 state is pinned back to |S|
template <int NC, typename real, typename TOPO>
__device__ __forceinline__ void wolff_reflect(const TOPO &topo, int p, real ap, const real (&n)[3], real (&s)[3]) {
    if (NC == 1) s[0] = -s[0];
    else {
        s[0] += real(2) * ap * n[0]; s[1] += real(2) * ap * n[1]; s[2] += real(2) * ap * n[2];
        if (sizeof(real) == 4) {
            real S = topo.S(topo.begin(p));
            real f = S * r_rsqrt<real>(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            s[0] *= f; s[1] *= f; s[2] *= f;
        }
    }
}

// reflect site p if it belongs to the accepted cluster, count the step at the seed, and (parentNext != nullptr) prepare
// the next step: fresh forest in the other buffer, projections on the next plane normal n2
template <int NC, typename real, bool FLAT, typename TOPO>
__device__ __forceinline__ void wolff_flip_site(const TOPO &topo, const WolffArgs &w, int r, int p, const real (&n)[3], const real (&n2)[3],
                                                int seedPos, bool accept, bool inCluster, double clusterSize, real *sp, const real *proj,
                                                int32_t *parentNext, real *projNext) {
    if (p == seedPos) {
        atomicAdd(w.cnt + (size_t)r * NCNT + CNT_WSTEPS, 1ull);
        atomicAdd(w.cnt + (size_t)r * NCNT + CNT_ATTEMPT, 1ull);
        if (accept) {
            atomicAdd(w.cnt + (size_t)r * NCNT + CNT_ACCEPT, 1ull);
            if (FLAT) atomicAdd(w.cnt + (size_t)r * NCNT + CNT_CLUSTER, (unsigned long long)(clusterSize + 0.5));
        }
    }
    real s[3];
    if (parentNext || (accept && inCluster)) load_spin<NC, real>(sp, w.N, p, s);
    if (accept && inCluster) {
        wolff_reflect<NC, real, TOPO>(topo, p, NC == 1 ? real(0) : proj[p], n, s);
        store_spin<NC, real>(sp, w.N, p, s);
    }
    if (parentNext) {
        parentNext[p] = p;
        if (NC > 1) projNext[p] = wolff_proj<real>(s, n2);
    }
}

// ---- one kernel per phase (lattices too large for one thread block) ----
// Every kernel walks its replica's sites with a block-uniform grid-stride loop: the plain sequence launches one thread per
// site, the hybrid a grid of a few blocks per SM whose blocks leave at once when the replica's step is already done.
template <int NC, typename real, typename TOPO>
__global__ void __launch_bounds__(256) k_wolff_init(TOPO topo, WolffArgs w) {
    int r = blockIdx.y;
    if (wolff_skip(w, r)) return;
    __shared__ SeedShared<NC, real> sh;
    real n[3], u; int seed;
    wolff_seed_block<NC, real>(w, r, sh, n, seed, u);
    const WolffBufs<real> b = wolff_bufs<real>(w, r);
    // hybrid: the clean half's forest is the identity already; its projections may be this step's (previous global flip)
    if (w.mode && (NC == 1 || (w.hostPrimed && (uint32_t)w.mode[WM_N * r + WM_PROJTAG] == w.tag))) return;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < w.N; p += gridDim.x * blockDim.x)
        wolff_init_site<NC, real>((const real *)w.spin + (size_t)r * NC * w.N, w.N, p, n, w.mode ? nullptr : b.parent, b.proj);
}

template <int NC, typename real, bool FULLJ, typename TOPO>
__global__ void __launch_bounds__(256) k_wolff_bonds(TOPO topo, WolffArgs w) {
    int r = blockIdx.y;
    if (wolff_skip(w, r)) return;
    __shared__ SeedShared<NC, real> sh;
    real n[3], uAcc; int seed;
    wolff_seed_block<NC, real>(w, r, sh, n, seed, uAcc);
    if (blockIdx.x == 0 && threadIdx.x == 0) { w.wres[2 * r] = 0.0; w.wres[2 * r + 1] = 0.0; }
    const WolffBufs<real> b = wolff_bufs<real>(w, r);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < w.N; p += gridDim.x * blockDim.x)
        wolff_bonds_site<NC, real, FULLJ, TOPO>(topo, w, r, p, n, (const real *)w.spin + (size_t)r * NC * w.N, b.proj, b.parent);
}

static __global__ void __launch_bounds__(256) k_wolff_flatten(WolffArgs w) {
    int r = blockIdx.y;
    if (wolff_skip(w, r)) return;
    int32_t *parent = wolff_bufs<float>(w, r).parent;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < w.N; p += gridDim.x * blockDim.x) wolff_flatten_site(parent, p);
}

template <int NC, typename real, bool FULLJ, typename TOPO>
__global__ void __launch_bounds__(256) k_wolff_residual(TOPO topo, WolffArgs w) {
    __shared__ double smem[2 * 32];
    int r = blockIdx.y;
    if (wolff_skip(w, r)) return;
    double v[2] = {0.0, 0.0};
    __shared__ SeedShared<NC, real> sh;
    real n[3], uAcc; int seed;
    wolff_seed_block<NC, real>(w, r, sh, n, seed, uAcc);
    const WolffBufs<real> b = wolff_bufs<real>(w, r);
    const int root = b.parent[topo.pos_of_site(seed)];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < w.N; p += gridDim.x * blockDim.x)
        wolff_residual_site<NC, real, FULLJ, TOPO>(topo, w, r, p, n, root, v, (const real *)w.spin + (size_t)r * NC * w.N, b.proj, b.parent);
    block_accumulate<2>(v, w.wres + 2 * r, smem);
}

// Reflect the seed's cluster and prepare the next step in the same pass: the OTHER parent buffer is reset to the
// identity and the projections on the next step's plane normal are written, so a step costs two passes over the
// lattice (bonds, flip) when no residual energy has to be reduced first.
//   FLAT = true : parent[] was flattened (residual mode), membership is parent[p] == parent[seed];
//   FLAT = false: membership by find() on the unflattened forest; the cluster is always accepted (residual == 0).
template <int NC, typename real, bool FLAT, typename TOPO>
__global__ void __launch_bounds__(256) k_wolff_flip(TOPO topo, WolffArgs w) {
    int r = blockIdx.y;
    if (wolff_skip(w, r)) return;
    __shared__ SeedShared<NC, real> sh, shNext;
    __shared__ int shSeedPos;
    real n[3], uAcc; int seed;
    wolff_seed_block<NC, real>(w, r, sh, n, seed, uAcc);
    real n2[3] = {0, 0, 0};
    if (NC > 1) {
        WolffArgs wn = w;
        wn.step = w.step + 1;
        real u2; int seed2;
        wolff_seed_block<NC, real>(wn, r, shNext, n2, seed2, u2);
    }
    if (threadIdx.x == 0) shSeedPos = topo.pos_of_site(seed);
    __syncthreads();
    const WolffBufs<real> b = wolff_bufs<real>(w, r);
    int32_t *parent = b.parent;
    const int seedPos = shSeedPos;
    bool accept = true;
    double csize = 0.0;
    int rootSeed = 0;
    if (FLAT) {
        double res = w.wres[2 * r];
        accept = res <= 0.0 || r_exp<real>((real)-res) > uAcc;     // heisenbergLib.c:423 / isingLib.c:225
        csize = w.wres[2 * r + 1];
        rootSeed = parent[seedPos];
    } else {
        rootSeed = uf_find(parent, seedPos);
    }
    int nIn = 0;
    for (int base = blockIdx.x * blockDim.x; base < w.N; base += gridDim.x * blockDim.x) {   // block-uniform trip count
        const int p = base + threadIdx.x;
        if (p >= w.N) continue;
        const bool inCluster = FLAT ? parent[p] == rootSeed : uf_find(parent, p) == rootSeed;
        nIn += inCluster ? 1 : 0;
        wolff_flip_site<NC, real, FLAT, TOPO>(topo, w, r, p, n, n2, seedPos, accept, inCluster, csize, (real *)w.spin + (size_t)r * NC * w.N,
                                              b.proj, b.parentNext, b.projNext);
        if (w.mode && p == seedPos) {   // the other half is clean from here on and holds the next step's projections
            w.mode[WM_N * r + WM_TOGGLE] = 1;
            w.mode[WM_N * r + WM_PROJTAG] = (int32_t)(w.tag + 1u);
        }
    }
    if (!FLAT) {   // cluster size by block counts (the residual kernel counts it in the other mode)
        __shared__ int shIn;
        if (threadIdx.x == 0) shIn = 0;
        __syncthreads();
        nIn = __reduce_add_sync(0xffffffffu, nIn);
        if ((threadIdx.x & 31) == 0 && nIn) atomicAdd(&shIn, nIn);
        __syncthreads();
        if (threadIdx.x == 0 && shIn) atomicAdd(w.cnt + (size_t)r * NCNT + CNT_CLUSTER, (unsigned long long)shIn);
    }
}

// ---- frontier growth (hybrid sequence, first kernel of every step) ----
// One thread block per replica grows the seed's cluster breadth-first: a thread per (frontier site, link) tests the bond
// with the same Philox word the global bond pass would use (rng_bond is keyed by the pair of reference ids), claims the
// neighbour with an atomic exchange on its visit stamp and appends it to the member queue.  Bond states are a function of
// (pair, step) alone, so the grown set IS the seed's cluster of the global bond-percolation pass: identical spins after the
// step whichever path ran.  Work is O(cluster x z) instead of O(N x z): above Tc, where a cluster is a few hundred sites of
// 1.7e7, that is the whole point (isingLib.c:165-236 grows its FIFO the same way but resets O(N) flags per step).
// If the cluster outgrows the queue the step is left to the global passes (mode[WM_PENDING] = 1); replicas whose clusters
// keep outgrowing it try again only every 8th step.
constexpr int WF_THREADS = 512;
template <int NC, typename real, bool FULLJ, typename TOPO>
__global__ void __launch_bounds__(WF_THREADS) k_wolff_frontier(TOPO topo, WolffArgs w, int maxL, int needResidual, int force, int qInSmem) {
    const int r = blockIdx.x, tid = threadIdx.x;
    int32_t *mode = w.mode + WM_N * r;
    __shared__ SeedShared<NC, real> sh;
    __shared__ int s_tail, s_over, s_go;
    __shared__ double red[2 * 32];
    real n[3], uAcc; int seed;
    wolff_seed_block<NC, real>(w, r, sh, n, seed, uAcc);
    if (tid == 0) {
        if (mode[WM_TOGGLE]) { mode[WM_CLEAN] ^= 1; mode[WM_TOGGLE] = 0; }
        const bool go = force == 1 || (force == 0 && (mode[WM_STREAK] < 2 || (w.step & 7) == 0));
        s_go = go ? 1 : 0;
        if (!go) mode[WM_PENDING] = 1;
        s_tail = 1; s_over = 0;
    }
    __syncthreads();
    if (!s_go) return;
    const uint32_t tag = w.tag;
    uint32_t *stamp = w.stamp + (size_t)r * w.N;
    // the member queue lives in shared memory when it fits (written by one thread, read by another a level later: a global
    // queue costs an L2 round trip per level)
    extern __shared__ int32_t wf_smem_queue[];
    int32_t *queue = qInSmem ? wf_smem_queue : w.queue + (size_t)r * w.cap;
    real *sp = (real *)w.spin + (size_t)r * NC * w.N;
    const real beta = (real)w.beta[r];
    const int seedPos = topo.pos_of_site(seed);
    if (tid == 0) { queue[0] = seedPos; stamp[seedPos] = tag; }
    __syncthreads();
    auto projOf = [&](int q) {
        real s[3];
        load_spin<NC, real>(sp, w.N, q, s);
        return NC == 1 ? s[0] : wolff_proj<real>(s, n);
    };
    int head = 0, tail = 1;
    while (head < tail) {
        const int work = (tail - head) * maxL;
        for (int i = tid; i < work; i += WF_THREADS) {
            const int m = i / maxL, k = i - m * maxL;
            const int p = queue[head + m];
            auto c = topo.begin(p);
            if (k >= topo.nlinks(c)) continue;
            int q; const real *J;
            if (!topo.link(c, k, q, J)) continue;
            // stamp and both spins are requested together: one memory round trip per level instead of two
            const uint32_t sq = *(volatile uint32_t *)(stamp + q);
            const real ap = projOf(p), aq = projOf(q);
            if (sq == tag) continue;   // a member already: the bond cannot add anything
            if (wolff_bond_active<NC, real, FULLJ>(w, r, beta, J, n, ap, aq, topo.site_id(c), topo.nbr_id(c, k, q), topo.occurrence(c, k, q))) {
                if (atomicExch(stamp + q, tag) != tag) {
                    const int idx = atomicAdd(&s_tail, 1);
                    if (idx < w.cap) queue[idx] = q; else s_over = 1;
                }
            }
        }
        __syncthreads();
        head = tail;
        tail = min(s_tail, w.cap);
        const int over = s_over;
        __syncthreads();
        if (over) break;
    }
    if (s_over) {
        if (tid == 0) { mode[WM_PENDING] = 1; mode[WM_STREAK] = min(mode[WM_STREAK] + 1, 1 << 20); }
        return;
    }
    const int csize = tail;
    bool accept = true;
    if (needResidual) {
        double v[2] = {0.0, 0.0};
        for (int m = tid; m < csize; m += WF_THREADS)
            wolff_residual_member<NC, real, FULLJ, TOPO>(topo, w, r, queue[m], n, v, sp, projOf,
                                                         [&](int q) { return *(volatile uint32_t *)(stamp + q) == tag; });
        const int lane = tid & 31, wp = tid >> 5;
        double s0 = warp_sum(v[0]);
        if (lane == 0) red[wp] = s0;
        __syncthreads();
        double res = 0.0;
        for (int i = 0; i < WF_THREADS / 32; i++) res += red[i];   // same order in every thread
        accept = res <= 0.0 || r_exp<real>((real)-res) > uAcc;       // heisenbergLib.c:423 / isingLib.c:225
        __syncthreads();
    }
    if (accept)
        for (int m = tid; m < csize; m += WF_THREADS) {
            const int p = queue[m];
            real s[3];
            load_spin<NC, real>(sp, w.N, p, s);
            wolff_reflect<NC, real, TOPO>(topo, p, NC == 1 ? real(0) : wolff_proj<real>(s, n), n, s);
            store_spin<NC, real>(sp, w.N, p, s);
        }
    if (tid == 0) {
        mode[WM_PENDING] = 0;
        mode[WM_STREAK] = 0;
        unsigned long long *cnt = w.cnt + (size_t)r * NCNT;
        atomicAdd(cnt + CNT_WSTEPS, 1ull);
        atomicAdd(cnt + CNT_ATTEMPT, 1ull);
        if (accept) { atomicAdd(cnt + CNT_ACCEPT, 1ull); atomicAdd(cnt + CNT_CLUSTER, (unsigned long long)csize); }
        mode[WM_NFRONT]++;
    }
}

static __global__ void __launch_bounds__(256) k_wolff_identity(int32_t *parent, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) parent[i] = (int32_t)i;   // positions are replica-local: n = N < 2^31
}

// launch sequence of one cluster update, shared by both paths.  primed: the forest/projection buffers of this
// step were prepared by the previous step's flip kernel; needResidual: anisotropic exchange, D or field present.
template <int NC, typename real, bool FJ, typename TOPO>
static int wolff_launch_step(const TOPO &topo, const WolffArgs &w, cudaStream_t stream, bool primed, bool needResidual) {
    dim3 g((unsigned)((w.N + 255) / 256), (unsigned)w.R);
    int launches = 0;
    if (!primed) { k_wolff_init<NC, real, TOPO><<<g, 256, 0, stream>>>(topo, w); launches++; }
    k_wolff_bonds<NC, real, FJ, TOPO><<<g, 256, 0, stream>>>(topo, w);
    if (needResidual) {
        k_wolff_flatten<<<g, 256, 0, stream>>>(w);
        k_wolff_residual<NC, real, FJ, TOPO><<<g, 256, 0, stream>>>(topo, w);
        k_wolff_flip<NC, real, true, TOPO><<<g, 256, 0, stream>>>(topo, w);
        return launches + 4;
    }
    k_wolff_flip<NC, real, false, TOPO><<<g, 256, 0, stream>>>(topo, w);
    return launches + 2;
}

// hybrid sequence of one cluster update (w.mode set): frontier growth, then the global passes for the replicas it left pending
// (their blocks return at once otherwise).  force: 0 adaptive, 1 always try the frontier first, 2 never (global passes only).
template <int NC, typename real, bool FJ, typename TOPO>
static int wolff_launch_hybrid(const TOPO &topo, const WolffArgs &w, cudaStream_t stream, int maxL, bool needResidual, int force, int gridX) {
    dim3 g((unsigned)std::min((w.N + 255) / 256, gridX), (unsigned)w.R);
    {
        const size_t qb = (size_t)w.cap * sizeof(int32_t);
        const bool inSmem = qb <= 160 * 1024;
        auto kern = k_wolff_frontier<NC, real, FJ, TOPO>;
        if (inSmem && qb > 40 * 1024) {
            static int attrDev[64];     // per device: the opt-in is cheap, but not free, at 1e5 steps per second
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev < 0 || dev >= 64 || attrDev[dev] != (int)qb) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qb);
                if (dev >= 0 && dev < 64) attrDev[dev] = (int)qb;
            }
        }
        kern<<<w.R, WF_THREADS, inSmem ? qb : 0, stream>>>(topo, w, maxL, needResidual ? 1 : 0, force, inSmem ? 1 : 0);
    }
    int launches = 1;
    if (NC > 1) { k_wolff_init<NC, real, TOPO><<<g, 256, 0, stream>>>(topo, w); launches++; }
    k_wolff_bonds<NC, real, FJ, TOPO><<<g, 256, 0, stream>>>(topo, w);
    if (needResidual) {
        k_wolff_flatten<<<g, 256, 0, stream>>>(w);
        k_wolff_residual<NC, real, FJ, TOPO><<<g, 256, 0, stream>>>(topo, w);
        k_wolff_flip<NC, real, true, TOPO><<<g, 256, 0, stream>>>(topo, w);
        return launches + 4;
    }
    k_wolff_flip<NC, real, false, TOPO><<<g, 256, 0, stream>>>(topo, w);
    return launches + 2;
}

}  // namespace mcg
