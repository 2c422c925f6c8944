// Structured path: translation-invariant lattices given as bond templates + supercell dims.
//
// What it replaces: the reference expands a parameter file into one Python object per orbital
// (Lattice.py:155-284), flattens that into O(N*maxL*9) Python floats (mcMain.py:150-223) and the
// C engine chases pointers through 264-byte AoS structs (heisenbergLib.c:129-151).  Here the
// lattice stays a descriptor: the engine finds the smallest colouring period (px,py,pz) such that
// colour(x,y,z,o) = f(x%px, y%py, z%pz, o) is a proper colouring of the bond graph, splits the
// lattice into nclass = px*py*pz*norb sublattice "classes", and stores every class as a dense
// [Xd][Yd][Zd] array (Xd=Lx/px ...), structure-of-arrays per spin component.  The neighbour of cell
// (X,Y,Z) of class q through link k is cell (X+cX, Y+cY, Z+cZ) of class q'(q,k): indices are
// computed, not stored; the per-class link table and exchange tensors live in shared memory.
// One colour per launch; classes of a colour are interleaved in block order so that the L2 serves
// the repeated neighbour reads and HBM traffic stays at (own read + own write + neighbour read).
//
// Per-sweep measurements (M, E) are fused into the colour passes: every site's final spin is known
// when its own colour is processed, and a bond's final energy is known when its later-coloured end
// is processed (both ends final), so  E = sum_i [ s'_i . H_i(lower colours) + onsite(s'_i) ].
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <vector>

#include "structured.hpp"
#include "nccl_dl.hpp"
#include "struct_pass.cuh"
#include "ising8.cuh"
#include "topo_pass.cuh"
#include "jit.hpp"
#include "kernels_wolff.cuh"
#include "kernels_extra.cuh"

namespace mcg {

// Block-spin tables (see the kernels below).  RgEntD: one (bond, direction) entry of an orbital's link list in bond order.
struct RgEntD {
    int o2;         // orbital of the coarse neighbour
    int off[3];     // +2d (forward) or -2d (backward), reduced into [0,L), internal axes
};
struct RgZones { int R[3], nz[3]; };   // per axis: coordinates closer than R to either edge are their own zone, the rest is "bulk"

struct StructuredSystem {
    int L[3], p[3], Xd, Yd, Zd, ncellc, nclass, norb, V;
    int nrows;
    bool hasSelf = false, selfPairs = false;
    int pair_s, pair_t, pair_d[3];
    int ncircuit = 0;
    std::vector<int> circuits;           // [ncircuit][3][4] internal axes
    std::vector<int> tvertsHost, ttrisHost;
    std::vector<SClassD> classes;        // sorted by colour
    std::vector<int> classOf;            // [(a*py+b)*pz+c)*norb+o] -> class index
    std::vector<int> colourClassStart;   // [C+1]
    std::vector<SLinkD> linksHost;       // [nclass][MAXLINK]
    std::vector<int> cXs, cYs;           // signed coarse offsets per link (fast path needs |c|<=1)
    std::vector<double> JHost;           // [nclass][MAXLINK][JW]
    std::vector<char> fastOK;            // per colour: pass table fits the __grid_constant__ fast path
    std::vector<std::vector<char>> passTables;   // per colour: PassTable<real> bytes
    std::vector<I8Table> i8Tables;       // per colour: link tables of the int8 Ising pass (precision 8)
    std::vector<void *> jitResolved;     // [colour*2 + partial] -> JitPass* once looked up (nullptr = not yet)
    void *jitTopo = nullptr;             // JitPass* of the specialised topological-charge kernel, once looked up
    SClassD *d_classes = nullptr;
    SLinkD *d_links = nullptr;
    void *d_J = nullptr;
    int *d_classOf = nullptr, *d_circuits = nullptr, *d_tverts = nullptr, *d_ttris = nullptr, *d_gmask = nullptr;
    bool asyncAttrSet[2][2] = {{false, false}, {false, false}};   // k_struct_async<MODE, PARTIAL>: dynamic shared memory opt-in done on this system's device
    AsTab *d_asTab = nullptr;                     // [nclass][PT_MAXL + 1] link tables of the asynchronous pipeline (fp32 full-tensor passes)
    std::vector<int> gmaskHost;
    bool groupInSC = true;
    int nvert = 0;
    double *d_classSums = nullptr;       // [R][nclass][4]
    // block-spin ("renormalised lattice") statistics, opt-in through mcg_lattice_desc.block_spin
    bool rgOn = false;
    int rgN[3] = {0, 0, 0}, nR = 0;      // chosen (all-even) cells per axis, chosen sites
    double rg_ci = 0, rg_cj = 0, rg_cij = 0;
    RgZones rgZ;
    std::vector<RgEntD> rgEntHost;       // [norb][MAXLINK]
    std::vector<int> rgNentHost;         // [norb]
    std::vector<int> rgPermHost;         // [nsig][MAXLINK]: (2*bond + transposed) << 8 | entry giving the k-th coarse neighbour
    std::vector<double> rgJHost;         // [nbond][9] unscaled, reference flat order
    std::vector<double> rgSDHost;        // [norb][4] signed S, D
    RgEntD *d_rgEnt = nullptr;
    int *d_rgNent = nullptr, *d_rgPerm = nullptr;
    double *d_rgJ = nullptr, *d_rgSD = nullptr, *d_ms = nullptr, *d_rsums = nullptr;
    double *d_stage = nullptr;           // [3N] host<->device staging, allocated on demand
    // slab decomposition along X (slab.cu): rows [rowLo, rowHi) are this rank's own, the coarse planes X = 0 and X = Xd - 1 are
    // ghosts; xoff = global x of local x = 0; LxGlobal = extent of the whole lattice.  Not decomposed: 0, nrows, 0, L[0].
    int rowLo = 0, rowHi = 0, xoff = 0, LxGlobal = 0;
    int slabRank = 0, slabWorld = 0;     // slabWorld == 0: not decomposed
    void *slabComm = nullptr;            // ncclComm_t of the slab ring (world > 1)
    cudaStream_t slabStream = nullptr;   // the exchange of a colour's boundary planes runs here, behind the interior rows
    cudaEvent_t slabEvA = nullptr, slabEvB = nullptr;
    void *d_slabBuf = nullptr;           // [4][max elements of one boundary plane set]: send left/right, receive right/left
    size_t slabBufElems = 0;
};

__device__ __forceinline__ void split_period(int v, int p, int ps, int &rem, int &quo) {
    if (ps >= 0) { rem = v & (p - 1); quo = v >> ps; }      // periods are almost always 1, 2 or 4
    else { quo = v / p; rem = v - quo * p; }
}
__device__ __forceinline__ int struct_pos(const StructArgs &a, int x, int y, int z, int o) {
    int ca, cb, cc, X, Y, Z;
    split_period(x, a.px, a.psx, ca, X);
    split_period(y, a.py, a.psy, cb, Y);
    split_period(z, a.pz, a.psz, cc, Z);
    int q = a.classOf[((ca * a.py + cb) * a.pz + cc) * a.norb + o];
    return ((q * a.Xd + X) * a.Yd + Y) * a.Zd + Z;
}
__device__ __forceinline__ int struct_site_id(const StructArgs &a, int p) {
    int q = p / a.ncellc, cell = p - q * a.ncellc;
    int Z = cell % a.Zd, Y = (cell / a.Zd) % a.Yd, X = cell / (a.Zd * a.Yd);
    const SClassD &c = a.classes[q];
    int x = X * a.px + c.a, y = Y * a.py + c.b, z = Z * a.pz + c.c;
    return ((x * a.Ly + y) * a.Lz + z) * a.norb + c.o;
}


// reference id of storage position p in the GLOBAL lattice: a slab's local x is shifted by xoff and wrapped (ghost planes of
// the first and last slab are periodic images); equal to struct_site_id when the lattice is not decomposed
__device__ __forceinline__ int struct_site_gid(const StructArgs &a, int p, int LxGlobal) {
    int q = p / a.ncellc, cell = p - q * a.ncellc;
    int Z = cell % a.Zd, Y = (cell / a.Zd) % a.Yd, X = cell / (a.Zd * a.Yd);
    const SClassD &c = a.classes[q];
    int x = X * a.px + c.a + a.xoff, y = Y * a.py + c.b, z = Z * a.pz + c.c;
    if (x < 0) x += LxGlobal;
    if (x >= LxGlobal) x -= LxGlobal;
    return ((x * a.Ly + y) * a.Lz + z) * a.norb + c.o;
}

// ---- topology of the structured path for the Wolff kernels (kernels_wolff.cuh) ----
template <int NC, typename real> struct StructTopo {
    StructArgs a;
    int sZ, sY, sC;      // log2 of Zd, Yd, ncellc when they are powers of two (else -1): p decodes with shifts
    struct Ctx { int p, q, X, Y, Z, x, y, z; };
    __device__ __forceinline__ Ctx begin(int p) const {
        Ctx c;
        c.p = p;
        int cell;
        if (sC >= 0) { c.q = p >> sC; cell = p & (a.ncellc - 1); } else { c.q = p / a.ncellc; cell = p - c.q * a.ncellc; }
        int t;
        if (sZ >= 0) { c.Z = cell & (a.Zd - 1); t = cell >> sZ; } else { t = cell / a.Zd; c.Z = cell - t * a.Zd; }
        if (sY >= 0) { c.Y = t & (a.Yd - 1); c.X = t >> sY; } else { c.X = t / a.Yd; c.Y = t - c.X * a.Yd; }
        const SClassD &cl = a.classes[c.q];
        c.x = c.X * a.px + cl.a; c.y = c.Y * a.py + cl.b; c.z = c.Z * a.pz + cl.c;
        return c;
    }
    __device__ __forceinline__ int site_id(const Ctx &c) const {
        return ((c.x * a.Ly + c.y) * a.Lz + c.z) * a.norb + a.classes[c.q].o;
    }
    // the bond template's source endpoint activates the bond; the neighbour's reference id follows from the
    // link's cell offset without decoding its storage position
    __device__ __forceinline__ bool owns(const Ctx &c, int k, int q, int ip, int &iq) const {
        const SLinkD &L = a.links[c.q * MAXLINK + k];
        if (!L.fwd) return false;
        int xn = c.x + L.dx; if (xn >= a.Lx) xn -= a.Lx;
        int yn = c.y + L.dy; if (yn >= a.Ly) yn -= a.Ly;
        int zn = c.z + L.dz; if (zn >= a.Lz) zn -= a.Lz;
        iq = ((xn * a.Ly + yn) * a.Lz + zn) * a.norb + L.o2;
        return true;
    }
    __device__ __forceinline__ int nbr_id(const Ctx &c, int k, int) const {
        const SLinkD &L = a.links[c.q * MAXLINK + k];
        int xn = c.x + L.dx; if (xn >= a.Lx) xn -= a.Lx;
        int yn = c.y + L.dy; if (yn >= a.Ly) yn -= a.Ly;
        int zn = c.z + L.dz; if (zn >= a.Lz) zn -= a.Lz;
        return ((xn * a.Ly + yn) * a.Lz + zn) * a.norb + L.o2;
    }
    // bond templates with the same (orbital, offset) are merged when the class tables are built: a pair is linked once
    __device__ __forceinline__ uint32_t occurrence(const Ctx &, int, int) const { return 0u; }
    __device__ __forceinline__ int site_id_of(int q) const { return struct_site_id(a, q); }
    __device__ __forceinline__ int pos_of_site(int id) const {
        int o = id % a.norb, cell = id / a.norb;
        int z = cell % a.Lz, y = (cell / a.Lz) % a.Ly, x = cell / (a.Lz * a.Ly);
        return struct_pos(a, x, y, z, o);
    }
    __device__ __forceinline__ int nlinks(const Ctx &c) const { return a.classes[c.q].nlink; }
    __device__ __forceinline__ bool link(const Ctx &c, int k, int &q, const real *&J) const {
        const SLinkD L = a.links[c.q * MAXLINK + k];
        if (L.self) return false;
        int Xn = c.X + L.cX; if (Xn >= a.Xd) Xn -= a.Xd;
        int Yn = c.Y + L.cY; if (Yn >= a.Yd) Yn -= a.Yd;
        int Zn = c.Z + L.cZ; if (Zn < 0) Zn += a.Zd; if (Zn >= a.Zd) Zn -= a.Zd;
        q = ((L.qn * a.Xd + Xn) * a.Yd + Yn) * a.Zd + Zn;
        J = (const real *)a.J + (size_t)(c.q * MAXLINK + k) * (NC == 1 ? 1 : 9);
        return true;
    }
    __device__ __forceinline__ real S(const Ctx &c) const { return (real)fabs(a.classes[c.q].S); }
    __device__ __forceinline__ void D(const Ctx &c, real (&d)[3]) const {
        const SClassD &cl = a.classes[c.q];
        d[0] = (real)cl.D[0]; d[1] = (real)cl.D[1]; d[2] = (real)cl.D[2];
    }
};

// MODE 0: update only   1: update + fused measurement   2: measurement only (no update)
template <int NC, typename real, bool FULLJ, int MODE, int V>
__global__ void __launch_bounds__(256) k_struct(StructArgs a, int q0, int nqc, int rowsPerBlock, int nrb, uint64_t sweep,
                                                real pAtt) {
    constexpr int JW = NC == 1 ? 1 : 9;
    __shared__ SLinkD sl[MAXLINK];
    __shared__ real sJ[MAXLINK * JW];
    __shared__ SClassD sc;
    __shared__ double red[4 * 32];
    const int bid = blockIdx.x;
    const int j = bid % nqc, tq = bid / nqc, rb = tq % nrb, r = tq / nrb;
    const int q = q0 + j;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) sc = a.classes[q];
    for (int i = tid; i < MAXLINK; i += blockDim.x * blockDim.y) sl[i] = a.links[q * MAXLINK + i];
    for (int i = tid; i < MAXLINK * JW; i += blockDim.x * blockDim.y) sJ[i] = ((const real *)a.J)[(size_t)q * MAXLINK * JW + i];
    __syncthreads();
    const int nlink = sc.nlink, lowmode = sc.lowmode;
    const real S = (real)fabs(sc.S);
    const real D[3] = {(real)sc.D[0], (real)sc.D[1], (real)sc.D[2]};
    const real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    const int idStrideZ = a.pz * a.norb;
    real accM[3] = {0, 0, 0}, accE = 0;
    int natt = 0, nacc = 0;

    const int rowEnd = min(a.rowHi, a.rowLo + (rb + 1) * rowsPerBlock);
    for (int row = a.rowLo + rb * rowsPerBlock + threadIdx.y; row < rowEnd; row += blockDim.y) {
        const int X = row / a.Yd, Y = row - X * a.Yd;
        const int rowBase = ((q * a.Xd + X) * a.Yd + Y) * a.Zd;
        const int x = X * a.px + sc.a + a.xoff, y = Y * a.py + sc.b;   // x: of the global lattice (slab decomposition), for the RNG only
        for (int zc = threadIdx.x; zc < a.Zc; zc += blockDim.x) {
            const int Z0 = zc * V;
            real s[3][V], H[3][V], Hl[3][V];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int v = 0; v < V; v++) { s[c][v] = 0; H[c][v] = 0; Hl[c][v] = 0; }
#pragma unroll
            for (int c = 0; c < NC; c++) vload<real, V>(sp + (size_t)c * a.N + rowBase + Z0, s[c]);
            for (int k = 0; k < nlink; k++) {
                const SLinkD L = sl[k];
                int Xn = X + L.cX; if (Xn >= a.Xd) Xn -= a.Xd;
                int Yn = Y + L.cY; if (Yn >= a.Yd) Yn -= a.Yd;
                const int nb = ((L.qn * a.Xd + Xn) * a.Yd + Yn) * a.Zd;
                real t[3][V];
#pragma unroll
                for (int c = 0; c < NC; c++) load_shifted<real, V>(sp + (size_t)c * a.N + nb, Z0, L.cZ, a.Zd, t[c]);
                const real *Jk = sJ + k * JW;
                const bool lowk = MODE == 1 && lowmode == 2 && L.low;
#pragma unroll
                for (int v = 0; v < V; v++) {
                    real tv[3] = {t[0][v], NC >= 2 ? t[1][v] : real(0), NC == 3 ? t[2][v] : real(0)};
                    real Hv[3] = {0, 0, 0};
                    add_field<NC, real, FULLJ>(Hv, Jk, tv);
                    H[0][v] += Hv[0]; H[1][v] += Hv[1]; H[2][v] += Hv[2];
                    if (lowk) { Hl[0][v] += Hv[0]; Hl[1][v] += Hv[1]; Hl[2][v] += Hv[2]; }
                }
            }
            const uint32_t id0 = (uint32_t)(((x * a.Ly + y) * a.Lz + (Z0 * a.pz + sc.c)) * a.norb + sc.o);
            ItemWords<NC, V> iw;   // same stream convention as pass_body
            if (MODE != 2 && V > 1) iw.begin(a.key, a.replica0 + r, sweep, id0, (uint32_t)idStrideZ, pAtt < real(1));
#pragma unroll
            for (int v = 0; v < V; v++) {
                real sv[3] = {s[0][v], s[1][v], s[2][v]};
                const real Hv[3] = {H[0][v], H[1][v], H[2][v]};
                if (MODE != 2) {
                    uint32_t w[4];
                    if (V > 1) { iw.need(a.key, a.replica0 + r, sweep, v + 1); iw.lane(v, pAtt < real(1), w); }
                    else if (NC == 1) { IsingWords one; one.get(a.key, a.replica0 + r, sweep, id0, pAtt < real(1), w[2], w[3]); }
                    else rng4(a.key, a.replica0 + r, STREAM_METRO, 0, sweep, id0, w);
                    if (!(pAtt < real(1)) || u01<real>(w[3]) < pAtt) {
                        natt++;
                        if (NC == 1) {
                            real corr = real(2) * (beta * sv[0] * Hv[0] - hf * sv[0]);      // isingLib.c:242
                            if (metro_accept<real>(corr, w[2])) { sv[0] = -sv[0]; nacc++; }
                        } else {
                            real n[3];
                            random_dir<NC, real>(w[0], w[1], n);
                            real s1n = real(-2) * (sv[0] * n[0] + sv[1] * n[1] + (NC == 3 ? sv[2] * n[2] : real(0)));
                            real tr[3] = {n[0] * s1n, n[1] * s1n, NC == 3 ? n[2] * s1n : real(0)};
                            real dE = tr[0] * Hv[0] + tr[1] * Hv[1] + (NC == 3 ? tr[2] * Hv[2] : real(0));
                            real t1[3] = {sv[0] + tr[0], sv[1] + tr[1], sv[2] + tr[2]};
                            real dOn = D[0] * (t1[0] * t1[0] - sv[0] * sv[0]) + D[1] * (t1[1] * t1[1] - sv[1] * sv[1]);
                            if (NC == 3) dOn += D[2] * (t1[2] * t1[2] - sv[2] * sv[2]);
                            dE = beta * (dE + dOn) - hf * (NC == 3 ? tr[2] : tr[0]);
                            if (metro_accept<real>(-dE, w[2])) {       // heisenbergLib.c:461
                                if (sizeof(real) == 4) {
                                    real f = S * r_rsqrt<real>(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                                    t1[0] *= f; t1[1] *= f; t1[2] *= f;
                                }
                                sv[0] = t1[0]; sv[1] = t1[1]; sv[2] = t1[2];
                                nacc++;
                            }
                        }
                    }
                    s[0][v] = sv[0]; s[1][v] = sv[1]; s[2][v] = sv[2];
                }
                if (MODE != 0) {
                    accM[0] += sv[0]; accM[1] += sv[1]; accM[2] += sv[2];
                    real eb;
                    if (MODE == 2) {
                        eb = real(0.5) * (sv[0] * Hv[0] + sv[1] * Hv[1] + sv[2] * Hv[2]);
                    } else if (lowmode == 1) {
                        eb = sv[0] * Hv[0] + sv[1] * Hv[1] + sv[2] * Hv[2];
                    } else if (lowmode == 2) {
                        eb = sv[0] * Hl[0][v] + sv[1] * Hl[1][v] + sv[2] * Hl[2][v];
                    } else {
                        eb = real(0);
                    }
                    accE += beta * eb + onsite_energy<NC, real>(sv, D, beta, hf);
                }
            }
            if (MODE != 2) {
#pragma unroll
                for (int c = 0; c < NC; c++) vstore<real, V>(sp + (size_t)c * a.N + rowBase + Z0, s[c]);
            }
        }
    }
    if (MODE != 2) {
        natt = __reduce_add_sync(0xffffffffu, natt);
        nacc = __reduce_add_sync(0xffffffffu, nacc);
        if ((tid & 31) == 0 && natt) {
            atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, (unsigned long long)natt);
            atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, (unsigned long long)nacc);
        }
    }
    if (MODE != 0) {
        double v[4] = {(double)accM[0], (double)accM[1], (double)accM[2], (double)accE};
        // block_accumulate indexes warps by threadIdx.x: flatten first
        int lane = tid & 31, w = tid >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double sum = warp_sum(v[i]);
            if (lane == 0) red[i * 32 + w] = sum;
        }
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double sum = lane < nw ? red[i * 32 + lane] : 0.0;
                sum = warp_sum(sum);
                if (lane == 0 && sum != 0.0) atomicAdd(a.classSums + ((size_t)r * a.nclass + q) * 4 + i, sum);
            }
        }
    }
}


// classSums -> the raw per-sweep sums the common finalize kernel consumes; clears classSums
__global__ void __launch_bounds__(32) k_struct_fold(StructArgs a, int R, int NC, int prec, int pair_s, int pair_t, int selfPairs, double nLat, double *sums, int nG,
                              const int *__restrict__ gmask, int groupInSC, double *gacc, const int32_t *slot) {
    // one warp per replica, lanes over the classes (a dipole stencil has 64 of them; one thread walking them through
    // read-modify-writes of global memory took 85 us per sweep)
    const int r = blockIdx.x, lane = threadIdx.x;
    if (r >= R) return;
    double v[11];   // tot[3], si[3], sj[3], E, sij
#pragma unroll
    for (int i = 0; i < 11; i++) v[i] = 0.0;
    for (int q = lane; q < a.nclass; q += 32) {
        const double *c = a.classSums + ((size_t)r * a.nclass + q) * 4;
        const SClassD &cl = a.classes[q];
        const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
        v[0] += c0; v[1] += c1; v[2] += c2;
        if (cl.o == pair_s) { v[3] += c0; v[4] += c1; v[5] += c2; }
        if (cl.o == pair_t) { v[6] += c0; v[7] += c1; v[8] += c2; }
        v[9] += c3;
        if (selfPairs && cl.o == pair_s) v[10] += cl.S * cl.S * (nLat / (a.nclass / a.norb));
    }
#pragma unroll
    for (int i = 0; i < 11; i++) v[i] = warp_sum(v[i]);
    double *s = sums + (size_t)r * NSUM;
    if (lane == 0) {
        for (int k = 0; k < 3; k++) { s[SUM_TOT + k] += v[k]; s[SUM_SI + k] += v[3 + k]; s[SUM_SJ + k] += v[6 + k]; }
        s[SUM_E] += v[9];
        s[SUM_SIJ] += v[10];
    }
    // orbital-group statistics (heisenbergLib.c:806-830): group sums are sums of class sums ("Supergroup")
    // or the listed orbitals of cell (0,0,0); the last "group" is totSpin/nLat
    if (nG > 0 && lane == 0) {
        const int n1 = nG + 1;
        double *G = gacc + (size_t)slot[r] * (n1 + 1) * n1;
        double g[17][3];
        for (int x = 0; x < nG; x++) {
            g[x][0] = g[x][1] = g[x][2] = 0.0;
            if (groupInSC) {
                for (int q = 0; q < a.nclass; q++)
                    if (gmask[x * a.norb + a.classes[q].o]) {
                        const double *c = a.classSums + ((size_t)r * a.nclass + q) * 4;
                        g[x][0] += c[0]; g[x][1] += c[1]; g[x][2] += c[2];
                    }
            } else {
                for (int o = 0; o < a.norb; o++)
                    if (gmask[x * a.norb + o]) {
                        int p = struct_pos(a, 0, 0, 0, o);
                        for (int k = 0; k < NC; k++)
                            g[x][k] += prec == 32 ? (double)((const float *)a.spin)[((size_t)r * NC + k) * a.N + p]
                                                  : ((const double *)a.spin)[((size_t)r * NC + k) * a.N + p];
                    }
            }
        }
        for (int k = 0; k < 3; k++) g[nG][k] = s[SUM_TOT + k] / nLat;
        for (int x = 0; x < n1; x++)
            for (int y = 0; y < n1; y++) {
                double d = g[x][0] * g[y][0] + g[x][1] * g[y][1] + g[x][2] * g[y][2];
                G[x * n1 + y] += d;
                if (x == y) G[n1 * n1 + x] += d * d;
            }
    }
    __syncwarp();
    for (int q = lane; q < a.nclass; q += 32) {
        double *c = a.classSums + ((size_t)r * a.nclass + q) * 4;
        c[0] = c[1] = c[2] = c[3] = 0.0;
    }
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_struct_pairs(StructArgs a, int ps, int pt, int d0, int d1, int d2, double *sums) {
    __shared__ double smem[32];
    int r = blockIdx.y;
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    int ncell = a.Lx * a.Ly * a.Lz;
    if (cell < ncell) {
        int z = cell % a.Lz, y = (cell / a.Lz) % a.Ly, x = cell / (a.Lz * a.Ly);
        int pi = struct_pos(a, x, y, z, ps);
        int pj = struct_pos(a, (x + d0 + a.Lx) % a.Lx, (y + d1 + a.Ly) % a.Ly, (z + d2 + a.Lz) % a.Lz, pt);
        const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
        for (int c = 0; c < NC; c++) v[0] += (double)sp[(size_t)c * a.N + pi] * (double)sp[(size_t)c * a.N + pj];
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_SIJ, smem);
}

// signed_area(): defined in kernels_generic.cuh for the table path; same formula here
__device__ __forceinline__ double signed_area_s(const double (&s1)[3], const double (&s2)[3], const double (&s3)[3], double l1,
                                                double l2, double l3) {
    double s1s2 = (s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2]) / l1 / l2;
    double s2s3 = (s2[0] * s3[0] + s2[1] * s3[1] + s2[2] * s3[2]) / l2 / l3;
    double s3s1 = (s3[0] * s1[0] + s3[1] * s1[1] + s3[2] * s1[2]) / l3 / l1;
    double cx = s2[1] * s3[2] - s2[2] * s3[1], cy = s2[2] * s3[0] - s2[0] * s3[2], cz = s2[0] * s3[1] - s2[1] * s3[0];
    double re = 1 + s1s2 + s2s3 + s3s1;
    double im = (s1[0] * cx + s1[1] * cy + s1[2] * cz) / l1 / l2 / l3;
    if (fabs(re) < 1e-6) return im > 0 ? MCG_REF_PI : -MCG_REF_PI;   // heisenbergLib.c:122-125
    return 2 * atan(im / re);
}

template <typename real>
__global__ void __launch_bounds__(256) k_struct_topo(StructArgs a, int ncirc, const int *__restrict__ circ, double *sums) {
    __shared__ double smem[32];
    int r = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    int ncell = a.Lx * a.Ly * a.Lz;
    if (t < ncell * ncirc) {
        int cell = t / ncirc, ic = t - cell * ncirc;
        int z = cell % a.Lz, y = (cell / a.Lz) % a.Ly, x = cell / (a.Lz * a.Ly);
        const real *sp = (const real *)a.spin + (size_t)r * 3 * a.N;
        double s[3][3], l[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int *e = circ + (ic * 3 + k) * 4;
            int xx = (x + e[1]) % a.Lx, yy = (y + e[2]) % a.Ly, zz = (z + e[3]) % a.Lz;
            int p = struct_pos(a, xx, yy, zz, e[0]);
            s[k][0] = sp[p]; s[k][1] = sp[(size_t)a.N + p]; s[k][2] = sp[2 * (size_t)a.N + p];
            l[k] = fabs(a.classes[p / a.ncellc].S);
        }
        v[0] = signed_area_s(s[0], s[1], s[2], l[0], l[1], l[2]);
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_AREA, smem);
}


constexpr int TOPO_MAXV = 12;
constexpr int TOPO_THREADS = 128;
// The vertex table is indexed by runtime triangle indices, so it lives in shared memory ([vertex][component][thread],
// conflict-free) rather than in a register array the compiler would have to demote to local memory.
// A block handles cells of ONE sublattice parity (x%px, y%py, z%pz) and one coarse X: every vertex of its cells then lies
// in a block-uniform class at a block-uniform coarse offset, so a vertex address costs three wrap-adds (the scheme of the
// colour passes) instead of a full lattice -> storage decode.  Thread (tz, ty) of block (zc + nzc*(yc + nyc*parity), X, r)
// handles coarse cell (X, yc*TY + ty, zc*TZ + tz).
template <typename real>
__global__ void __launch_bounds__(TOPO_THREADS) k_struct_topo_cells(StructArgs a, int ncirc, int nvert, const int *__restrict__ verts,
                                                                    const int *__restrict__ tris, int nzc, int nyc, double *sums) {
    __shared__ double smem[32];
    __shared__ real ilen[TOPO_MAXV], lsh[TOPO_MAXV];
    __shared__ int tsh[3 * 64], vbase[TOPO_MAXV], vcX[TOPO_MAXV], vcY[TOPO_MAXV], vcZ[TOPO_MAXV];
    extern __shared__ __align__(16) unsigned char topo_dyn[];
    real *sv = reinterpret_cast<real *>(topo_dyn);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int par = blockIdx.x / (nzc * nyc), rest = blockIdx.x - par * (nzc * nyc);
    const int yc = rest / nzc, zc = rest - yc * nzc;
    if (tid < nvert) {
        const int *e = verts + 4 * tid;   // (orbital, dx, dy, dz) with offsets already reduced into [0, L)
        const int cc = par % a.pz, cb = (par / a.pz) % a.py, ca = par / (a.pz * a.py);
        const int na = ca + e[1], nb = cb + e[2], nc = cc + e[3];
        const int q = a.classOf[(((na % a.px) * a.py + nb % a.py) * a.pz + nc % a.pz) * a.norb + e[0]];
        vbase[tid] = q * a.ncellc;
        vcX[tid] = (na / a.px) % a.Xd; vcY[tid] = (nb / a.py) % a.Yd; vcZ[tid] = (nc / a.pz) % a.Zd;
        lsh[tid] = (real)fabs(a.classes[q].S);   // |S| depends on the orbital only
        ilen[tid] = real(1) / lsh[tid];
    }
    for (int i = tid; i < 3 * ncirc; i += TOPO_THREADS) tsh[i] = tris[i] * 3 * TOPO_THREADS;
    __syncthreads();
    const int r = blockIdx.z;
    const int X = blockIdx.y, Y = yc * blockDim.y + threadIdx.y, Z = zc * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (Y < a.Yd && Z < a.Zd) {
        const real *sp = (const real *)a.spin + (size_t)r * 3 * a.N;
        for (int k = 0; k < nvert; k++) {
            int Xn = X + vcX[k]; if (Xn >= a.Xd) Xn -= a.Xd;
            int Yn = Y + vcY[k]; if (Yn >= a.Yd) Yn -= a.Yd;
            int Zn = Z + vcZ[k]; if (Zn >= a.Zd) Zn -= a.Zd;
            const real *q = sp + (vbase[k] + (Xn * a.Yd + Yn) * a.Zd + Zn);
            real s0 = q[0], s1 = q[a.N], s2 = q[2 * (size_t)a.N];
            if (sizeof(real) == 4) { s0 *= ilen[k]; s1 *= ilen[k]; s2 *= ilen[k]; }   // fp32: unit vectors once per vertex
            real *d = sv + (size_t)k * 3 * TOPO_THREADS + tid;
            d[0] = s0; d[TOPO_THREADS] = s1; d[2 * TOPO_THREADS] = s2;
        }
        double acc = 0.0;
        float accf = 0.f;
        for (int t = 0; t < ncirc; t++) {
            const real *p0 = sv + tsh[3 * t] + tid, *p1 = sv + tsh[3 * t + 1] + tid, *p2 = sv + tsh[3 * t + 2] + tid;
            real s0[3] = {p0[0], p0[TOPO_THREADS], p0[2 * TOPO_THREADS]};
            real s1[3] = {p1[0], p1[TOPO_THREADS], p1[2 * TOPO_THREADS]};
            real s2[3] = {p2[0], p2[TOPO_THREADS], p2[2 * TOPO_THREADS]};
            if (sizeof(real) == 4) accf += (float)tri_area_unit<real>(s0, s1, s2);   // a cell's few triangles: fp32 partial sum
            else {
                int i0 = tsh[3 * t] / (3 * TOPO_THREADS), i1 = tsh[3 * t + 1] / (3 * TOPO_THREADS), i2 = tsh[3 * t + 2] / (3 * TOPO_THREADS);
                acc += (double)tri_area<real>(s0, s1, s2, lsh[i0], lsh[i1], lsh[i2]);
            }
        }
        v[0] = acc + (double)accf;
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_AREA, smem);
}

// ---------------------------------------------------------------------------------------------
// Block-spin ("renormalised lattice") statistics on the structured path: tuple slots 11-19 (Ising 6, 7).
//   chosen sites   = every orbital of the cells with all-even coordinates            Lattice.py:185-196
//   cluster        = the same orbital in the 2x2x2 cells above it                    Lattice.py:198-205
//   coarse links   = the bond templates with doubled cell offsets, in insertion order Lattice.py:265-271, 60-66
//   coarse energy  = 1/2 sum_k m.J_k.m_k + onsite(un-renormalised spin), where J_k is the exchange of the
//                    k-th ORIGINAL link of the site (getCorrEnergy_rnorm, heisenbergLib.c:255-286)
// Both link lists are insertion-ordered by (id of the bond's source site, bond index): a forward entry is made
// when the site itself is visited, a backward entry when the source site is.  Which comes first depends on the
// periodic wrap, so the ranks are evaluated per site from the keys instead of being tabulated.
// ---------------------------------------------------------------------------------------------
struct RgS {
    int nR, n0, n1, n2, nLat, fullJ;
    RgZones Z;
    const RgEntD *ent;
    const int *nent, *perm;
    const double *J, *SD;
    double *ms, *rsums;
    unsigned long long meas;
    int ps, pt, pd0, pd1, pd2;
};
__host__ __device__ __forceinline__ int rg_zone(int x, int L, int R, int nz) {
    if (nz == L) return x;                       // small axis: every coordinate is its own zone
    return x < R ? x : (x >= L - R ? x - (L - R) + R + 1 : R);
}
__device__ __forceinline__ int wrap_add(int v, int d, int L) { v += d; return v >= L ? v - L : v; }
__device__ __forceinline__ int rg_row(const StructArgs &a, const RgS &g, int x, int y, int z, int o) {
    return (((x >> 1) * g.n1 + (y >> 1)) * g.n2 + (z >> 1)) * a.norb + o;
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_struct_rg_majority(StructArgs a, RgS g) {
    int r = blockIdx.y;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= g.nR) return;
    int o = row % a.norb, c = row / a.norb;
    int z = 2 * (c % g.n2), y = 2 * ((c / g.n2) % g.n1), x = 2 * (c / (g.n2 * g.n1));
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    double s[3] = {0, 0, 0};
    const int offs[8] = {0, 1, 2, 4, 3, 5, 6, 7};   // bit2 = x, bit1 = y, bit0 = z: the member order of Lattice.py:198-205
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int dx = (offs[k] >> 2) & 1, dy = (offs[k] >> 1) & 1, dz = offs[k] & 1;
        if ((dx && a.Lx == 1) || (dy && a.Ly == 1) || (dz && a.Lz == 1)) continue;   // addOrbIntoCluster skips repeated members
        int p = struct_pos(a, wrap_add(x, dx, a.Lx), wrap_add(y, dy, a.Ly), wrap_add(z, dz, a.Lz), o);
        s[0] += sp[p];
        if (NC >= 2) s[1] += sp[(size_t)a.N + p];
        if (NC == 3) s[2] += sp[2 * (size_t)a.N + p];
    }
    double *out = g.ms + ((size_t)r * g.nR + row) * 3;
    if (NC == 1) {
        int po = struct_pos(a, x, y, z, o);
        double mag = fabs((double)sp[po]), v;
        if (s[0] > 0) v = mag;
        else if (s[0] < 0) v = -mag;
        else {
            uint32_t w[4];
            rng4(a.key, a.replica0 + r, STREAM_RG, 0, g.meas, (uint32_t)(((x * a.Ly + y) * a.Lz + z) * a.norb + o), w);
            v = u01<double>(w[0]) > 0.5 ? mag : -mag;
        }
        out[0] = v; out[1] = 0; out[2] = 0;
        return;
    }
    double len = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    if (!(len < 1e-5)) { double il = 1.0 / len; s[0] *= il; s[1] *= il; s[2] *= il; }
    double S = g.SD[4 * o];
    out[0] = s[0] * S; out[1] = s[1] * S; out[2] = s[2] * S;
}

template <int NC, typename real>
__global__ void __launch_bounds__(128) k_struct_rg_sums(StructArgs a, RgS g, int evenShift) {
    __shared__ double smem[NRS * 32];
    int r = blockIdx.y;
    double v[NRS];
#pragma unroll
    for (int i = 0; i < NRS; i++) v[i] = 0.0;
    const double *ms = g.ms + (size_t)r * g.nR * 3;
    const real *sp = (const real *)a.spin + (size_t)r * NC * a.N;
    const double beta = a.beta[r], hf = a.beta[r] * a.field[r];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.nR; t += gridDim.x * blockDim.x) {   // one chosen site per iteration
        int o = t % a.norb, c = t / a.norb;
        int z = 2 * (c % g.n2), y = 2 * ((c / g.n2) % g.n1), x = 2 * (c / (g.n2 * g.n1));
        // the insertion orders of the two link lists depend on which images wrap: one precomputed permutation per edge zone
        int sig = ((rg_zone(x, a.Lx, g.Z.R[0], g.Z.nz[0]) * g.Z.nz[1] + rg_zone(y, a.Ly, g.Z.R[1], g.Z.nz[1])) * g.Z.nz[2] +
                   rg_zone(z, a.Lz, g.Z.R[2], g.Z.nz[2])) * a.norb + o;
        const int *perm = g.perm + (size_t)sig * MAXLINK;
        const RgEntD *ent = g.ent + (size_t)o * MAXLINK;
        const int n = g.nent[o];
        double m[3] = {ms[3 * t], ms[3 * t + 1], ms[3 * t + 2]};
        double corr = 0;
        for (int k = 0; k < n; k++) {
            const int pk = perm[k];
            const double *Jb = g.J + (size_t)(pk >> 9) * 9;
            const bool tr = (pk >> 8) & 1;   // J^T on the target's list (Lattice.py:136-137)
            const RgEntD E = ent[pk & 255];
            int nr = rg_row(a, g, wrap_add(x, E.off[0], a.Lx), wrap_add(y, E.off[1], a.Ly), wrap_add(z, E.off[2], a.Lz), E.o2);
            double nn[3] = {ms[3 * nr], ms[3 * nr + 1], ms[3 * nr + 2]};
            if (NC == 1) corr += Jb[0] * m[0] * nn[0];
            else if (NC == 2) {
                corr += m[0] * nn[0] * Jb[0] + m[1] * nn[1] * Jb[1];
                if (g.fullJ) corr += m[0] * nn[1] * Jb[tr ? 6 : 3] + m[1] * nn[0] * Jb[tr ? 3 : 6];
            } else {
                corr += m[0] * nn[0] * Jb[0] + m[1] * nn[1] * Jb[1] + m[2] * nn[2] * Jb[2];
                if (g.fullJ)
                    corr += m[0] * nn[1] * Jb[tr ? 6 : 3] + m[0] * nn[2] * Jb[tr ? 7 : 4] + m[1] * nn[2] * Jb[tr ? 8 : 5] +
                            m[1] * nn[0] * Jb[tr ? 3 : 6] + m[2] * nn[0] * Jb[tr ? 4 : 7] + m[2] * nn[1] * Jb[tr ? 5 : 8];
            }
        }
        // on-site part with the UN-renormalised spin of the chosen site (heisenbergLib.c:799, isingLib.c:390)
        int p = struct_pos(a, x, y, z, o);
        double s[3] = {(double)sp[p], NC >= 2 ? (double)sp[(size_t)a.N + p] : 0.0, NC == 3 ? (double)sp[2 * (size_t)a.N + p] : 0.0};
        double eo;
        if (NC == 1) eo = -hf * s[0];
        else {
            const double *D = g.SD + 4 * o + 1;
            eo = beta * (D[0] * s[0] * s[0] + D[1] * s[1] * s[1] + (NC == 3 ? D[2] * s[2] * s[2] : 0.0)) - hf * (NC == 3 ? s[2] : s[0]);
        }
        v[RS_E] += 0.5 * beta * corr + eo;
        // coarse pair statistics (heisenbergLib.c:748-790): a correlated pair contributes its i member (cell c, orbital ps) when c
        // is a coarse cell, its j member (cell c + kl, orbital pt) when c + kl is one - as c runs over the lattice both are
        // exactly the coarse cells - and the product when both are, i.e. for every coarse cell if kl is even
        if (o == g.ps) {
            v[RS_I] += m[0]; v[RS_I + 1] += m[1]; v[RS_I + 2] += m[2];
            if (evenShift) {
                int rj = rg_row(a, g, wrap_add(x, g.pd0, a.Lx), wrap_add(y, g.pd1, a.Ly), wrap_add(z, g.pd2, a.Lz), g.pt);
                v[RS_IJ] += m[0] * ms[3 * rj] + m[1] * ms[3 * rj + 1] + m[2] * ms[3 * rj + 2];
            }
        }
        if (o == g.pt) { v[RS_J] += m[0]; v[RS_J + 1] += m[1]; v[RS_J + 2] += m[2]; }
    }
    block_accumulate<NRS>(v, g.rsums + (size_t)r * NRS, smem);
}

template <int NC, typename real>
__global__ void __launch_bounds__(256) k_struct_init(StructArgs a, double flunc, int LxGlobal) {
    int r = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    int q = p / a.ncellc;
    double S = a.classes[q].S;   // signed S of the orbital
    double aS = fabs(S);
    if (NC == 1) { sp[p] = (real)S; return; }
    if (flunc == 0.0) {   // the reference's default start (win.py never passes flunc): polarised along x, no RNG needed
        sp[p] = (real)S;
        sp[(size_t)a.N + p] = real(0);
        if (NC == 3) sp[2 * (size_t)a.N + p] = real(0);
        return;
    }
    uint32_t w[4];
    rng4(a.key, a.replica0 + r, STREAM_INIT, 0, 0, (uint32_t)struct_site_gid(a, p, LxGlobal), w);
    double n[3];
    if (sizeof(real) == 4) { float nf[3]; random_dir<NC, float>(w[0], w[1], nf); n[0] = nf[0]; n[1] = nf[1]; n[2] = nf[2]; }
    else random_dir<NC, double>(w[0], w[1], n);
    double v[3] = {S + flunc * n[0], flunc * n[1], NC == 3 ? flunc * n[2] : 0.0};
    double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(len < 1e-5)) { v[0] /= len; v[1] /= len; v[2] /= len; }
    sp[p] = (real)(v[0] * aS);
    sp[(size_t)a.N + p] = (real)(v[1] * aS);
    if (NC == 3) sp[2 * (size_t)a.N + p] = (real)(v[2] * aS);
}

template <int NC, typename real, bool GATHER>
__global__ void __launch_bounds__(256) k_struct_frame(StructArgs a, int r, double *buf) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    real *sp = (real *)a.spin + (size_t)r * NC * a.N;
    size_t i = (size_t)struct_site_id(a, p);
    if (NC == 1) {
        if (GATHER) buf[i] = sp[p]; else sp[p] = (real)buf[i];
        return;
    }
    if (GATHER) {
        buf[3 * i] = sp[p]; buf[3 * i + 1] = sp[(size_t)a.N + p];
        buf[3 * i + 2] = NC == 3 ? (double)sp[2 * (size_t)a.N + p] : 0.0;
    } else {
        sp[p] = (real)buf[3 * i]; sp[(size_t)a.N + p] = (real)buf[3 * i + 1];
        if (NC == 3) sp[2 * (size_t)a.N + p] = (real)buf[3 * i + 2];
    }
}


// ---- int8 Ising planes (precision 8, ising8.cuh): initial state, frame gather / scatter, pair correlation ----
static __global__ void __launch_bounds__(256) k_i8_init(StructArgs a) {
    const int r = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    ((unsigned char *)a.spin)[(size_t)r * a.N + p] = a.classes[p / a.ncellc].S < 0 ? 1 : 0;   // byte = "spin is down"; initSpin carries the signed S (isingLib.c:23-40)
}
template <bool GATHER> static __global__ void __launch_bounds__(256) k_i8_frame(StructArgs a, int r, double *buf) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.N) return;
    unsigned char *sp = (unsigned char *)a.spin + (size_t)r * a.N;
    const size_t i = (size_t)struct_site_id(a, p);
    if (GATHER) buf[i] = (sp[p] ? -1.0 : 1.0) * fabs(a.classes[p / a.ncellc].S);
    else sp[p] = buf[i] < 0 ? 1 : 0;
}
static __global__ void __launch_bounds__(256) k_i8_pairs(StructArgs a, int ps, int pt, int d0, int d1, int d2, double *sums) {
    __shared__ double smem[32];
    const int r = blockIdx.y, cell = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (cell < a.Lx * a.Ly * a.Lz) {
        const int z = cell % a.Lz, y = (cell / a.Lz) % a.Ly, x = cell / (a.Lz * a.Ly);
        const int pi = struct_pos(a, x, y, z, ps);
        const int pj = struct_pos(a, (x + d0 + a.Lx) % a.Lx, (y + d1 + a.Ly) % a.Ly, (z + d2 + a.Lz) % a.Lz, pt);
        const unsigned char *sp = (const unsigned char *)a.spin + (size_t)r * a.N;
        v[0] = (sp[pi] != sp[pj] ? -1.0 : 1.0) * fabs(a.classes[pi / a.ncellc].S) * fabs(a.classes[pj / a.ncellc].S);
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_SIJ, smem);
}

// ---------------------------------------------------------------------------------------------
// JIT: generate the prologue (lattice as literals), compile struct_pass.cuh with NVRTC for sm_100a,
// load the cubin through the driver API.  libnvrtc / libcuda are dlopen'ed lazily so that the
// library itself has no link-time dependency on them; any failure falls back to the offline
// (runtime-table) CUDA kernel - never to a CPU path.
// ---------------------------------------------------------------------------------------------
}  // namespace mcg
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <sstream>
namespace mcg {

struct JitApi {
    bool ok = false, rtc_ok = false;
    std::string why;
    decltype(&nvrtcCreateProgram) createProgram;
    decltype(&nvrtcCompileProgram) compileProgram;
    decltype(&nvrtcGetCUBINSize) getCUBINSize;
    decltype(&nvrtcGetCUBIN) getCUBIN;
    decltype(&nvrtcGetProgramLogSize) getLogSize;
    decltype(&nvrtcGetProgramLog) getLog;
    decltype(&nvrtcDestroyProgram) destroyProgram;
    decltype(&cuModuleLoadData) moduleLoadData;
    decltype(&cuModuleGetFunction) moduleGetFunction;
    decltype(&cuLaunchKernel) launchKernel;
};

static JitApi &jit_api() {
    static JitApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *rtc = nullptr;
        for (const char *n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
            if ((rtc = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        void *drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!rtc) { api.why = "cannot dlopen libnvrtc"; return; }
#define MCG_SYM(lib, field, name)                                               \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name));        \
    if (!api.field) { api.why = std::string("missing symbol ") + name; return; }
        MCG_SYM(rtc, createProgram, "nvrtcCreateProgram")
        MCG_SYM(rtc, compileProgram, "nvrtcCompileProgram")
        MCG_SYM(rtc, getCUBINSize, "nvrtcGetCUBINSize")
        MCG_SYM(rtc, getCUBIN, "nvrtcGetCUBIN")
        MCG_SYM(rtc, getLogSize, "nvrtcGetProgramLogSize")
        MCG_SYM(rtc, getLog, "nvrtcGetProgramLog")
        MCG_SYM(rtc, destroyProgram, "nvrtcDestroyProgram")
        api.rtc_ok = true;
        if (!drv) { api.why = "cannot dlopen libcuda.so.1"; return; }
        MCG_SYM(drv, moduleLoadData, "cuModuleLoadData")
        MCG_SYM(drv, moduleGetFunction, "cuModuleGetFunction")
        MCG_SYM(drv, launchKernel, "cuLaunchKernel")
#undef MCG_SYM
        api.ok = true;
    });
    return api;
}

static std::string csrc_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&csrc_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "mcsolver_b200/csrc";
}

static std::string lit(double v, bool f32) {
    char buf[64];
    if (f32) snprintf(buf, sizeof buf, "%.9ef", (double)(float)v);
    else snprintf(buf, sizeof buf, "%.17e", v);
    return buf;
}

std::string jit_prologue(const mcg_system *s, int colour, bool partial) {
    const StructuredSystem *st = s->st;
    const bool f32 = s->prec == 32;
    const int q0 = st->colourClassStart[colour], nqc = st->colourClassStart[colour + 1] - q0;
    const int JW = s->NC == 1 ? 1 : 9;
    // resident blocks per SM the kernel is compiled for (register budget 65536 / (256 * minb)): four for the short link lists
    // (sc: 64 registers, measured best of 2..5), two once a class has more than 8 links - the unrolled field sums then need the
    // registers more than the occupancy (CrI3, 12 links: +7 % over four)
    int maxLinks = 0;
    for (int j = 0; j < nqc; j++) maxLinks = std::max(maxLinks, st->classes[q0 + j].nlink);
    const int minb = getenv("MCG_JIT_MINB") ? atoi(getenv("MCG_JIT_MINB")) : (f32 && maxLinks <= 8 ? 4 : 2);
    std::ostringstream o;
    o << "#define MCG_JIT 1\ntypedef " << (f32 ? "float" : "double") << " jit_real;\n";
    if (getenv("MCG_NO_F32X2")) o << "#define MCG_NO_F32X2 1\n";   // A/B switch: scalar fp32 arithmetic instead of packed pairs
    // bulk L2 prefetch of the next row's operands (struct_pass.cuh)
    // 3 (own row + first link's row) pays where the pass is latency-bound with nothing saturated: fp64 state on short rows
    // (sc 256^3: 0.66 -> 0.75 of the HBM roofline); it costs 2-18 % elsewhere (fp32, long 2D rows), and asking for every
    // neighbour row (1) doubles the L2 request traffic and loses 7-20 % everywhere: profiles/r02b_prefetch_ab.txt
    const int pf = getenv("MCG_JIT_PF") ? atoi(getenv("MCG_JIT_PF")) : (!f32 && st->Zd / st->V <= 64 && st->Xd > 1 ? 3 : 0);
    o << "#define JIT_PF " << pf << "\n";
    if (getenv("MCG_JIT_ACC")) o << "#define JIT_ACC " << atoi(getenv("MCG_JIT_ACC")) << "\n";   // A/B: see struct_pass.cuh
    o << "#define JIT_NC " << s->NC << "\n#define JIT_FULLJ " << (s->fullJ ? "true" : "false") << "\n#define JIT_V " << st->V
      << "\n#define JIT_PARTIAL " << (partial ? "true" : "false") << "\n#define JIT_NQC " << nqc << "\n#define JIT_MINB " << minb << "\n";
    o << "#define JIT_Xd " << st->Xd << "\n#define JIT_Yd " << st->Yd << "\n#define JIT_Zd " << st->Zd << "\n#define JIT_Zc "
      << st->Zd / st->V << "\n#define JIT_N " << s->N << "\n#define JIT_px " << st->p[0] << "\n#define JIT_py " << st->p[1]
      << "\n#define JIT_pz " << st->p[2] << "\n#define JIT_norb " << st->norb << "\n#define JIT_Ly " << st->L[1]
      << "\n#define JIT_Lz " << st->L[2] << "\n#define JIT_nrows " << st->nrows << "\n#define JIT_nclass " << st->nclass << "\n";
    o << "#define JIT_xoff " << st->xoff << "\n";      // the row range [rowLo, rowHi) is read from the kernel arguments
    o << "namespace mcg {\ntemplate <int J, int K> struct CtLinkData;\ntemplate <int J> struct CtClassData;\n";
    for (int j = 0; j < nqc; j++) {
        const SClassD &cl = st->classes[q0 + j];
        o << "template <> struct CtClassData<" << j << "> { static constexpr int nl=" << cl.nlink << ", ca=" << cl.a << ", cb=" << cl.b
          << ", cc=" << cl.c << ", co=" << cl.o << ", lowmode=" << cl.lowmode << ", nlow=" << cl.pad << "; static constexpr jit_real S=" << lit(std::fabs(cl.S), f32)
          << "; static __device__ constexpr jit_real D(int e) { constexpr jit_real v[3] = {" << lit(cl.D[0], f32) << "," << lit(cl.D[1], f32)
          << "," << lit(cl.D[2], f32) << "}; return v[e]; } };\n";
        for (int k = 0; k < cl.nlink; k++) {
            size_t li = (size_t)(q0 + j) * MAXLINK + k;
            int cx = st->cXs[li], cy = st->cYs[li];
            int delta = (st->linksHost[li].qn - (q0 + j)) * st->ncellc + (cx * st->Yd + cy) * st->Zd;
            o << "template <> struct CtLinkData<" << j << "," << k << "> { static constexpr int delta=" << delta << ", mxp=" << (cx > 0)
              << ", mxm=" << (cx < 0) << ", myp=" << (cy > 0) << ", mym=" << (cy < 0) << ", cZ=" << st->linksHost[li].cZ
              << ", low=" << st->linksHost[li].low << "; static __device__ constexpr jit_real J(int e) { constexpr jit_real v[9] = {";
            for (int e = 0; e < 9; e++) o << (e ? "," : "") << lit(e < JW ? st->JHost[li * JW + e] : 0.0, f32);
            o << "}; return v[e]; } };\n";
        }
    }
    o << "}\n#include \"struct_pass.cuh\"\n";
    return o.str();
}

// the int8 Ising pass (ising8.cuh) with the lattice as literals: same entry-point names, so the loader below serves both
std::string jit_i8_prologue(const mcg_system *s, int colour, bool partial) {
    const StructuredSystem *st = s->st;
    const I8Table &T = st->i8Tables[colour];
    const int minb = getenv("MCG_JIT_MINB") ? atoi(getenv("MCG_JIT_MINB")) : 4;
    std::ostringstream o;
    o << "#define MCG_JIT_I8 1\n#define JIT_NW " << st->V / 4 << "\n#define JIT_PARTIAL " << (partial ? "true" : "false") << "\n#define JIT_NQC " << T.nqc
      << "\n#define JIT_MINB " << minb << "\n#define JIT_JS2 " << lit(T.JS2, false) << "\n#define JIT_S " << lit(T.S, false) << "\n";
    o << "#define JIT_Xd " << st->Xd << "\n#define JIT_Yd " << st->Yd << "\n#define JIT_Zd " << st->Zd << "\n#define JIT_Zc "
      << st->Zd / st->V << "\n#define JIT_N " << s->N << "\n#define JIT_px " << st->p[0] << "\n#define JIT_py " << st->p[1]
      << "\n#define JIT_pz " << st->p[2] << "\n#define JIT_norb " << st->norb << "\n#define JIT_Ly " << st->L[1]
      << "\n#define JIT_Lz " << st->L[2] << "\n#define JIT_nrows " << st->nrows << "\n#define JIT_nclass " << st->nclass << "\n";
    o << "namespace mcg {\ntemplate <int J> struct I8Ct;\n";
    for (int j = 0; j < T.nqc; j++) {
        const I8Class &c = T.c[j];
        o << "template <> struct I8Ct<" << j << "> { static constexpr int nl=" << c.nl << ", nlow=" << c.nlow << ", lowmode=" << c.lowmode << ", ca=" << c.ca
          << ", cb=" << c.cb << ", cc=" << c.cc << ", co=" << c.co << ";\n";
        auto arr = [&](const char *name, auto get) {
            o << "  static __device__ constexpr int " << name << "(int k) { constexpr int v[" << c.nl << "] = {";
            for (int k = 0; k < c.nl; k++) o << (k ? "," : "") << get(k);
            o << "}; return v[k]; }\n";
        };
        arr("delta", [&](int k) { return c.delta[k]; });
        arr("wx", [&](int k) { return (int)c.wx[k]; });
        arr("wy", [&](int k) { return (int)c.wy[k]; });
        arr("cz", [&](int k) { return (int)c.cz[k]; });
        o << "};\n";
    }
    o << "}\n#include \"ising8.cuh\"\n";
    return o.str();
}

struct JitPass {
    CUfunction f[2] = {nullptr, nullptr};
    bool failed = false;
};
static std::mutex g_jit_mutex;
static std::map<std::pair<int, std::string>, JitPass> g_jit_cache;   // (device, prologue) -> loaded kernels

// ---- on-disk cubin cache: $MCG_CACHE_DIR or <library dir>/build/jitcache/<fnv1a(prologue + kernel headers)>.cubin ----
static uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
static std::string read_file(const std::string &path) {
    std::string out;
    if (FILE *f = fopen(path.c_str(), "rb")) {
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
        fclose(f);
    }
    return out;
}
// identity of a specialised module: generated prologue + the kernel headers it includes
static uint64_t cache_key(const std::string &src) {
    uint64_t h = fnv1a(src);
    for (const char *f : {"/struct_pass.cuh", "/ising8.cuh", "/topo_pass.cuh", "/devmath.cuh", "/rng.cuh"}) h = fnv1a(read_file(csrc_dir() + f), h);
    return h;
}
uint64_t structured_jit_key(const mcg_system *s, int colour) {
    if (!s->structured || colour < 0 || colour >= s->C) throw Error(MCG_ERR_ARG, "not a structured system / colour out of range");
    return cache_key(s->prec == 8 ? jit_i8_prologue(s, colour, false) : jit_prologue(s, colour, false));
}
static std::string cache_path(const std::string &src) {
    const char *dir = getenv("MCG_CACHE_DIR");
    std::string d;
    if (dir && dir[0]) d = dir;
    else d = csrc_dir() + "/../build/jitcache";
    if (getenv("MCG_NO_DISK_CACHE")) return "";
    uint64_t h = cache_key(src);
    std::string mk = "mkdir -p '" + d + "' 2>/dev/null";
    if (system(mk.c_str()) != 0) return "";
    char name[64];
    snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)h);
    return d + name;
}

// compile (or fetch) the specialised kernels; also usable without a GPU up to the cubin (tests)
std::vector<char> jit_compile_cubin(const std::string &src, std::string &log) {
    JitApi &api = jit_api();
    if (!api.rtc_ok) { log = api.why; return {}; }
    const std::string cpath = cache_path(src);
    if (!cpath.empty()) {
        std::string c = read_file(cpath);
        if (c.size() > 1024) return std::vector<char>(c.begin(), c.end());
    }
    nvrtcProgram prog;
    if (api.createProgram(&prog, src.c_str(), "mcg_pass.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return {}; }
    std::string inc = "-I" + csrc_dir();
    // 128: "loop is not reachable" - the scalar item code after the packed fp32 branch of pass_body, by construction
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str(), "-default-device", "--diag-suppress=128"};
    nvrtcResult res = api.compileProgram(prog, 6, opts);
    size_t ls = 0;
    api.getLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); api.getLog(prog, &log[0]); }
    std::vector<char> cubin;
    if (res == NVRTC_SUCCESS) {
        size_t cs = 0;
        api.getCUBINSize(prog, &cs);
        cubin.resize(cs);
        api.getCUBIN(prog, cubin.data());
        if (!cpath.empty()) {   // write-then-rename so concurrent ranks never read a partial file
            std::string tmp = cpath + "." + std::to_string((long long)getpid());
            if (FILE *f = fopen(tmp.c_str(), "wb")) {
                bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
                fclose(f);
                if (!ok || rename(tmp.c_str(), cpath.c_str()) != 0) remove(tmp.c_str());
            }
        }
        if (const char *dump = getenv("MCG_JIT_DUMP")) {   // debugging aid: keep the last cubin for cuobjdump -sass
            if (FILE *f = fopen(dump, "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
        }
    }
    api.destroyProgram(&prog);
    return cubin;
}

static bool jit_enabled(const mcg_system *s) {
    const char *e = getenv("MCG_JIT");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return (long long)s->N * s->R >= (1ll << 20);   // creation-time compile (~seconds) only pays for big jobs
}

// Full unrolling only pays while the specialised body stays inside the instruction cache: measured on the
// 32-link dipole stencil (4 classes x 32 links x 4 sites unrolled) the JIT kernel ran 3x SLOWER than the
// runtime-table kernel and took minutes to compile.  Budget: <= 16 links per class, <= 64 link bodies per module.
static bool jit_worthwhile(const mcg_system *s, int colour) {
    const StructuredSystem *st = s->st;
    int q0 = st->colourClassStart[colour], nqc = st->colourClassStart[colour + 1] - q0, total = 0;
    for (int j = 0; j < nqc; j++) {
        if (st->classes[q0 + j].nlink > 16) return false;
        total += st->classes[q0 + j].nlink;
    }
    return total <= 64;
}

bool jit_launch_pass(mcg_system *s, int colour, int mode, const StructArgs &a, int q0, int rowsPerBlock, int nrb, uint64_t sweep,
                     double pAtt, dim3 grid, dim3 block) {
    if (!jit_enabled(s) || !jit_worthwhile(s, colour)) return false;
    JitApi &api = jit_api();
    if (!api.ok) return false;
    StructuredSystem *st = s->st;
    const int slotIdx = colour * 2 + (pAtt < 1.0 ? 1 : 0);
    if (st->jitResolved.empty()) st->jitResolved.assign((size_t)s->C * 2, nullptr);
    JitPass *jp = static_cast<JitPass *>(st->jitResolved[slotIdx]);
    if (!jp) {
        std::lock_guard<std::mutex> lock(g_jit_mutex);
        auto key = std::make_pair(s->device, s->prec == 8 ? jit_i8_prologue(s, colour, pAtt < 1.0) : jit_prologue(s, colour, pAtt < 1.0));
        auto it = g_jit_cache.find(key);
        if (it == g_jit_cache.end()) {
            JitPass np;
            std::string log;
            std::vector<char> cubin = jit_compile_cubin(key.second, log);
            CUmodule mod = nullptr;
            if (cubin.empty() || api.moduleLoadData(&mod, cubin.data()) != CUDA_SUCCESS ||
                api.moduleGetFunction(&np.f[0], mod, "mcg_pass_m0") != CUDA_SUCCESS ||
                api.moduleGetFunction(&np.f[1], mod, "mcg_pass_m1") != CUDA_SUCCESS) {
                np.failed = true;
                fprintf(stderr, "mcsolver_b200: JIT specialisation unavailable (%s); using the offline CUDA kernel\n", log.substr(0, 2000).c_str());
            }
            it = g_jit_cache.emplace(key, np).first;
        }
        jp = &it->second;               // std::map nodes are stable: the pointer stays valid
        st->jitResolved[slotIdx] = jp;
    }
    if (jp->failed) return false;
    float pf = (float)pAtt;
    double pd = pAtt;
    void *params[] = {(void *)&a, &q0, &rowsPerBlock, &nrb, &sweep, s->prec == 32 ? (void *)&pf : (void *)&pd};
    CUresult r = api.launchKernel(jp->f[mode], grid.x, 1, 1, block.x, block.y, 1, 0, (CUstream)s->stream, params, nullptr);
    if (r != CUDA_SUCCESS) throw Error(MCG_ERR_CUDA, "cuLaunchKernel of the JIT pass kernel failed");
    s->jitLaunches++;
    return true;
}

// ---- specialised topological-charge kernel (topo_pass.cuh) ----
static bool jit_topo_worthwhile(const StructuredSystem *st) {
    const int npar = st->p[0] * st->p[1] * st->p[2];
    return st->ncircuit > 0 && st->ncircuit <= 32 && st->nvert <= 16 && npar * st->nvert <= 256;
}

static int topo_ypt(const StructuredSystem *st) { return st->Xd >= TOPO_XPT ? 1 : TOPO_YPT; }

std::string jit_topo_prologue(const mcg_system *s) {
    const StructuredSystem *st = s->st;
    const bool f32 = s->prec == 32;
    const int px = st->p[0], py = st->p[1], pz = st->p[2], no = st->norb, npar = px * py * pz;
    std::ostringstream o;
    o << "#define MCG_JIT_TOPO 1\ntypedef " << (f32 ? "float" : "double") << " jit_real;\n";
    o << "#define JT_NV " << st->nvert << "\n#define JT_NT " << st->ncircuit << "\n#define JT_NPAR " << npar << "\n#define JT_Xd " << st->Xd
      << "\n#define JT_Yd " << st->Yd << "\n#define JT_Zd " << st->Zd << "\n#define JT_N " << s->N << "\n#define JT_YPT " << topo_ypt(st) << "\n";
    o << "namespace mcg {\ntemplate <int PAR, int K> struct CtVert;\ntemplate <int T> struct CtTri;\n";
    for (int par = 0; par < npar; par++) {
        const int cc = par % pz, cb = (par / pz) % py, ca = par / (pz * py);
        for (int k = 0; k < st->nvert; k++) {
            const int *e = st->tvertsHost.data() + 4 * k;   // (orbital, dx, dy, dz), offsets in [0, L)
            const int na = ca + e[1], nb = cb + e[2], nc = cc + e[3];
            const int q = st->classOf[(((na % px) * py + nb % py) * pz + nc % pz) * no + e[0]];
            o << "template <> struct CtVert<" << par << "," << k << "> { static constexpr int base=" << q * st->ncellc << ", cX=" << (na / px) % st->Xd
              << ", cY=" << (nb / py) % st->Yd << ", cZ=" << (nc / pz) % st->Zd << "; static constexpr jit_real len=" << lit(std::fabs(st->classes[q].S), f32)
              << "; };\n";
        }
    }
    for (int t = 0; t < st->ncircuit; t++)
        o << "template <> struct CtTri<" << t << "> { static constexpr int i0=" << st->ttrisHost[3 * t] << ", i1=" << st->ttrisHost[3 * t + 1]
          << ", i2=" << st->ttrisHost[3 * t + 2] << "; };\n";
    o << "}\n#include \"topo_pass.cuh\"\n";
    return o.str();
}

static bool jit_launch_topo(mcg_system *s, const StructArgs &a, int nzc, int nyc, dim3 grid, dim3 block) {
    StructuredSystem *st = s->st;
    if (!jit_enabled(s) || !jit_topo_worthwhile(st)) return false;
    JitApi &api = jit_api();
    if (!api.ok) return false;
    JitPass *jp = static_cast<JitPass *>(st->jitTopo);
    if (!jp) {
        std::lock_guard<std::mutex> lock(g_jit_mutex);
        auto key = std::make_pair(s->device, jit_topo_prologue(s));
        auto it = g_jit_cache.find(key);
        if (it == g_jit_cache.end()) {
            JitPass np;
            std::string log;
            std::vector<char> cubin = jit_compile_cubin(key.second, log);
            CUmodule mod = nullptr;
            if (cubin.empty() || api.moduleLoadData(&mod, cubin.data()) != CUDA_SUCCESS ||
                api.moduleGetFunction(&np.f[0], mod, "mcg_topo") != CUDA_SUCCESS) {
                np.failed = true;
                fprintf(stderr, "mcsolver_b200: JIT topological-charge kernel unavailable (%s); using the offline CUDA kernel\n", log.substr(0, 2000).c_str());
            }
            it = g_jit_cache.emplace(key, np).first;
        }
        jp = &it->second;
        st->jitTopo = jp;
    }
    if (jp->failed) return false;
    double *sums = s->d_sums;
    const int ypt = topo_ypt(st);
    const int npar = st->p[0] * st->p[1] * st->p[2];
    nyc = (st->Yd + (int)block.y * ypt - 1) / ((int)block.y * ypt);
    grid.x = (unsigned)(nzc * nyc * npar);
    void *params[] = {(void *)&a, &nzc, &nyc, &sums};
    CUresult r = api.launchKernel(jp->f[0], grid.x, (grid.y + TOPO_XPT - 1) / TOPO_XPT, grid.z, block.x, block.y, 1, 0, (CUstream)s->stream, params, nullptr);
    if (r != CUDA_SUCCESS) throw Error(MCG_ERR_CUDA, "cuLaunchKernel of the JIT topological-charge kernel failed");
    s->jitLaunches++;
    return true;
}

// ---------------------------------------------------------------------------------------------
// host: link templates, colouring period search, class tables
// ---------------------------------------------------------------------------------------------
struct Tmpl {
    int o2;
    int d[3];       // internal axes, reduced mod L into (-L/2, L/2]
    double J[9];
    bool self;
    bool fwd;       // created from the bond's source side
};

static int mod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

static StructArgs struct_args(const mcg_system *s) {
    const StructuredSystem *st = s->st;
    StructArgs a;
    a.Xd = st->Xd; a.Yd = st->Yd; a.Zd = st->Zd; a.Zc = st->Zd / st->V; a.ncellc = st->ncellc; a.nclass = st->nclass;
    a.nrows = st->nrows; a.N = s->N;
    a.px = st->p[0]; a.py = st->p[1]; a.pz = st->p[2]; a.norb = st->norb;
    auto lg = [](int p) { int s = 0; while ((1 << s) < p) s++; return (1 << s) == p ? s : -1; };
    a.psx = lg(a.px); a.psy = lg(a.py); a.psz = lg(a.pz); a.Lx = st->L[0]; a.Ly = st->L[1]; a.Lz = st->L[2];
    a.classes = st->d_classes; a.links = st->d_links; a.J = st->d_J; a.classOf = st->d_classOf;
    a.spin = s->d_spin; a.beta = s->d_beta; a.field = s->d_field; a.cnt = s->d_cnt; a.classSums = st->d_classSums;
    a.key = make_rng_key(s->seed);
    a.replica0 = s->replica0;
    a.rowLo = st->rowLo; a.rowHi = st->rowHi; a.xoff = st->xoff;
    return a;
}

// plan of a slab decomposition along X: the descriptor handed to structured_build_host is the LOCAL lattice (own planes plus one
// ghost coarse cell on either side), coloured with the period of the whole lattice
struct SlabPlan { int rank, world, period[3], LxGlobal; };

static void structured_build_host(mcg_system *s, const mcg_lattice_desc *d, std::vector<SLinkD> &links, std::vector<double> &Jt,
                                  const SlabPlan *plan = nullptr) {
    MCG_REQUIRE(d->model >= 1 && d->model <= 3, "model must be 1, 2 or 3");
    MCG_REQUIRE(d->norb >= 1 && d->S, "norb/S invalid");
    for (int k = 0; k < 3; k++) MCG_REQUIRE(d->L[k] >= 1, "supercell dims must be >= 1");
    MCG_REQUIRE((long long)d->L[0] * d->L[1] * d->L[2] * d->norb < (1ll << 31), "lattice too large for 32-bit site ids");
    MCG_REQUIRE(d->nbond >= 0 && (d->nbond == 0 || d->bonds), "bonds invalid");
    MCG_REQUIRE(d->pair_s >= 0 && d->pair_s < d->norb && d->pair_t >= 0 && d->pair_t < d->norb, "pair orbital out of range");
    std::unique_ptr<StructuredSystem> stp(new StructuredSystem());
    StructuredSystem *st = stp.get();
    const int no = d->norb;
    st->norb = no;
    // canonical internal axes: drop unit dims, left-pad with 1 (keeps the row-major site id formula)
    int ax[3], nax = 0;
    for (int k = 0; k < 3; k++) if (d->L[k] > 1) ax[nax++] = k;
    int map[3] = {-1, -1, -1};   // internal axis i <- original axis map[i] (or -1)
    for (int i = 0; i < nax; i++) map[3 - nax + i] = ax[i];
    for (int i = 0; i < 3; i++) st->L[i] = map[i] >= 0 ? d->L[map[i]] : 1;
    auto conv = [&](const int32_t *v, int *o) { for (int i = 0; i < 3; i++) o[i] = map[i] >= 0 ? v[map[i]] : 0; };
    const int *L = st->L;

    s->model = d->model; s->NC = d->model; s->structured = true;
    s->N = L[0] * L[1] * L[2] * no;
    bool full = false;
    // link templates per orbital, with the reference's merge rule (Lattice.py:36-52) applied per template
    std::vector<std::vector<Tmpl>> tm(no);
    auto add_link = [&](int o, int o2, const int *dd, const double *J9, bool transpose) {
        Tmpl t;
        t.o2 = o2;
        for (int k = 0; k < 3; k++) { int m = mod(dd[k], L[k]); if (m > L[k] / 2) m -= L[k]; t.d[k] = m; }
        static const int T9[9] = {0, 1, 2, 6, 7, 8, 3, 4, 5};
        for (int k = 0; k < 9; k++) t.J[k] = d->model == 1 ? (k == 0 ? J9[0] : 0.0) : J9[transpose ? T9[k] : k];
        t.self = (o2 == o && t.d[0] == 0 && t.d[1] == 0 && t.d[2] == 0);
        t.fwd = !transpose;
        if (t.self && transpose) return;   // Lattice.py:261: no back link to oneself
        for (auto &e : tm[o])
            if (e.o2 == t.o2 && e.d[0] == t.d[0] && e.d[1] == t.d[1] && e.d[2] == t.d[2]) {
                double diff = 0;
                for (int k = 0; k < 9; k++) diff += std::fabs(e.J[k] - t.J[k]);
                if (diff < 1e-5) return;
                for (int k = 0; k < 9; k++) e.J[k] += t.J[k];
                return;
            }
        tm[o].push_back(t);
    };
    for (int b = 0; b < d->nbond; b++) {
        const mcg_bond &B = d->bonds[b];
        MCG_REQUIRE(B.src >= 0 && B.src < no && B.tgt >= 0 && B.tgt < no, "bond orbital index out of range");
        int dd[3], nd[3];
        conv(B.d, dd);
        for (int k = 0; k < 3; k++) nd[k] = -dd[k];
        add_link(B.src, B.tgt, dd, B.J, false);
        add_link(B.tgt, B.src, nd, B.J, true);
    }
    for (int o = 0; o < no; o++) {
        MCG_REQUIRE((int)tm[o].size() <= MAXLINK, "more than 32 links per site: use the table path");
        for (auto &t : tm[o]) {
            if (t.self) st->hasSelf = true;
            if (d->model != 1) for (int k = 3; k < 9; k++) if (std::fabs(t.J[k]) > 1e-6) full = true;   // mcMain.py:174
        }
    }
    s->fullJ = full;
    s->isoNoOnsite = true;
    for (int o = 0; o < no; o++) {
        for (int k = 0; k < 3 && d->D && d->model != 1; k++) if (d->D[3 * o + k] != 0.0) s->isoNoOnsite = false;
        for (auto &t : tm[o]) {
            bool iso = d->model == 1 || (t.J[0] == t.J[1] && (d->model == 2 || t.J[1] == t.J[2]));
            for (int k = 3; k < 9 && d->model != 1; k++) if (t.J[k] != 0.0 && (d->model == 3 || k == 3 || k == 6)) iso = false;
            if (!iso) s->isoNoOnsite = false;
        }
    }

    // ---- colouring period search ----
    auto candidates = [&](int Ld, int ax) {
        std::vector<int> c;
        if (plan) { c.push_back(plan->period[ax]); return c; }   // a slab is coloured like the whole lattice
        if (Ld == 1) { c.push_back(1); return c; }
        for (int p = 1; p <= 6; p++) if (Ld % p == 0) c.push_back(p);
        if (Ld > 6 && Ld <= 16) c.push_back(Ld);
        return c;
    };
    int best[3] = {0, 0, 0}, bestC = 1 << 30, bestN = 1 << 30;
    std::vector<int> bestColour;
    const std::vector<int> candX = candidates(L[0], 0), candY = candidates(L[1], 1), candZ = candidates(L[2], 2);
    for (int px : candX) for (int py : candY) for (int pz : candZ) {
        int ncls = px * py * pz * no;
        if (ncls > 1024) continue;
        auto cid = [&](int a, int b, int c, int o) { return ((a * py + b) * pz + c) * no + o; };
        std::vector<std::vector<int>> adj(ncls);
        bool ok = true;
        for (int a = 0; a < px && ok; a++) for (int b = 0; b < py && ok; b++) for (int c = 0; c < pz && ok; c++)
            for (int o = 0; o < no && ok; o++)
                for (auto &t : tm[o]) {
                    if (t.self) continue;
                    int n = cid(mod(a + t.d[0], px), mod(b + t.d[1], py), mod(c + t.d[2], pz), t.o2);
                    if (n == cid(a, b, c, o)) { ok = false; break; }
                    adj[cid(a, b, c, o)].push_back(n);
                    adj[n].push_back(cid(a, b, c, o));
                }
        if (!ok) continue;
        std::vector<int> col(ncls, -1);
        int C = 0;
        for (int i = 0; i < ncls; i++) {
            std::vector<char> used(ncls + 1, 0);
            for (int n : adj[i]) if (col[n] >= 0) used[col[n]] = 1;
            int c = 0;
            while (used[c]) c++;
            col[i] = c;
            C = std::max(C, c + 1);
        }
        if (C < bestC || (C == bestC && ncls < bestN)) {
            bestC = C; bestN = ncls; best[0] = px; best[1] = py; best[2] = pz; bestColour = col;
        }
    }
    MCG_REQUIRE(best[0] > 0, "no periodic colouring with period <= 6 divides this supercell: use the table path");
    for (int k = 0; k < 3; k++) st->p[k] = best[k];
    const int px = best[0], py = best[1], pz = best[2];
    st->nclass = bestN; s->C = bestC;
    st->Xd = L[0] / px; st->Yd = L[1] / py; st->Zd = L[2] / pz;
    st->ncellc = st->Xd * st->Yd * st->Zd;
    st->nrows = st->Xd * st->Yd;
    st->rowLo = 0; st->rowHi = st->nrows; st->xoff = 0; st->LxGlobal = L[0];
    if (plan) {
        MCG_REQUIRE(st->Xd >= 4, "slab too thin: a rank needs at least two coarse planes of its own");
        st->rowLo = st->Yd; st->rowHi = (st->Xd - 1) * st->Yd;
        st->LxGlobal = plan->LxGlobal;
        st->xoff = plan->rank * (plan->LxGlobal / plan->world) - px;
        st->slabRank = plan->rank; st->slabWorld = plan->world;
    }
    int Vmax = s->prec == 32 ? 4 : 2;
    st->V = (st->Zd % Vmax == 0) ? Vmax : 1;
    if (s->prec == 8) {   // int8 Ising items: 16 sites (one 16-byte load) or 4 (one word)
        MCG_REQUIRE(d->model == 1, "precision 8 (int8 spins) is for the Ising model");
        MCG_REQUIRE(st->Zd % 4 == 0, "precision 8 needs the innermost coarse dimension to be a multiple of 4: use precision 32");
        st->V = st->Zd % 16 == 0 ? 16 : 4;
    }

    // classes sorted by colour (stable), classOf = inverse
    std::vector<int> order(bestN);
    for (int i = 0; i < bestN; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bestColour[a] < bestColour[b]; });
    st->classOf.assign(bestN, -1);
    for (int q = 0; q < bestN; q++) st->classOf[order[q]] = q;
    st->colourClassStart.assign(bestC + 1, 0);
    for (int i = 0; i < bestN; i++) st->colourClassStart[bestColour[i] + 1]++;
    for (int c = 0; c < bestC; c++) st->colourClassStart[c + 1] += st->colourClassStart[c];
    st->classes.resize(bestN);
    links.assign((size_t)bestN * MAXLINK, SLinkD{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
    st->cXs.assign((size_t)bestN * MAXLINK, 0);
    st->cYs.assign((size_t)bestN * MAXLINK, 0);
    const int JW = d->model == 1 ? 1 : 9;
    Jt.assign((size_t)bestN * MAXLINK * JW, 0.0);
    for (int q = 0; q < bestN; q++) {
        int id = order[q];
        int o = id % no, c = (id / no) % pz, b = (id / (no * pz)) % py, a = id / (no * pz * py);
        SClassD &cl = st->classes[q];
        cl.a = a; cl.b = b; cl.c = c; cl.o = o; cl.colour = bestColour[id]; cl.nlink = (int)tm[o].size(); cl.pad = 0;
        cl.S = d->S[o];
        for (int k = 0; k < 3; k++) cl.D[k] = (d->D && d->model != 1) ? d->D[3 * o + k] : 0.0;
        int nlow = 0, nreal = 0;
        for (int k = 0; k < cl.nlink; k++) {
            const Tmpl &t = tm[o][k];
            SLinkD &l = links[(size_t)q * MAXLINK + k];
            int na = a + t.d[0], nb = b + t.d[1], nc = c + t.d[2];
            auto fdiv = [](int v, int m) { return (v >= 0) ? v / m : -((-v + m - 1) / m); };
            int cX = fdiv(na, px), cY = fdiv(nb, py), cZ = fdiv(nc, pz);
            int nid = ((mod(na, px) * py + mod(nb, py)) * pz + mod(nc, pz)) * no + t.o2;
            l.qn = st->classOf[nid];
            l.cX = mod(cX, st->Xd); l.cY = mod(cY, st->Yd);
            int z = mod(cZ, st->Zd); if (z > st->Zd / 2) z -= st->Zd;
            l.cZ = z;
            { int v = mod(cX, st->Xd); if (v > st->Xd / 2) v -= st->Xd; st->cXs[(size_t)q * MAXLINK + k] = v; }
            { int v = mod(cY, st->Yd); if (v > st->Yd / 2) v -= st->Yd; st->cYs[(size_t)q * MAXLINK + k] = v; }
            l.self = t.self ? 1 : 0;
            l.fwd = t.fwd ? 1 : 0;
            l.o2 = t.o2; l.dx = mod(t.d[0], L[0]); l.dy = mod(t.d[1], L[1]); l.dz = mod(t.d[2], L[2]);
            l.low = (!t.self && bestColour[nid] < bestColour[id]) ? 1 : 0;
            if (!t.self) { nreal++; nlow += l.low; }
            for (int e = 0; e < JW; e++) Jt[((size_t)q * MAXLINK + k) * JW + e] = t.J[e];
        }
        cl.lowmode = nlow == 0 ? 0 : (nlow == nreal ? 1 : 2);
        // links to lower colours first: the field of the already final neighbours (the bond-energy part of the fused
        // measurement) is then simply the running field sum after the first nlow links - no second accumulator per link
        {
            std::vector<int> perm(cl.nlink);
            for (int k = 0; k < cl.nlink; k++) perm[k] = k;
            // within each half: aligned rows, then Z shift -1, +1, others (runs with a fixed shift, struct_pass.cuh: PassTable::gend)
            auto zkind = [](int cz) { return cz == 0 ? 0 : cz == -1 ? 1 : cz == 1 ? 2 : 3; };
            std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) {
                const SLinkD &lx = links[(size_t)q * MAXLINK + x], &ly = links[(size_t)q * MAXLINK + y];
                if (lx.low != ly.low) return lx.low > ly.low;
                return zkind(lx.cZ) < zkind(ly.cZ);
            });
            std::vector<SLinkD> l2(cl.nlink);
            std::vector<int> cx2(cl.nlink), cy2(cl.nlink);
            std::vector<double> j2((size_t)cl.nlink * JW);
            for (int k = 0; k < cl.nlink; k++) {
                const size_t src = (size_t)q * MAXLINK + perm[k];
                l2[k] = links[src]; cx2[k] = st->cXs[src]; cy2[k] = st->cYs[src];
                for (int e = 0; e < JW; e++) j2[(size_t)k * JW + e] = Jt[src * JW + e];
            }
            for (int k = 0; k < cl.nlink; k++) {
                const size_t dst = (size_t)q * MAXLINK + k;
                links[dst] = l2[k]; st->cXs[dst] = cx2[k]; st->cYs[dst] = cy2[k];
                for (int e = 0; e < JW; e++) Jt[dst * JW + e] = j2[(size_t)k * JW + e];
            }
        }
        cl.pad = nlow;   // number of leading links into lower colours
    }
    st->linksHost = links;
    st->JHost = Jt;
    // ---- per-colour pass tables for the fast kernel ----
    st->fastOK.assign(bestC, 0);
    st->passTables.assign(bestC, std::vector<char>());
    auto build_pass = [&](auto realTag, int colour) {
        typedef decltype(realTag) real;
        int q0 = st->colourClassStart[colour], nqc = st->colourClassStart[colour + 1] - q0;
        if (nqc < 1 || nqc > PT_MAXC || st->hasSelf) return;
        int nl = 0;
        for (int j = 0; j < nqc; j++) nl = std::max(nl, st->classes[q0 + j].nlink);
        if (nl < 1) return;
        int G = 1;
        int nlp = nl;
        if (nlp > PT_MAXL) return;
        std::vector<char> buf(sizeof(PassTable<real>), 0);
        PassTable<real> *P = reinterpret_cast<PassTable<real> *>(buf.data());
        P->nl = nlp; P->nqc = nqc;
        P->uniformJ = 1;
        for (int j = 0; j < nqc; j++)
            for (int k = 0; k < st->classes[q0 + j].nlink; k++)
                for (int c = 0; c < 3 && JW == 9; c++)
                    if (Jt[((size_t)(q0 + j) * MAXLINK + k) * JW + c] != Jt[(size_t)q0 * MAXLINK * JW + c]) P->uniformJ = 0;
        for (int j = 0; j < nqc; j++) {
            const SClassD &cl = st->classes[q0 + j];
            P->ca[j] = cl.a; P->cb[j] = cl.b; P->cc[j] = cl.c; P->co[j] = cl.o; P->lowmode[j] = cl.lowmode; P->nlow[j] = cl.pad;
            P->S[j] = (real)std::fabs(cl.S);
            for (int e = 0; e < 3; e++) P->D[j][e] = (real)cl.D[e];
            {   // run boundaries; the zero-J pads of a shorter class are not visited at all
                auto zkind = [](int cz) { return cz == 0 ? 0 : cz == -1 ? 1 : cz == 1 ? 2 : 3; };
                int k = 0;
                for (int half = 0; half < 2; half++)
                    for (int kind = 0; kind < 4; kind++) {
                        const int hend = half == 0 ? cl.pad : cl.nlink;
                        while (k < hend && zkind(links[(size_t)(q0 + j) * MAXLINK + k].cZ) == kind) k++;
                        P->gend[j][4 * half + kind] = k;
                    }
                if (k != cl.nlink) return;   // not sorted as expected: leave this colour to the generic kernel
            }
            for (int k = 0; k < nlp; k++) {
                PLink<real> &L = P->L[j][k];
                if (k >= cl.nlink) { L.delta = 0; L.mxp = L.mxm = L.myp = L.mym = 0; L.cZ = 0; L.low = 0; continue; }   // zero-J pad: own cell
                size_t li = (size_t)(q0 + j) * MAXLINK + k;
                int cx = st->cXs[li], cy = st->cYs[li];
                if (cx < -1 || cx > 1 || cy < -1 || cy > 1) return;
                L.delta = (links[li].qn - (q0 + j)) * st->ncellc + (cx * st->Yd + cy) * st->Zd;
                L.mxp = cx > 0; L.mxm = cx < 0; L.myp = cy > 0; L.mym = cy < 0;
                L.cZ = links[li].cZ; L.low = links[li].low;
                for (int e = 0; e < JW; e++) L.J[e] = (real)Jt[li * JW + e];
            }
        }
        st->passTables[colour] = buf;
        st->fastOK[colour] = (char)G;
    };
    for (int c = 0; c < bestC && s->prec != 8; c++) {
        if (s->prec == 32) build_pass(float(0), c); else build_pass(double(0), c);
    }
    if (s->prec == 8) {
        // ising8.cuh: one exchange constant and one |S| for the whole lattice, every bond image within one row / one cell
        const char *why = "precision 8 (int8 Ising spins) needs ONE exchange constant, one |S|, no self-images and bonds within the "
                          "neighbouring coarse cells: use precision 32 for this lattice";
        MCG_REQUIRE(!st->hasSelf && d->block_spin == 0, d->block_spin ? "block_spin statistics are not available at precision 8" : why);
        const double J0 = Jt.empty() ? 0.0 : Jt[0], S0 = std::fabs(st->classes[0].S);
        st->i8Tables.assign(bestC, I8Table());
        for (int c = 0; c < bestC; c++) {
            I8Table &T = st->i8Tables[c];
            const int q0 = st->colourClassStart[c], nqc = st->colourClassStart[c + 1] - q0;
            MCG_REQUIRE(nqc >= 1 && nqc <= PT_MAXC, why);
            T.nqc = nqc; T.JS2 = J0 * S0 * S0; T.S = S0;
            for (int j = 0; j < nqc; j++) {
                const SClassD &cl = st->classes[q0 + j];
                I8Class &C = T.c[j];
                MCG_REQUIRE(std::fabs(cl.S) == S0 && cl.nlink >= 1 && cl.nlink <= I8_MAXZ, why);
                C.nl = cl.nlink; C.nlow = cl.pad; C.lowmode = cl.lowmode; C.ca = cl.a; C.cb = cl.b; C.cc = cl.c; C.co = cl.o;
                for (int k = 0; k < cl.nlink; k++) {
                    const size_t li = (size_t)(q0 + j) * MAXLINK + k;
                    const int cx = st->cXs[li], cy = st->cYs[li], cz = links[li].cZ;
                    MCG_REQUIRE(Jt[li] == J0 && cx >= -1 && cx <= 1 && cy >= -1 && cy <= 1 && cz >= -1 && cz <= 1, why);
                    C.delta[k] = (links[li].qn - (q0 + j)) * st->ncellc + (cx * st->Yd + cy) * st->Zd;
                    C.wx[k] = (signed char)cx; C.wy[k] = (signed char)cy; C.cz[k] = (signed char)cz;
                }
            }
        }
    }
    // orbital groups
    s->nG = (d->model != 1 && d->ngroup > 0) ? d->ngroup : 0;
    MCG_REQUIRE(s->nG <= 16, "at most 16 orbital groups on the structured path");
    if (s->nG > 0) {
        MCG_REQUIRE(d->group_mask, "group_mask is NULL");
        st->gmaskHost.assign(d->group_mask, d->group_mask + (size_t)s->nG * no);
        st->groupInSC = d->group_in_sc != 0;
    }
    // block-spin statistics (opt-in): the bond templates in the reference's order with pre-reduced offsets
    if (d->block_spin) {
        for (int k = 0; k < 3; k++)
            MCG_REQUIRE(L[k] == 1 || L[k] % 2 == 0, "block_spin needs even supercell dimensions on the structured path (odd sizes: table path)");
        for (int mult = 1; mult <= 2; mult++)      // every image of the original and of the doubled bonds must be a distinct site
            for (int o = 0; o < no; o++) {
                std::vector<std::array<int, 4>> seen;
                auto visit = [&](int o2, const int *dd, int sgn) {
                    std::array<int, 4> e = {o2, mod(sgn * mult * dd[0], L[0]), mod(sgn * mult * dd[1], L[1]), mod(sgn * mult * dd[2], L[2])};
                    bool self = o2 == o && e[1] == 0 && e[2] == 0 && e[3] == 0;
                    MCG_REQUIRE(!self && std::find(seen.begin(), seen.end(), e) == seen.end(),
                                "block_spin: bond images coincide on this supercell (merged links): use the table path");
                    seen.push_back(e);
                };
                for (int b = 0; b < d->nbond; b++) {
                    int dd[3];
                    conv(d->bonds[b].d, dd);
                    if (d->bonds[b].src == o) visit(d->bonds[b].tgt, dd, 1);
                    if (d->bonds[b].tgt == o) visit(d->bonds[b].src, dd, -1);
                }
                MCG_REQUIRE((int)seen.size() <= MAXLINK, "more than 32 links per site");
            }
        st->rgOn = true;
        for (int k = 0; k < 3; k++) st->rgN[k] = (L[k] + 1) / 2;
        st->nR = st->rgN[0] * st->rgN[1] * st->rgN[2] * no;
        const int nb = d->nbond;
        std::vector<std::array<int, 3>> dd(nb);
        for (int b = 0; b < nb; b++) conv(d->bonds[b].d, dd[b].data());
        auto sym = [&](int v, int k) { int m = mod(v, L[k]); return m > L[k] / 2 ? m - L[k] : m; };
        for (int k = 0; k < 3; k++) {   // edge zones: wide enough for every image used by either list
            int R = 0;
            for (int b = 0; b < nb; b++) R = std::max({R, std::abs(sym(dd[b][k], k)), std::abs(sym(2 * dd[b][k], k))});
            st->rgZ.R[k] = R;
            st->rgZ.nz[k] = L[k] <= 2 * R + 1 ? L[k] : 2 * R + 1;
        }
        // entries of every orbital in bond order: (bond, forward | backward)
        struct HostEnt { int b, act; };
        std::vector<std::vector<HostEnt>> ents(no);
        st->rgEntHost.assign((size_t)no * MAXLINK, RgEntD{0, {0, 0, 0}});
        st->rgNentHost.assign(no, 0);
        for (int o = 0; o < no; o++) {
            for (int b = 0; b < nb; b++) {
                if (d->bonds[b].src == o) ents[o].push_back({b, 0});
                if (d->bonds[b].tgt == o) ents[o].push_back({b, 1});
            }
            st->rgNentHost[o] = (int)ents[o].size();
            for (size_t e = 0; e < ents[o].size(); e++) {
                RgEntD &E = st->rgEntHost[(size_t)o * MAXLINK + e];
                const int b = ents[o][e].b, sgn = ents[o][e].act ? -1 : 1;
                E.o2 = ents[o][e].act ? d->bonds[b].src : d->bonds[b].tgt;
                for (int k = 0; k < 3; k++) E.off[k] = mod(sgn * 2 * dd[b][k], L[k]);
            }
        }
        // insertion order of the original and of the doubled-bond list for one representative site per zone:
        // key = (id of the bond's source site, bond, direction)   (Lattice.py:247-271 visit sites in id order)
        const int nsig = st->rgZ.nz[0] * st->rgZ.nz[1] * st->rgZ.nz[2] * no;
        st->rgPermHost.assign((size_t)nsig * MAXLINK, 0);
        auto rep = [&](int zone, int k) {
            const int R = st->rgZ.R[k];
            if (st->rgZ.nz[k] == L[k]) return zone;
            return zone < R ? zone : (zone == R ? R : L[k] - R + (zone - R - 1));
        };
        auto sid = [&](int x, int y, int z, int o) { return (long long)((mod(x, L[0]) * L[1] + mod(y, L[1])) * L[2] + mod(z, L[2])) * no + o; };
        for (int zx = 0; zx < st->rgZ.nz[0]; zx++) for (int zy = 0; zy < st->rgZ.nz[1]; zy++) for (int zz = 0; zz < st->rgZ.nz[2]; zz++)
            for (int o = 0; o < no; o++) {
                const int x = rep(zx, 0), y = rep(zy, 1), z = rep(zz, 2);
                const int n = (int)ents[o].size();
                std::vector<long long> kO(n), kR(n);
                for (int e = 0; e < n; e++) {
                    const int b = ents[o][e].b;
                    if (!ents[o][e].act) kO[e] = kR[e] = (sid(x, y, z, o) * nb + b) * 2;
                    else {
                        kO[e] = (sid(x - dd[b][0], y - dd[b][1], z - dd[b][2], d->bonds[b].src) * nb + b) * 2 + 1;
                        kR[e] = (sid(x - 2 * dd[b][0], y - 2 * dd[b][1], z - 2 * dd[b][2], d->bonds[b].src) * nb + b) * 2 + 1;
                    }
                }
                int *perm = st->rgPermHost.data() + (size_t)((((zx * st->rgZ.nz[1] + zy) * st->rgZ.nz[2] + zz) * no) + o) * MAXLINK;
                std::vector<int> slotJ(n), slotE(n);
                for (int e = 0; e < n; e++) {
                    int rO = 0, rR = 0;
                    for (int f = 0; f < n; f++) { rO += kO[f] < kO[e]; rR += kR[f] < kR[e]; }
                    slotJ[rO] = 2 * ents[o][e].b + ents[o][e].act;   // exchange of the k-th ORIGINAL link ...
                    slotE[rR] = e;                                   // ... paired with the k-th coarse neighbour
                }
                for (int k = 0; k < n; k++) perm[k] = (slotJ[k] << 8) | slotE[k];
            }
        st->rgJHost.assign((size_t)std::max(1, nb) * 9, 0.0);
        for (int b = 0; b < nb; b++)
            for (int k = 0; k < 9; k++) st->rgJHost[(size_t)b * 9 + k] = d->model == 1 ? (k == 0 ? d->bonds[b].J[0] : 0.0) : d->bonds[b].J[k];
        for (int o = 0; o < no; o++) {
            st->rgSDHost.push_back(d->S[o]);
            for (int k = 0; k < 3; k++) st->rgSDHost.push_back((d->D && d->model != 1) ? d->D[3 * o + k] : 0.0);
        }
    }
    // measurement templates
    st->pair_s = d->pair_s; st->pair_t = d->pair_t;
    conv(d->pair_d, st->pair_d);
    for (int k = 0; k < 3; k++) st->pair_d[k] = mod(st->pair_d[k], L[k]);
    st->selfPairs = st->pair_s == st->pair_t && st->pair_d[0] == 0 && st->pair_d[1] == 0 && st->pair_d[2] == 0;
    s->nLat = L[0] * L[1] * L[2];
    if (st->rgOn) {   // pairs with a chosen i member, a chosen j member, both (heisenbergLib.c:748-790)
        double cells = (double)st->rgN[0] * st->rgN[1] * st->rgN[2];
        bool evenShift = !((st->pair_d[0] | st->pair_d[1] | st->pair_d[2]) & 1);
        st->rg_ci = cells; st->rg_cj = cells; st->rg_cij = evenShift ? cells : 0.0;
    }
    st->ncircuit = d->model == 3 ? d->ncircuit : 0;
    s->nTri = st->ncircuit * s->nLat;
    for (int i = 0; i < st->ncircuit * 3; i++) {
        const int32_t *e = d->circuits + 4 * i;
        MCG_REQUIRE(e[0] >= 0 && e[0] < no, "circuit orbital out of range");
        int dd[3];
        conv(e + 1, dd);
        st->circuits.push_back(e[0]);
        for (int k = 0; k < 3; k++) st->circuits.push_back(mod(dd[k], L[k]));
    }
    // distinct vertices of the circuit templates (cell-relative) and triangles as indices into them
    std::vector<int> tverts, ttris;
    for (int i = 0; i < st->ncircuit * 3; i++) {
        const int *e = st->circuits.data() + 4 * i;
        int found = -1;
        for (int k = 0; k < (int)tverts.size() / 4; k++)
            if (tverts[4 * k] == e[0] && tverts[4 * k + 1] == e[1] && tverts[4 * k + 2] == e[2] && tverts[4 * k + 3] == e[3]) found = k;
        if (found < 0) { found = (int)tverts.size() / 4; tverts.insert(tverts.end(), e, e + 4); }
        ttris.push_back(found);
    }
    st->nvert = (int)tverts.size() / 4;
    st->tvertsHost = tverts; st->ttrisHost = ttris;
    s->S_host.clear();
    s->maxL = 0;
    for (int o = 0; o < no; o++) s->maxL = std::max(s->maxL, (int)tm[o].size());

    s->st = stp.release();
}

static void structured_create_impl(mcg_system *s, const mcg_lattice_desc *d, const SlabPlan *plan) {
    std::vector<SLinkD> links;
    std::vector<double> Jt;
    structured_build_host(s, d, links, Jt, plan);
    StructuredSystem *st = s->st;
    // ---- upload ----
    auto up = [&](const void *src, size_t bytes) { void *p = pool_alloc(std::max<size_t>(bytes, 16)); if (bytes) MCG_CUDA(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice)); return p; };
    st->d_classes = (SClassD *)up(st->classes.data(), st->classes.size() * sizeof(SClassD));
    st->d_links = (SLinkD *)up(links.data(), links.size() * sizeof(SLinkD));
    if (s->prec == 64) st->d_J = up(Jt.data(), Jt.size() * sizeof(double));
    else { std::vector<float> jf(Jt.begin(), Jt.end()); st->d_J = up(jf.data(), jf.size() * sizeof(float)); }
    st->d_classOf = (int *)up(st->classOf.data(), st->classOf.size() * sizeof(int));
    st->d_circuits = (int *)up(st->circuits.data(), st->circuits.size() * sizeof(int));
    st->d_gmask = (int *)up(st->gmaskHost.data(), st->gmaskHost.size() * sizeof(int));
    st->d_tverts = (int *)up(st->tvertsHost.data(), st->tvertsHost.size() * sizeof(int));
    st->d_ttris = (int *)up(st->ttrisHost.data(), st->ttrisHost.size() * sizeof(int));
    if (st->rgOn) {
        st->d_rgEnt = (RgEntD *)up(st->rgEntHost.data(), st->rgEntHost.size() * sizeof(RgEntD));
        st->d_rgNent = (int *)up(st->rgNentHost.data(), st->rgNentHost.size() * sizeof(int));
        st->d_rgPerm = (int *)up(st->rgPermHost.data(), st->rgPermHost.size() * sizeof(int));
        st->d_rgJ = (double *)up(st->rgJHost.data(), st->rgJHost.size() * sizeof(double));
        st->d_rgSD = (double *)up(st->rgSDHost.data(), st->rgSDHost.size() * sizeof(double));
        st->d_ms = (double *)pool_alloc((size_t)s->R * st->nR * 3 * sizeof(double));
        st->d_rsums = (double *)pool_alloc((size_t)s->R * NRS * sizeof(double));
        MCG_CUDA(cudaMemset(st->d_rsums, 0, (size_t)s->R * NRS * sizeof(double)));
    }
    if (s->prec == 32 && s->NC == 3 && s->fullJ) {
        // the asynchronous pipeline's view of the pass tables: per class the links in pass order (16-byte records the kernel
        // copies into shared memory), then one record for the item itself
        std::vector<AsTab> at((size_t)st->nclass * (PT_MAXL + 1));
        memset(at.data(), 0, at.size() * sizeof(AsTab));
        for (size_t c = 0; c < st->passTables.size(); c++) {
            if (st->passTables[c].empty()) continue;
            const PassTable<float> &P = *reinterpret_cast<const PassTable<float> *>(st->passTables[c].data());
            const int q0 = st->colourClassStart[c];
            for (int j = 0; j < P.nqc; j++)
                for (int k = 0; k < P.gend[j][7]; k++) {
                    const PLink<float> &L = P.L[j][k];
                    AsTab &t = at[(size_t)(q0 + j) * (PT_MAXL + 1) + k];
                    for (int e = 0; e < 4; e++) t.j[e] = make_float4(L.J[2 * e], L.J[2 * e], L.J[2 * e + 1], L.J[2 * e + 1]);
                    t.j[4] = make_float4(L.J[8], L.J[8], 0.f, 0.f);
                    t.a = make_int4(L.delta, L.mxp - L.mxm, L.myp - L.mym, L.cZ);
                }
        }
        st->d_asTab = (AsTab *)up(at.data(), at.size() * sizeof(AsTab));
    }
    size_t cs = (size_t)s->R * st->nclass * 4 * sizeof(double);
    st->d_classSums = (double *)pool_alloc(cs);
    MCG_CUDA(cudaMemset(st->d_classSums, 0, cs));
    s->d_spin = pool_alloc((size_t)s->R * s->NC * s->N * s->real_size());
}

void structured_create(mcg_system *s, const mcg_lattice_desc *d) { structured_create_impl(s, d, nullptr); }

// host-only check used by the CPU tests: build the class tables of a descriptor and compile the
// specialised pass kernels of every colour with NVRTC (no device needed up to the cubin)
static std::string jit_key_hex(const std::string &src) {
    char b[32];
    snprintf(b, sizeof b, "%016llx", (unsigned long long)cache_key(src));
    return b;
}

int structured_jit_check(const mcg_lattice_desc *d, int precision, std::string &report) {
    mcg_system tmp;
    tmp.prec = precision;
    tmp.R = 1;
    std::vector<SLinkD> links;
    std::vector<double> Jt;
    structured_build_host(&tmp, d, links, Jt);
    int ncompiled = 0;
    std::ostringstream o;
    o << "colours=" << tmp.C << " classes=" << tmp.st->nclass << " period=" << tmp.st->p[0] << "x" << tmp.st->p[1] << "x" << tmp.st->p[2]
      << " V=" << tmp.st->V << "\n";
    for (int c = 0; c < tmp.C && precision != 8; c++) {
        if (!tmp.st->fastOK[c] || tmp.st->V == 1 || !jit_worthwhile(&tmp, c)) { o << "colour " << c << ": not eligible for specialisation\n"; continue; }
        std::string log;
        const std::string src = jit_prologue(&tmp, c, false);
        std::vector<char> cubin = jit_compile_cubin(src, log);
        o << "colour " << c << ": module " << jit_key_hex(src) << " cubin " << cubin.size() << " bytes" << (log.empty() ? "" : " log: " + log.substr(0, 1500)) << "\n";
        if (cubin.empty()) { report = o.str(); return -1; }
        ncompiled++;
    }
    for (int c = 0; c < tmp.C && precision == 8; c++) {
        std::string log;
        const std::string src = jit_i8_prologue(&tmp, c, false);
        std::vector<char> cubin = jit_compile_cubin(src, log);
        o << "colour " << c << " (int8): module " << jit_key_hex(src) << " cubin " << cubin.size() << " bytes" << (log.empty() ? "" : " log: " + log.substr(0, 1500)) << "\n";
        if (cubin.empty()) { report = o.str(); return -1; }
        ncompiled++;
    }
    if (jit_topo_worthwhile(tmp.st) && tmp.NC == 3) {
        std::string log;
        const std::string src = jit_topo_prologue(&tmp);
        std::vector<char> cubin = jit_compile_cubin(src, log);
        o << "topological charge: module " << jit_key_hex(src) << " cubin " << cubin.size() << " bytes" << (log.empty() ? "" : " log: " + log.substr(0, 1500)) << "\n";
        if (cubin.empty()) { report = o.str(); return -1; }
    }
    report = o.str();
    return ncompiled;
}

void structured_destroy(StructuredSystem *st) {
    if (!st) return;
    void *bufs[] = {st->d_classes, st->d_links, st->d_J, st->d_classOf, st->d_circuits, st->d_classSums, st->d_stage, st->d_tverts, st->d_ttris, st->d_gmask, st->d_rgEnt, st->d_rgNent, st->d_rgPerm, st->d_rgJ, st->d_rgSD, st->d_ms, st->d_rsums, st->d_asTab};
    for (void *b : bufs) pool_free(b);
    pool_free(st->d_slabBuf);
    if (st->slabStream) { cudaStreamSynchronize(st->slabStream); cudaStreamDestroy(st->slabStream); }
    if (st->slabEvA) cudaEventDestroy(st->slabEvA);
    if (st->slabEvB) cudaEventDestroy(st->slabEvB);
    if (st->slabComm && nccl_api().ok) nccl_api().commDestroy((ncclComm_t)st->slabComm);
    delete st;
}

static void slab_exchange_classes(mcg_system *s, int q0, int nqc);
static void slab_allreduce_sums(mcg_system *s);

template <typename F> static void sdispatch(const mcg_system *s, F &&f) {
    bool d = s->prec == 64, fj = s->fullJ;
    switch (s->NC) {
    case 1: d ? f.template operator()<1, double, false>() : f.template operator()<1, float, false>(); break;
    case 2:
        if (d) fj ? f.template operator()<2, double, true>() : f.template operator()<2, double, false>();
        else fj ? f.template operator()<2, float, true>() : f.template operator()<2, float, false>();
        break;
    default:
        if (d) fj ? f.template operator()<3, double, true>() : f.template operator()<3, double, false>();
        else fj ? f.template operator()<3, float, true>() : f.template operator()<3, float, false>();
    }
}

void structured_init_spins(mcg_system *s, double flunc) {
    StructArgs a = struct_args(s);
    dim3 g((s->N + 255) / 256, s->R);
    s->launches++;
    if (s->prec == 8) k_i8_init<<<g, 256, 0, s->stream>>>(a);
    else sdispatch(s, [&]<int NC, typename real, bool FJ>() { k_struct_init<NC, real><<<g, 256, 0, s->stream>>>(a, flunc, s->st->LxGlobal); });
    MCG_CUDA(cudaGetLastError());
}

static double *stage(mcg_system *s) {
    if (!s->st->d_stage) s->st->d_stage = (double *)pool_alloc(3 * (size_t)s->N * sizeof(double));
    return s->st->d_stage;
}

void structured_set_spins(mcg_system *s, int r, const double *spins) {
    StructArgs a = struct_args(s);
    double *buf = stage(s);
    size_t n = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    MCG_CUDA(cudaMemcpyAsync(buf, spins, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    s->launches++;
    if (s->prec == 8) k_i8_frame<false><<<(s->N + 255) / 256, 256, 0, s->stream>>>(a, r, buf);
    else sdispatch(s, [&]<int NC, typename real, bool FJ>() { k_struct_frame<NC, real, false><<<(s->N + 255) / 256, 256, 0, s->stream>>>(a, r, buf); });
    MCG_CUDA(cudaGetLastError());
    MCG_CUDA(cudaStreamSynchronize(s->stream));
    // a slab takes its local planes, ghosts included, as given: mcg_slab_sync() brings the ghosts up to date once every rank has
    // set every replica (it is collective)
}

void structured_get_spins(mcg_system *s, int r, double *spins) {
    StructArgs a = struct_args(s);
    double *buf = stage(s);
    size_t n = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    s->launches++;
    if (s->prec == 8) k_i8_frame<true><<<(s->N + 255) / 256, 256, 0, s->stream>>>(a, r, buf);
    else sdispatch(s, [&]<int NC, typename real, bool FJ>() { k_struct_frame<NC, real, true><<<(s->N + 255) / 256, 256, 0, s->stream>>>(a, r, buf); });
    MCG_CUDA(cudaGetLastError());
    MCG_CUDA(cudaMemcpyAsync(spins, buf, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    MCG_CUDA(cudaStreamSynchronize(s->stream));
}

void structured_rng_layout(const mcg_system *s, int32_t *stride, int32_t *group) {
    const StructuredSystem *st = s->st;
    if (st->V > 1) { *stride = st->p[2] * st->norb; *group = st->V; }   // ItemWords (rng.cuh): ids id0 + v * pz * norb, v < V
    else { *stride = 0; *group = 0; }
}

void structured_colour_order(const mcg_system *s, int32_t *order) {
    const StructuredSystem *st = s->st;
    for (int p = 0; p < s->N; p++) {
        int q = p / st->ncellc, cell = p % st->ncellc;
        int Z = cell % st->Zd, Y = (cell / st->Zd) % st->Yd, X = cell / (st->Zd * st->Yd);
        const SClassD &c = st->classes[q];
        int x = X * st->p[0] + c.a, y = Y * st->p[1] + c.b, z = Z * st->p[2] + c.c;
        order[p] = ((x * st->L[1] + y) * st->L[2] + z) * st->norb + c.o;
    }
}

// rowFrom/rowTo: update only these rows (a slab launches its two boundary planes ahead of its interior); -1 = all own rows
template <int MODE> static void launch_pass(mcg_system *s, int colour, uint64_t sweep, double pAtt, int rowFrom = -1, int rowTo = -1) {
    StructuredSystem *st = s->st;
    StructArgs a = struct_args(s);
    if (rowFrom >= 0) { a.rowLo = rowFrom; a.rowHi = rowTo; }
    if (a.rowHi <= a.rowLo) return;
    int q0 = st->colourClassStart[colour], nqc = st->colourClassStart[colour + 1] - q0;
    if (nqc == 0) return;
    int Zc = st->Zd / st->V;
    int bx = 1;
    // int8 items (16 sites): four items of a row per thread amortise the per-row address set-up
    const int bxTarget = s->prec == 8 ? std::max(1, Zc / 4) : Zc;
    while (bx < bxTarget && bx < 64) bx <<= 1;
    int by = 256 / bx;
    int iters = 8;
    const int nrowsOwn = a.rowHi - a.rowLo;          // all rows, this rank's slab of them, or a part of the slab
    auto nblocks = [&](int it) { return (long long)s->R * nqc * ((nrowsOwn + by * it - 1) / (by * it)); };
    while (iters > 1 && nblocks(iters) < 2368) iters >>= 1;   // >= 16 blocks per SM on 148 SMs when the lattice allows
    int rowsPerBlock = by * iters;
    int nrb = (nrowsOwn + rowsPerBlock - 1) / rowsPerBlock;
    dim3 block(bx, by), grid((unsigned)(s->R * nrb * nqc));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = s->profilePasses && MODE != 2;
    if (prof) { MCG_CUDA(cudaEventCreate(&e0)); MCG_CUDA(cudaEventCreate(&e1)); MCG_CUDA(cudaEventRecord(e0, s->stream)); }
    s->launches++;
    if (s->prec == 8) {   // int8 Ising pass (ising8.cuh): items of 16 or 4 sites
        const I8Table &T8 = st->i8Tables[colour];
        if constexpr (MODE != 2) {
            if (jit_launch_pass(s, colour, MODE, a, q0, rowsPerBlock, nrb, sweep, pAtt, grid, block)) {
                if (prof) { MCG_CUDA(cudaEventRecord(e1, s->stream)); s->passEvents.emplace_back(e0, e1); }
                return;
            }
        }
        bool small = true;
        for (int j = 0; j < T8.nqc; j++) small = small && T8.c[j].nl <= RtI8Class::NLMAX;
        auto go = [&]<bool PARTIAL>() {
            if (st->V == 16) {
                if (small) k_i8_pass<MODE, PARTIAL, 4, true><<<grid, block, 0, s->stream>>>(a, T8, q0, rowsPerBlock, nrb, sweep, pAtt);
                else k_i8_pass<MODE, PARTIAL, 4, false><<<grid, block, 0, s->stream>>>(a, T8, q0, rowsPerBlock, nrb, sweep, pAtt);
            } else {
                if (small) k_i8_pass<MODE, PARTIAL, 1, true><<<grid, block, 0, s->stream>>>(a, T8, q0, rowsPerBlock, nrb, sweep, pAtt);
                else k_i8_pass<MODE, PARTIAL, 1, false><<<grid, block, 0, s->stream>>>(a, T8, q0, rowsPerBlock, nrb, sweep, pAtt);
            }
        };
        if (MODE != 2 && pAtt < 1.0) go.template operator()<true>(); else go.template operator()<false>();
        if (prof) { MCG_CUDA(cudaEventRecord(e1, s->stream)); s->passEvents.emplace_back(e0, e1); }
        return;
    }
    const int G = (MODE != 2 && st->V > 1 && !getenv("MCG_NO_FAST")) ? st->fastOK[colour] : 0;
    if (G) {
        bool launched = false;
        if constexpr (MODE != 2) launched = jit_launch_pass(s, colour, MODE, a, q0, rowsPerBlock, nrb, sweep, pAtt, grid, block);
        // full 3x3 tensors on fp32 Heisenberg items that the specialiser did not take (dipole stencils): asynchronous link pipeline
        if constexpr (MODE != 2) {
            const bool noAsync = getenv("MCG_NO_ASYNC") != nullptr;   // read per launch: tests switch it inside one process
            if (!launched && !noAsync && st->d_asTab && st->V == 4 && block.x * block.y == 256) {
                constexpr size_t dyn = (size_t)ASYNC_D * 3 * 256 * sizeof(float4);
                const PassTable<float> &P = *reinterpret_cast<const PassTable<float> *>(st->passTables[colour].data());
                auto go = [&]<bool PARTIAL>() {
                    auto kern = k_struct_async<MODE, PARTIAL>;
                    // function attributes are per device: set once per system (a process may hold systems on several GPUs)
                    if (!st->asyncAttrSet[MODE][PARTIAL]) {
                        MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                        st->asyncAttrSet[MODE][PARTIAL] = true;
                    }
                    kern<<<grid, block, dyn, s->stream>>>(a, P, st->d_asTab, q0, rowsPerBlock, nrb, sweep, (float)pAtt);
                };
                if (pAtt < 1.0) go.template operator()<true>(); else go.template operator()<false>();
                launched = true;
            }
        }
        if (!launched)
            sdispatch(s, [&]<int NC, typename real, bool FJ>() {
                if constexpr (MODE != 2) {
                    constexpr int VV = sizeof(real) == 4 ? 4 : 2;
                    const PassTable<real> &P = *reinterpret_cast<const PassTable<real> *>(st->passTables[colour].data());
                    if (pAtt < 1.0) k_struct_fast<NC, real, FJ, MODE, VV, true><<<grid, block, 0, s->stream>>>(a, P, q0, rowsPerBlock, nrb, sweep, (real)pAtt);
                    else k_struct_fast<NC, real, FJ, MODE, VV, false><<<grid, block, 0, s->stream>>>(a, P, q0, rowsPerBlock, nrb, sweep, (real)pAtt);
                }
            });
    } else
    sdispatch(s, [&]<int NC, typename real, bool FJ>() {
        if (st->V == 1) k_struct<NC, real, FJ, MODE, 1><<<grid, block, 0, s->stream>>>(a, q0, nqc, rowsPerBlock, nrb, sweep, (real)pAtt);
        else if constexpr (sizeof(real) == 4) k_struct<NC, real, FJ, MODE, 4><<<grid, block, 0, s->stream>>>(a, q0, nqc, rowsPerBlock, nrb, sweep, (real)pAtt);
        else k_struct<NC, real, FJ, MODE, 2><<<grid, block, 0, s->stream>>>(a, q0, nqc, rowsPerBlock, nrb, sweep, (real)pAtt);
    });
    if (prof) { MCG_CUDA(cudaEventRecord(e1, s->stream)); s->passEvents.emplace_back(e0, e1); }
}

static void launch_block_spin(mcg_system *s, const StructArgs &a) {
    StructuredSystem *st = s->st;
    RgS g;
    g.nR = st->nR; g.n0 = st->rgN[0]; g.n1 = st->rgN[1]; g.n2 = st->rgN[2]; g.nLat = s->nLat;
    g.fullJ = s->fullJ ? 1 : 0;
    g.Z = st->rgZ; g.ent = st->d_rgEnt; g.nent = st->d_rgNent; g.perm = st->d_rgPerm; g.J = st->d_rgJ;
    g.SD = st->d_rgSD; g.ms = st->d_ms; g.rsums = st->d_rsums; g.meas = s->measCtr;
    g.ps = st->pair_s; g.pt = st->pair_t; g.pd0 = st->pair_d[0]; g.pd1 = st->pair_d[1]; g.pd2 = st->pair_d[2];
    sdispatch(s, [&]<int NC, typename real, bool FJ>() {
        k_struct_rg_majority<NC, real><<<dim3((st->nR + 255) / 256, s->R), 256, 0, s->stream>>>(a, g);
        k_struct_rg_sums<NC, real><<<dim3(std::min((st->nR + 127) / 128, 148 * 8), s->R), 128, 0, s->stream>>>(a, g, st->rg_cij > 0 ? 1 : 0);
    });
    k_extra_finalize<<<(s->R + 63) / 64, 64, 0, s->stream>>>(s->model, s->R, s->nLat, st->nR, st->rg_ci, st->rg_cj, st->rg_cij, 0, s->d_sums,
                                                              st->d_rsums, nullptr, s->d_acc, nullptr, s->d_slot);
    s->launches += 3;
    s->measCtr++;
}

static void fold_and_extras(mcg_system *s) {
    StructuredSystem *st = s->st;
    StructArgs a = struct_args(s);
    s->launches++;
    // a slab folds the sums of its own planes; the all-reduce below makes them the whole lattice's on every rank
    const double nLatOwn = st->slabWorld ? (double)s->nLatGlobal / st->slabWorld : (double)s->nLat;
    k_struct_fold<<<s->R, 32, 0, s->stream>>>(a, s->R, s->NC, s->prec, st->pair_s, st->pair_t, st->selfPairs ? 1 : 0, nLatOwn, s->d_sums,
                                                          s->nG, st->d_gmask, st->groupInSC ? 1 : 0, s->d_gacc, s->d_slot);
    slab_allreduce_sums(s);
    if (!st->selfPairs) {
        dim3 g((s->nLat + 255) / 256, s->R);
        s->launches++;
        if (s->prec == 8) k_i8_pairs<<<g, 256, 0, s->stream>>>(a, st->pair_s, st->pair_t, st->pair_d[0], st->pair_d[1], st->pair_d[2], s->d_sums);
        else sdispatch(s, [&]<int NC, typename real, bool FJ>() {
            k_struct_pairs<NC, real><<<g, 256, 0, s->stream>>>(a, st->pair_s, st->pair_t, st->pair_d[0], st->pair_d[1], st->pair_d[2], s->d_sums);
        });
    }
    if (st->ncircuit > 0 && s->NC == 3) {
        s->launches++;
        const int npar = st->p[0] * st->p[1] * st->p[2];
        if (st->nvert <= TOPO_MAXV && st->ncircuit <= 64 && st->Xd <= 65535 && s->R <= 65535) {
            int tz = 1;
            while (tz < st->Zd && tz < TOPO_THREADS) tz <<= 1;
            const int ty = TOPO_THREADS / tz;
            const int nzc = (st->Zd + tz - 1) / tz, nyc = (st->Yd + ty - 1) / ty;
            dim3 gc((unsigned)(nzc * nyc * npar), (unsigned)st->Xd, (unsigned)s->R), bc(tz, ty);
            size_t sh = (size_t)3 * st->nvert * TOPO_THREADS * (s->prec == 64 ? 8 : 4);
            if (jit_launch_topo(s, a, nzc, nyc, gc, bc)) { /* specialised kernel launched */ }
            else if (s->prec == 64) k_struct_topo_cells<double><<<gc, bc, sh, s->stream>>>(a, st->ncircuit, st->nvert, st->d_tverts, st->d_ttris, nzc, nyc, s->d_sums);
            else k_struct_topo_cells<float><<<gc, bc, sh, s->stream>>>(a, st->ncircuit, st->nvert, st->d_tverts, st->d_ttris, nzc, nyc, s->d_sums);
        } else {
            dim3 g((s->nTri + 255) / 256, s->R);
            if (s->prec == 64) k_struct_topo<double><<<g, 256, 0, s->stream>>>(a, st->ncircuit, st->d_circuits, s->d_sums);
            else k_struct_topo<float><<<g, 256, 0, s->stream>>>(a, st->ncircuit, st->d_circuits, s->d_sums);
        }
    }
    if (st->rgOn) launch_block_spin(s, a);
    MCG_CUDA(cudaGetLastError());
}

int structured_wolff_step(mcg_system *s, const WolffArgs &w, bool primed, bool needResidual, int force) {
    MCG_REQUIRE(s->prec != 8, "Wolff updates run on fp32/fp64 state: create the system with precision 32");
    MCG_REQUIRE(!s->st->slabWorld, "Wolff updates are not decomposed over slabs: run them on one GPU");
    StructArgs a = struct_args(s);
    int launches = 0;
    sdispatch(s, [&]<int NC, typename real, bool FJ>() {
        auto lg = [](int v) { int sh = 0; while ((1 << sh) < v) sh++; return (1 << sh) == v ? sh : -1; };
        StructTopo<NC, real> topo{a, lg(a.Zd), lg(a.Yd), lg(a.ncellc)};
        if (w.mode) {
            int maxL = 1;
            for (const SClassD &c : s->st->classes) maxL = std::max(maxL, c.nlink);
            launches = wolff_launch_hybrid<NC, real, FJ>(topo, w, s->stream, maxL, needResidual, force, 148 * 8);
        } else
        launches = wolff_launch_step<NC, real, FJ>(topo, w, s->stream, primed, needResidual);
    });
    return launches;
}

void structured_measure_sums(mcg_system *s) {
    for (int c = 0; c < s->C; c++) launch_pass<2>(s, c, 0, 1.0);
    fold_and_extras(s);
}

void structured_sweeps(mcg_system *s, int64_t n, double pAtt, bool fusedMeasure) {
    bool fuse = fusedMeasure && !s->st->hasSelf;
    for (int64_t it = 0; it < n; it++) {
        bool last = it == n - 1;
        for (int c = 0; c < s->C; c++) {
            StructuredSystem *st = s->st;
            const int q0c = st->colourClassStart[c], nqcc = st->colourClassStart[c + 1] - q0c;
            auto pass = [&](int from, int to) {
                if (last && fuse) launch_pass<1>(s, c, s->sweepCtr, pAtt, from, to);
                else launch_pass<0>(s, c, s->sweepCtr, pAtt, from, to);
            };
            if (st->slabWorld > 1 && st->Xd >= 6 && !getenv("MCG_SLAB_NO_OVERLAP")) {
                // slab over several ranks: the two boundary coarse planes first, then their way to the neighbours' ghosts (pack,
                // ncclSend/ncclRecv, unpack on the side stream) runs behind the interior rows.  No hazard: the interior reads
                // other colours' planes 1 .. Xd-2 and writes this colour's planes 2 .. Xd-3; the exchange reads this colour's
                // planes 1 and Xd-2 and writes its ghost planes 0 and Xd-1.
                pass(st->rowLo, st->rowLo + st->Yd);
                pass(st->rowHi - st->Yd, st->rowHi);
                MCG_CUDA(cudaEventRecord(st->slabEvA, s->stream));
                MCG_CUDA(cudaStreamWaitEvent(st->slabStream, st->slabEvA, 0));
                cudaStream_t main = s->stream;
                s->stream = st->slabStream;
                slab_exchange_classes(s, q0c, nqcc);
                s->stream = main;
                MCG_CUDA(cudaEventRecord(st->slabEvB, st->slabStream));
                pass(st->rowLo + st->Yd, st->rowHi - st->Yd);
                MCG_CUDA(cudaStreamWaitEvent(s->stream, st->slabEvB, 0));
            } else {
                pass(-1, -1);
                // the neighbours' ghosts of the classes just written, before any other colour reads them
                if (st->slabWorld) slab_exchange_classes(s, q0c, nqcc);
            }
        }
        s->sweepCtr++;
    }
    MCG_CUDA(cudaGetLastError());
    if (fusedMeasure) {
        if (fuse) fold_and_extras(s);
        else structured_measure_sums(s);
    }
}


// =============================================================================================
// Slab decomposition of ONE lattice along X over the ranks of a node (SURVEY 8e "largest lattice", optional in the reference's
// terms: it has no decomposition at all, README.md:99).  Every rank holds Lx/world planes of its own plus one ghost coarse cell
// (px planes) on either side; a colour pass updates the own rows only, then the two boundary coarse planes of the classes it
// just wrote travel to the neighbours' ghosts (ncclSend/ncclRecv on the compute stream, one message per direction; one rank:
// the periodic images are copied in place).  The Philox counters are keyed by the GLOBAL reference site ids, so the slabs
// together perform - bit for bit - the trajectory the undivided lattice performs on one GPU (tests/test_gpu_slab.py).  The raw
// measurement sums are all-reduced before the non-linear fold, so every rank accumulates the observables of the whole lattice.
// =============================================================================================
template <typename real>
static __global__ void __launch_bounds__(256) k_slab_pack(const real *__restrict__ spin, real *__restrict__ toLeft, real *__restrict__ toRight, int R, int NC,
                                                          size_t N, int q0, int nqc, int Xd, int plane) {
    const size_t total = (size_t)R * NC * nqc * plane;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % plane);
        size_t t = e / plane;
        const int j = (int)(t % nqc); t /= nqc;
        const int c = (int)(t % NC), r = (int)(t / NC);
        const real *base = spin + ((size_t)r * NC + c) * N + (size_t)(q0 + j) * Xd * plane;
        toLeft[e] = base[(size_t)plane + i];                  // first own coarse plane
        toRight[e] = base[(size_t)(Xd - 2) * plane + i];      // last own coarse plane
    }
}
// fromRight: the right neighbour's first own plane -> ghost X = Xd-1;  fromLeft: the left neighbour's last own plane -> ghost X = 0.
// SELF (one rank): both neighbours are this rank, the periodic images are read straight from the own planes.
template <typename real, bool SELF>
static __global__ void __launch_bounds__(256) k_slab_unpack(real *__restrict__ spin, const real *__restrict__ fromRight, const real *__restrict__ fromLeft, int R,
                                                            int NC, size_t N, int q0, int nqc, int Xd, int plane) {
    const size_t total = (size_t)R * NC * nqc * plane;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % plane);
        size_t t = e / plane;
        const int j = (int)(t % nqc); t /= nqc;
        const int c = (int)(t % NC), r = (int)(t / NC);
        real *base = spin + ((size_t)r * NC + c) * N + (size_t)(q0 + j) * Xd * plane;
        base[(size_t)(Xd - 1) * plane + i] = SELF ? base[(size_t)plane + i] : fromRight[e];
        base[i] = SELF ? base[(size_t)(Xd - 2) * plane + i] : fromLeft[e];
    }
}

// ghost planes of the classes [q0, q0 + nqc) - one colour, or all classes - brought up to date; stream-ordered, no host sync
static void slab_exchange_classes(mcg_system *s, int q0, int nqc) {
    StructuredSystem *st = s->st;
    if (!st->slabWorld || nqc == 0) return;
    const int plane = st->Yd * st->Zd;
    const size_t total = (size_t)s->R * s->NC * nqc * plane;
    const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    const size_t rs = s->real_size();
    s->launches++;
    if (st->slabWorld == 1) {
        if (s->prec == 32) k_slab_unpack<float, true><<<grid, 256, 0, s->stream>>>((float *)s->d_spin, nullptr, nullptr, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
        else k_slab_unpack<double, true><<<grid, 256, 0, s->stream>>>((double *)s->d_spin, nullptr, nullptr, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
        MCG_CUDA(cudaGetLastError());
        return;
    }
    MCG_REQUIRE(total <= st->slabBufElems, "slab exchange buffer too small");
    char *buf = (char *)st->d_slabBuf;
    void *toLeft = buf, *toRight = buf + st->slabBufElems * rs, *fromRight = buf + 2 * st->slabBufElems * rs, *fromLeft = buf + 3 * st->slabBufElems * rs;
    if (s->prec == 32) k_slab_pack<float><<<grid, 256, 0, s->stream>>>((const float *)s->d_spin, (float *)toLeft, (float *)toRight, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
    else k_slab_pack<double><<<grid, 256, 0, s->stream>>>((const double *)s->d_spin, (double *)toLeft, (double *)toRight, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
    NcclApi &api = nccl_api();
    ncclComm_t comm = (ncclComm_t)st->slabComm;
    const int W = st->slabWorld, left = (st->slabRank + W - 1) % W, right = (st->slabRank + 1) % W;
    // a pair of ranks matches its messages in order: with two ranks the neighbour's "to its left" arrives first, which is this
    // rank's right ghost - hence receive-from-right before receive-from-left
    MCG_NCCL(api.groupStart());
    MCG_NCCL(api.send(toLeft, total * rs, ncclChar, left, comm, s->stream));
    MCG_NCCL(api.send(toRight, total * rs, ncclChar, right, comm, s->stream));
    MCG_NCCL(api.recv(fromRight, total * rs, ncclChar, right, comm, s->stream));
    MCG_NCCL(api.recv(fromLeft, total * rs, ncclChar, left, comm, s->stream));
    MCG_NCCL(api.groupEnd());
    s->launches++;
    if (s->prec == 32) k_slab_unpack<float, false><<<grid, 256, 0, s->stream>>>((float *)s->d_spin, (const float *)fromRight, (const float *)fromLeft, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
    else k_slab_unpack<double, false><<<grid, 256, 0, s->stream>>>((double *)s->d_spin, (const double *)fromRight, (const double *)fromLeft, s->R, s->NC, (size_t)s->N, q0, nqc, st->Xd, plane);
    MCG_CUDA(cudaGetLastError());
}

void structured_slab_exchange_all(mcg_system *s) { slab_exchange_classes(s, 0, s->st->nclass); }
bool structured_is_slab(const mcg_system *s) { return s->st && s->st->slabWorld > 0; }
void structured_slab_info(const mcg_system *s, int32_t *info) {
    const StructuredSystem *st = s->st;
    const int px = st->p[0];
    info[0] = st->slabRank; info[1] = st->slabWorld; info[2] = st->xoff + px;   // first own global x
    info[3] = (st->Xd - 2) * px;                                                 // own planes
    info[4] = px;                                                                // ghost planes on either side
    info[5] = st->LxGlobal;
}

// raw per-sweep sums of the own rows -> sums of the whole lattice on every rank
static void slab_allreduce_sums(mcg_system *s) {
    StructuredSystem *st = s->st;
    if (st->slabWorld <= 1) return;
    MCG_NCCL(nccl_api().allReduce(s->d_sums, s->d_sums, (size_t)s->R * NSUM, ncclDouble, ncclSum, (ncclComm_t)st->slabComm, s->stream));
}

// host-only: how the whole lattice is cut (no device needed): {rank, world, first own x, own planes, ghost planes per side, L[0]}
static SlabPlan slab_plan(const mcg_lattice_desc *global, int precision, int device, int rank, int world);
void structured_slab_plan(const mcg_lattice_desc *global, int precision, int rank, int world, int32_t *info) {
    const SlabPlan p = slab_plan(global, precision, 0, rank, world);
    const int own = p.LxGlobal / world;
    info[0] = rank; info[1] = world; info[2] = rank * own; info[3] = own; info[4] = p.period[0]; info[5] = p.LxGlobal;
}

static SlabPlan slab_plan(const mcg_lattice_desc *global, int precision, int device, int rank, int world) {
    MCG_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad slab rank/world");
    MCG_REQUIRE(precision == 32 || precision == 64, "slab decomposition: fp32 or fp64 state");
    MCG_REQUIRE(global->L[0] > 1 && global->L[1] > 1 && global->L[2] > 1, "slab decomposition is for three-dimensional supercells (it cuts the first axis)");
    MCG_REQUIRE(global->ncircuit == 0 && !global->block_spin, "slab decomposition: topological charge and block-spin statistics are not decomposed - run them on one GPU");
    MCG_REQUIRE(global->pair_s == global->pair_t && global->pair_d[0] == 0 && global->pair_d[1] == 0 && global->pair_d[2] == 0,
                "slab decomposition: the correlated pair must be a site with itself (the default)");
    // period of the whole lattice: the slabs must be coloured exactly like it
    SlabPlan plan;
    {
        mcg_system tmp;
        tmp.prec = precision; tmp.R = 1;
        tmp.device = device;          // its destructor selects its device
        std::vector<SLinkD> l; std::vector<double> j;
        structured_build_host(&tmp, global, l, j);
        for (int k = 0; k < 3; k++) plan.period[k] = tmp.st->p[k];
        MCG_REQUIRE(tmp.st->V > 1, "slab decomposition needs vector items (innermost coarse dimension a multiple of 4, fp64: 2)");
        for (int c = 0; c < tmp.C; c++) MCG_REQUIRE(tmp.st->fastOK[c], "slab decomposition needs bonds that reach at most one colouring period along every axis");
    }
    const int px = plan.period[0], Lx = global->L[0];
    MCG_REQUIRE(Lx % world == 0 && (Lx / world) % px == 0 && Lx / world >= 2 * px, "the first supercell dimension must split into equal slabs of whole colouring periods, two or more per rank");
    plan.rank = rank; plan.world = world; plan.LxGlobal = Lx;
    return plan;
}

void structured_create_slab(mcg_system *s, const mcg_lattice_desc *global, int rank, int world, const char *commId) {
    const SlabPlan plan = slab_plan(global, s->prec, s->device, rank, world);
    const int px = plan.period[0], Lx = global->L[0];
    mcg_lattice_desc local = *global;
    local.L[0] = Lx / world + 2 * px;
    local.ngroup = 0;      // orbital-group statistics are products of sums: not decomposed
    structured_create_impl(s, &local, &plan);
    StructuredSystem *st = s->st;
    MCG_REQUIRE(!st->hasSelf, "slab decomposition: bonds of a site with its own periodic image are not supported");
    // observables are those of the whole lattice
    s->nLatGlobal = s->nLat / (st->Xd * px) * Lx;
    s->NGlobal = s->N / (st->Xd * px) * Lx;
    if (world > 1) {
        MCG_REQUIRE(commId, "slab decomposition over several ranks needs the communicator id (mcg_comm_unique_id on rank 0)");
        NcclApi &api = nccl_api();
        if (!api.ok) throw Error(MCG_ERR_NCCL, api.why);
        ncclUniqueId u;
        memcpy(&u, commId, sizeof u);
        ncclComm_t comm = nullptr;
        MCG_CUDA(cudaSetDevice(s->device));
        MCG_NCCL(api.commInitRank(&comm, world, u, rank));
        st->slabComm = comm;
        int maxq = 0;
        for (int c = 0; c < s->C; c++) maxq = std::max(maxq, st->colourClassStart[c + 1] - st->colourClassStart[c]);
        maxq = std::max(maxq, st->nclass);      // the initial exchange moves every class at once
        st->slabBufElems = (size_t)s->R * s->NC * maxq * st->Yd * st->Zd;
        st->d_slabBuf = pool_alloc(4 * st->slabBufElems * s->real_size());
        MCG_CUDA(cudaStreamCreateWithFlags(&st->slabStream, cudaStreamNonBlocking));
        MCG_CUDA(cudaEventCreateWithFlags(&st->slabEvA, cudaEventDisableTiming));
        MCG_CUDA(cudaEventCreateWithFlags(&st->slabEvB, cudaEventDisableTiming));
    }
}

}  // namespace mcg
