// The Metropolis colour-pass kernel of the structured path, written once and compiled twice:
//
//  * OFFLINE (nvcc, this header included by structured.cu): lattice dimensions and the colour's
//    link table are runtime values; the table travels as a __grid_constant__ kernel parameter.
//  * JIT (NVRTC at system creation, -DMCG_JIT + a generated prologue): the lattice IS the program.
//    Supercell dims, class coordinates, per-link cell offsets, wrap masks, Z shifts and the exchange
//    tensors are literals; the link loop is unrolled at compile time (no table loads, no index
//    multiplies by runtime strides, no branches on the link kind, FFMA with immediate J).
//    profiles/r01: the offline kernel spent ~150 of its ~315 instructions per attempt on exactly
//    that bookkeeping and was issue-bound at 1/3 of the HBM roofline.
//
// Must stay free of host/std includes (NVRTC).  Arithmetic is identical in both builds.
#pragma once
#include "devmath.cuh"

namespace mcg {

constexpr int MAXLINK = 32;

struct SLinkD {
    int qn;          // neighbour class
    int cX, cY;      // coarse offsets, already reduced to [0,Xd) / [0,Yd)
    int cZ;          // coarse Z offset in (-Zd/2, Zd/2]
    int low;         // neighbour class has a lower colour (its spins are final in this sweep)
    int self;        // link to the site itself (supercell dimension 1)
    int fwd;         // this endpoint is the bond template's source: it activates the bond in the Wolff pass
    int o2, dx, dy, dz;   // neighbour orbital and cell offset reduced to [0,L): its reference id without decoding
};
struct SClassD {
    int a, b, c, o, colour, nlink, lowmode, pad;   // pad: number of leading links into lower colours (links are sorted low-first)
    double S, D[3];
};

struct StructArgs {
    int Xd, Yd, Zd, Zc, ncellc, nclass, nrows, N;
    int px, py, pz, norb, Lx, Ly, Lz;
    int psx, psy, psz;       // log2 of the colouring period per axis, or -1 if it is not a power of two
    const SClassD *classes;
    const SLinkD *links;
    const void *J;
    const int *classOf;
    void *spin;
    const double *beta, *field;
    unsigned long long *cnt;
    double *classSums;
    RngKey key;
    uint32_t replica0;
    // slab decomposition along X (slab.cu): only the rows [rowLo, rowHi) are updated and measured - the coarse planes outside are
    // ghosts, copies of the neighbouring ranks' boundary planes - and the reference site ids behind the Philox counters are those
    // of the GLOBAL lattice: local x + xoff.  Whole lattice on one GPU: rowLo = 0, rowHi = nrows, xoff = 0.
    int rowLo, rowHi, xoff;
};

template <typename real, int V> struct Vec;
template <> struct Vec<float, 4> { typedef float4 type; };
template <> struct Vec<float, 1> { typedef float type; };
template <> struct Vec<double, 2> { typedef double2 type; };
template <> struct Vec<double, 1> { typedef double type; };

template <typename real, int V> __device__ __forceinline__ void vload(const real *__restrict__ p, real (&o)[V]) {
    typedef typename Vec<real, V>::type VT;
    VT v = *reinterpret_cast<const VT *>(p);
    const real *e = reinterpret_cast<const real *>(&v);
#pragma unroll
    for (int i = 0; i < V; i++) o[i] = e[i];
}
template <typename real, int V> __device__ __forceinline__ void vstore(real *__restrict__ p, const real (&o)[V]) {
    typedef typename Vec<real, V>::type VT;
    VT v;
    real *e = reinterpret_cast<real *>(&v);
#pragma unroll
    for (int i = 0; i < V; i++) e[i] = o[i];
    *reinterpret_cast<VT *>(p) = v;
}

// V consecutive cells of a neighbour row, shifted by cZ cells with periodic wrap
template <typename real, int V>
__device__ __forceinline__ void load_shifted(const real *__restrict__ row, int Z0, int cZ, int Zd, real (&o)[V]) {
    if (cZ == 0) {
        vload<real, V>(row + Z0, o);
    } else if (V > 1 && cZ == -1) {
        real t[V];
        vload<real, V>(row + Z0, t);
        int zl = Z0 == 0 ? Zd - 1 : Z0 - 1;
        o[0] = row[zl];
#pragma unroll
        for (int i = 1; i < V; i++) o[i] = t[i - 1];
    } else if (V > 1 && cZ == 1) {
        real t[V];
        vload<real, V>(row + Z0, t);
        int zr = Z0 + V >= Zd ? 0 : Z0 + V;
#pragma unroll
        for (int i = 0; i < V - 1; i++) o[i] = t[i + 1];
        o[V - 1] = row[zr];
    } else {
#pragma unroll
        for (int i = 0; i < V; i++) {
            int z = Z0 + i + cZ;
            if (z < 0) z += Zd;
            if (z >= Zd) z -= Zd;
            o[i] = row[z];
        }
    }
}

// the NC component rows of one neighbour at once: ONE decision on the shift for all components (with a runtime link
// table the per-component version re-branched three times and kept three copies of the address arithmetic alive)
template <typename real, int V, int NC>
__device__ __forceinline__ void load_shifted_nc(const real *__restrict__ sp, size_t N, int nb, int Z0, int cZ, int Zd, real (&o)[3][V]) {
    // rows are addressed as (sp + c*N) + nb: the component planes are loop invariants, the link adds a 32-bit offset
    if (cZ == 0) {
#pragma unroll
        for (int c = 0; c < NC; c++) vload<real, V>(sp + (size_t)c * N + nb + Z0, o[c]);
    } else if (V > 1 && cZ == -1) {
        const int zl = Z0 == 0 ? Zd - 1 : Z0 - 1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const real *row = sp + (size_t)c * N + nb;
            real t[V];
            vload<real, V>(row + Z0, t);
            o[c][0] = row[zl];
#pragma unroll
            for (int i = 1; i < V; i++) o[c][i] = t[i - 1];
        }
    } else if (V > 1 && cZ == 1) {
        const int zr = Z0 + V >= Zd ? 0 : Z0 + V;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const real *row = sp + (size_t)c * N + nb;
            real t[V];
            vload<real, V>(row + Z0, t);
#pragma unroll
            for (int i = 0; i < V - 1; i++) o[c][i] = t[i + 1];
            o[c][V - 1] = row[zr];
        }
    } else {
        int z[V];
#pragma unroll
        for (int i = 0; i < V; i++) {
            z[i] = Z0 + i + cZ;
            if (z[i] < 0) z[i] += Zd;
            if (z[i] >= Zd) z[i] -= Zd;
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const real *row = sp + (size_t)c * N + nb;
#pragma unroll
            for (int i = 0; i < V; i++) o[c][i] = row[z[i]];
        }
    }
}

// ---- link / class tables of one colour pass ----
constexpr int PT_MAXC = 8, PT_MAXL = 32;
template <typename real> struct PLink {
    int delta;                  // (qn-q)*ncellc + (cX*Yd + cY)*Zd with signed cX,cY in {-1,0,1}
    int mxp, mxm, myp, mym;     // 0/1: which per-row wrap correction applies
    int cZ, low, pad;
    real J[9];
};
template <typename real> struct PassTable {
    int nl, nqc, uniformJ, pad1;   // uniformJ: every link of every class of the pass carries the same diagonal exchange
    int ca[PT_MAXC], cb[PT_MAXC], cc[PT_MAXC], co[PT_MAXC], lowmode[PT_MAXC], nlow[PT_MAXC];
    // links are sorted (lower colours first; within each half by Z-shift kind 0, -1, +1, other): gend[j][4*half + kind] is
    // the end index of that run, so every run is a loop whose shift is a compile-time constant
    int gend[PT_MAXC][8];
    real S[PT_MAXC], D[PT_MAXC][3];
    PLink<real> L[PT_MAXC][PT_MAXL];
};

// runtime views (offline build)
// CZ: the link's Z shift when the run it belongs to fixes it (0, -1, +1), 2 = read it from the table
template <typename real, int CZ = 2> struct RtLink {
    const PLink<real> &L;
    int k;
    __device__ __forceinline__ int idx() const { return k; }
    __device__ __forceinline__ int delta() const { return L.delta; }
    __device__ __forceinline__ int mxp() const { return L.mxp; }
    __device__ __forceinline__ int mxm() const { return L.mxm; }
    __device__ __forceinline__ int myp() const { return L.myp; }
    __device__ __forceinline__ int mym() const { return L.mym; }
    __device__ __forceinline__ int cZ() const { return CZ == 2 ? L.cZ : CZ; }
    __device__ __forceinline__ int low() const { return L.low; }
    __device__ __forceinline__ real J(int e) const { return L.J[e]; }
    // the neighbour's V cells of NC components; base = first component's row of the neighbour class
    template <typename R, int V, int NC> __device__ __forceinline__ void fetch(const R *__restrict__ sp, int nb, size_t N, int Z0, int Zd, R (&t)[3][V]) const {
        load_shifted_nc<R, V, NC>(sp, N, nb, Z0, cZ(), Zd, t);
    }
};
template <typename real> struct RtClass {
    const PassTable<real> &T;
    int j;
    __device__ __forceinline__ int nl() const { return T.nl; }
    __device__ __forceinline__ int ca() const { return T.ca[j]; }
    __device__ __forceinline__ int cb() const { return T.cb[j]; }
    __device__ __forceinline__ int cc() const { return T.cc[j]; }
    __device__ __forceinline__ int co() const { return T.co[j]; }
    __device__ __forceinline__ int lowmode() const { return T.lowmode[j]; }
    __device__ __forceinline__ int nlow() const { return T.nlow[j]; }
    __device__ __forceinline__ real S() const { return T.S[j]; }
    __device__ __forceinline__ real D(int e) const { return T.D[j][e]; }
    __device__ __forceinline__ bool uniformJ() const { return T.uniformJ != 0; }
    // f on every link; snap once after the leading nlow links (the lower-colour neighbours).  Two loops instead of a
    // test inside one: the snapshot is a dozen register moves that were predicated into every iteration
    // the item's own spins: read up front (the asynchronous pipeline fetches them behind the last link instead)
    template <typename R, int V, int NC> __device__ __forceinline__ void own_early(const R *__restrict__ own, size_t N, R (&sv)[3][V]) const {
#pragma unroll
        for (int c = 0; c < NC; c++) vload<R, V>(own + (size_t)c * N, sv[c]);
    }
    template <typename R, int V, int NC> __device__ __forceinline__ void own_late(R (&)[3][V]) const {}
    template <int CZ, typename F> __device__ __forceinline__ void run(int &k, int kend, F &&f) const {
        // two links per trip: the second link's loads are issued before the first one's multiply-adds retire
#pragma unroll 2
        for (; k < kend; k++) f(RtLink<real, CZ>{T.L[j][k], k});
    }
    template <typename X, typename P, typename F, typename G> __device__ __forceinline__ void for_links(const X &, P &&, F &&f, G &&snap) const {
        int k = 0;
#pragma unroll 1
        for (int part = 0; part < 2; part++) {
            const int *ge = T.gend[j] + 4 * part;
            run<0>(k, ge[0], f);
            run<-1>(k, ge[1], f);
            run<1>(k, ge[2], f);
            run<2>(k, ge[3], f);
            if (part == 0) snap();
        }
    }
};

// links whose coefficients arrive as splat pairs: the field sums of a full tensor run as packed FFMA2
template <typename LK> struct LinkPacked { static constexpr bool value = false; };

// what a link loop needs to know about the item being updated (the row-dependent part is the `pre` callable)
template <typename real> struct LinkCtx {
    const real *sp;   // replica's first component plane
    size_t N;
    int Z0, Zd;
};

#ifndef MCG_JIT
// ---- asynchronous link pipeline (offline full-tensor fp32 kernel: dipole stencils, 32 links of 9 coefficients) ----
// The pass is latency-bound (ncu: long-scoreboard stalls 7.6 of 11.7 cycles per issue at 24 warps per SM): registers allow
// one or two neighbours in flight per thread.  Here every thread keeps ASYNC_D links in flight with cp.async into ITS OWN
// shared-memory slots - nobody else reads them, so there is no block barrier, only cp.async.wait_group - and the Z shift is
// applied while copying (four 4-byte copies instead of one 16-byte copy), so the consumer is the same three LDS.128 for
// every link.
constexpr int ASYNC_D = 4;
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp_async16q(unsigned dst, unsigned long long src) {   // src: a global address
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// keeps a loop invariant in its register: without it the compiler re-derives thread ids, the shared window base and the
// replica's 64-bit plane offset for every link (ptxas prefers rematerialising to holding them under the 80-register cap)
__device__ __forceinline__ void pin(unsigned &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(unsigned long long &v) { asm volatile("" : "+l"(v)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

// the class's link table as the pipeline reads it from shared memory: four 16-byte loads per link instead of 15 scalar ones
struct AsTab {
    float4 j[5];   // the nine coefficients as splat pairs (J0,J0,J1,J1)(J2,J2,J3,J3)...(J8,J8,0,0): operands of FFMA2 as loaded
    int4 a;        // delta, cX sign, cY sign, cZ
};
struct AsAddr {    // just enough of a link for nb_of
    int4 a;
    __device__ __forceinline__ int delta() const { return a.x; }
    __device__ __forceinline__ int mxp() const { return a.y > 0; }
    __device__ __forceinline__ int mxm() const { return a.y < 0; }
    __device__ __forceinline__ int myp() const { return a.z > 0; }
    __device__ __forceinline__ int mym() const { return a.z < 0; }
};
struct AsLink {
    float4 j0, j1, j2, j3, j4;
    int k;
    unsigned slot;        // shared-memory address of this thread's slot in the stage that holds link k; components ASYNC_D * 4096 B apart
    __device__ __forceinline__ int idx() const { return k; }
    // the consumer never needs the neighbour's position (fetch reads the stage): nb_of on this type folds to a dead value
    __device__ __forceinline__ int delta() const { return 0; }
    __device__ __forceinline__ int mxp() const { return 0; }
    __device__ __forceinline__ int mxm() const { return 0; }
    __device__ __forceinline__ int myp() const { return 0; }
    __device__ __forceinline__ int mym() const { return 0; }
    __device__ __forceinline__ float J(int e) const {
        return e == 0 ? j0.x : e == 1 ? j0.z : e == 2 ? j1.x : e == 3 ? j1.z : e == 4 ? j2.x : e == 5 ? j2.z : e == 6 ? j3.x : e == 7 ? j3.z : j4.x;
    }
    __device__ __forceinline__ F2 J2(int e) const {
        return e == 0 ? F2{j0.x, j0.y} : e == 1 ? F2{j0.z, j0.w} : e == 2 ? F2{j1.x, j1.y} : e == 3 ? F2{j1.z, j1.w} : e == 4 ? F2{j2.x, j2.y}
             : e == 5 ? F2{j2.z, j2.w} : e == 6 ? F2{j3.x, j3.y} : e == 7 ? F2{j3.z, j3.w} : F2{j4.x, j4.y};
    }
    template <typename R, int V, int NC> __device__ __forceinline__ void fetch(const R *__restrict__, int, size_t, int, int, R (&t)[3][V]) const {
        static_assert(V == 4 && sizeof(R) == 4, "async pipeline: fp32 items of four sites");
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float4 v = lds128(slot + c * (ASYNC_D * 4096));
            t[c][0] = v.x; t[c][1] = v.y; t[c][2] = v.z; t[c][3] = v.w;
        }
    }
};
template <> struct LinkPacked<AsLink> { static constexpr bool value = true; };
struct AsClass : RtClass<float> {
    float4 *stages;      // [3 components][ASYNC_D][256 threads]
    const AsTab *tab;    // [nlinks + 1] in shared memory; the last entry is the item itself (offset 0, no shift)
    static constexpr unsigned STAGE_B = 256 * 16, COMP_B = ASYNC_D * STAGE_B;
    // called by every thread of the block before the pass: src = this class's records in global memory (built on the host)
    __device__ __forceinline__ void stage_table(AsTab *dst, const AsTab *__restrict__ src) const {
        const int n16 = (T.gend[j][7] + 1) * (int)(sizeof(AsTab) / 16);
        for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < n16; i += blockDim.x * blockDim.y)
            reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(src)[i];
        __syncthreads();
    }
    __device__ __forceinline__ unsigned slot0() const { return smem_u32(stages) + (threadIdx.y * blockDim.x + threadIdx.x) * 16; }
    template <typename R, int V, int NC> __device__ __forceinline__ void own_early(const R *__restrict__, size_t, R (&)[3][V]) const {}
    template <typename R, int V, int NC> __device__ __forceinline__ void own_late(R (&sv)[3][V]) const {
        cp_async_wait<0>();
        const int n = T.gend[j][7];
        const unsigned p = slot0() + (unsigned)(n & (ASYNC_D - 1)) * STAGE_B;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float4 v = lds128(p + c * COMP_B);
            sv[c][0] = v.x; sv[c][1] = v.y; sv[c][2] = v.z; sv[c][3] = v.w;
        }
    }
    template <typename P, typename F, typename G> __device__ __forceinline__ void for_links(const LinkCtx<float> &x, P &&pre, F &&f, G &&snap) const {
        const int n = T.gend[j][7], nlo = T.nlow[j];
        unsigned st0 = slot0();
        unsigned long long ib = (unsigned long long)__cvta_generic_to_global(x.sp + x.Z0);
        pin(st0);
        pin(ib);
        const unsigned long long planeB = (unsigned long long)x.N * 4;
        auto issue = [&](int k) {
            if (k <= n) {
                const int4 a = tab[k].a;
                const unsigned long long base = ib + (long long)pre(AsAddr{a}) * 4;
                const unsigned dst = st0 + (unsigned)(k & (ASYNC_D - 1)) * STAGE_B;
                const int cz = a.w;
                if (cz == 0) {
                    cp_async16q(dst, base);
                    cp_async16q(dst + COMP_B, base + planeB);
                    cp_async16q(dst + 2 * COMP_B, base + 2 * planeB);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        int dz = i + cz;
                        if (x.Z0 + dz < 0) dz += x.Zd;
                        if (x.Z0 + dz >= x.Zd) dz -= x.Zd;
#pragma unroll
                        for (int c = 0; c < 3; c++) cp_async4(dst + c * COMP_B + 4 * i, reinterpret_cast<const float *>(base + c * planeB) + dz);
                    }
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < ASYNC_D - 1; k++) issue(k);
        if (nlo == 0) snap();
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            issue(k + ASYNC_D - 1);
            cp_async_wait<ASYNC_D - 1>();
            f(AsLink{tab[k].j[0], tab[k].j[1], tab[k].j[2], tab[k].j[3], tab[k].j[4], k, st0 + (unsigned)(k & (ASYNC_D - 1)) * STAGE_B});
            if (k == nlo - 1) snap();
        }
    }
};
#endif

// cp.async.bulk.prefetch.L2: one instruction asks the L2 for `bytes` (multiple of 16) starting at a 16-byte aligned address
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(bytes) : "memory");
}

// dims: runtime fields offline, literals under JIT (the generated prologue defines JIT_Xd ...)
#ifdef MCG_JIT
#ifndef JIT_PF
#define JIT_PF 0
#endif
#ifndef JIT_ACC
#define JIT_ACC 0      // per-thread fused M,E accumulators: 0 fp32 (a thread sums at most a few dozen sites), 1 fp64 - see pass_body
#endif
#define MCG_DIM(a, f) (JIT_##f)
#else
#define MCG_DIM(a, f) ((a).f)
#endif

template <bool NARROW, typename real> struct AccT { typedef double type; };
template <typename real> struct AccT<true, real> { typedef real type; };
template <int I> struct IC { static constexpr int value = I; };
template <int I, int N, typename F> __device__ __forceinline__ void ct_for(F &&f) {
    if constexpr (I < N) {
        f(IC<I>{});
        ct_for<I + 1, N>(f);
    }
}

// position of the neighbour's row.  Runtime tables: the four wrap corrections vanish on all but the boundary rows, so
// they sit behind a test (4 multiply-adds and 4 table loads per link otherwise); literals (JIT) fold either way.
template <typename LK>
__device__ __forceinline__ int nb_of(const LK &L, int rowBase, bool edgeRow, int wxp, int wxm, int wyp, int wym) {
#ifdef MCG_JIT
    return rowBase + L.delta() + L.mxp() * wxp + L.mxm() * wxm + L.myp() * wyp + L.mym() * wym;
#else
    int nb = rowBase + L.delta();
    if (edgeRow) nb += L.mxp() * wxp + L.mxm() * wxm + L.myp() * wyp + L.mym() * wym;
    return nb;
#endif
}

// One colour pass over the rows [rb*rowsPerBlock, ...) of class q = q0 + j for replica r.
// MODE 0: update only   1: update + fused measurement (classSums += M, E contributions)
// PARTIAL: sweeps in which a site attempts only with probability pAtt (ninterval < N); the common
// full sweep drops the fourth uniform, its compare and the attempt counter from the per-site code.
template <int NC, typename real, bool FULLJ, int MODE, int V, bool PARTIAL, typename CLS>
__device__ __forceinline__ void pass_body(const StructArgs &a, const CLS cls, int q, int r, int rb, int rowsPerBlock, uint64_t sweep,
                                          real pAtt, double *red, float4 *hls = nullptr) {
    // hls (offline full-tensor kernel only): per-thread shared-memory parking for the lower-colour field snapshot, so its
    // 12 registers are free while the remaining links stream in
#ifndef MCG_JIT
    constexpr bool PARK = FULLJ && NC == 3 && MODE == 1 && V == 4 && sizeof(real) == 4;
#else
    constexpr bool PARK = false;
#endif
    const int Xd = MCG_DIM(a, Xd), Yd = MCG_DIM(a, Yd), Zd = MCG_DIM(a, Zd), Zc = MCG_DIM(a, Zc), N = MCG_DIM(a, N);
    const int px = MCG_DIM(a, px), py = MCG_DIM(a, py), pz = MCG_DIM(a, pz), norb = MCG_DIM(a, norb);
    const int Ly = MCG_DIM(a, Ly), Lz = MCG_DIM(a, Lz), nclass = MCG_DIM(a, nclass);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lowmode = cls.lowmode();
    const real S = cls.S();
    const real D0 = cls.D(0), D1 = cls.D(1), D2 = cls.D(2);
    const bool hasD = D0 != real(0) || D1 != real(0) || D2 != real(0);   // a literal under JIT: the D terms fold away
    // fp32 state: |s| is pinned back to S every 8th sweep (rounding drifts it by ~1e-7 per accepted move)
    const bool renorm = sizeof(real) == 4 && (sweep & 7) == 0;
    const real beta = (real)a.beta[r], hf = (real)(a.beta[r] * a.field[r]);
    real *sp = (real *)a.spin + (size_t)r * NC * N;
    const int idStrideZ = pz * norb;
    const int planeY = Yd * Zd, planeX = planeY * Xd;
    // fused M and E.  A thread sums its own sites - a few dozen - in fp32 and everything from the warp reduction on is fp64; the
    // rounding of those short partial sums averages out over the ~1e6 threads of a pass (fused E == recomputed E to ~1e-8 at
    // 256^3).  JIT_ACC=1 makes the per-thread sums fp64 from the item level (5e-9 ... 5e-11): -1.6 % in bursts but -9.4 % sustained
    // on the sc 256^3 fp32 pass - under the board power cap the eight fp64-pipe instructions per item are paid in clock
    // (profiles/r02b_accumulators.txt) - so it is opt-in (MCG_JIT_ACC=1).
#ifdef MCG_JIT
    typedef typename AccT<JIT_ACC == 0, real>::type acc_t;
#else
    typedef double acc_t;
#endif
    acc_t accM[3] = {0, 0, 0}, accE = 0;
    auto acc_add = [&](double m0, double m1, double m2, double e) {
        accM[0] += (acc_t)m0; accM[1] += (acc_t)m1; accM[2] += (acc_t)m2; accE += (acc_t)e;
    };
    int natt = 0, nacc = 0;

    const int rowLo = a.rowLo, rowHi = a.rowHi, xoff = MCG_DIM(a, xoff);   // the row range is a launch argument in both builds
    const int rowEnd = min(rowHi, rowLo + (rb + 1) * rowsPerBlock);
    for (int row = rowLo + rb * rowsPerBlock + threadIdx.y; row < rowEnd; row += blockDim.y) {
        const int X = row / Yd, Y = row - X * Yd;
        const int rowBase = ((q * Xd + X) * Yd + Y) * Zd;
        const int wxp = X == Xd - 1 ? -planeX : 0, wxm = X == 0 ? planeX : 0;
        const int wyp = Y == Yd - 1 ? -planeY : 0, wym = Y == 0 ? planeY : 0;
        const bool edgeRow = (wxp | wxm | wyp | wym) != 0;
#ifdef MCG_JIT
        if constexpr (JIT_PF != 0) {
            // Bulk L2 prefetch of everything the NEXT row of this thread row will read (own row and every neighbour row, whole
            // rows of Zd cells): one cp.async.bulk.prefetch.L2 per row and component, issued by one thread, a full item's worth
            // of arithmetic (microseconds) ahead of the loads.  The passes that run two blocks per SM (fp64 state, long link
            // lists) are bound by exposed DRAM latency, not by issue slots or bandwidth (profiles/r02a: fp64 pass 49 % issue
            // utilisation, 51 % DRAM utilisation, long-scoreboard stalls on the first use of a neighbour value).
            const int rowN = row + (int)blockDim.y;
            if (threadIdx.x == 0 && rowN < rowEnd) {
                const int Xn = rowN / Yd, Yn = rowN - Xn * Yd;
                const int rowBaseN = ((q * Xd + Xn) * Yd + Yn) * Zd;
                const int wxpN = Xn == Xd - 1 ? -planeX : 0, wxmN = Xn == 0 ? planeX : 0;
                const int wypN = Yn == Yd - 1 ? -planeY : 0, wymN = Yn == 0 ? planeY : 0;
                constexpr unsigned rowBytes = (unsigned)(JIT_Zd * sizeof(real));
                static_assert(rowBytes % 16 == 0, "bulk prefetch: rows are multiples of 16 bytes");
#pragma unroll
                for (int c = 0; c < NC; c++) bulk_prefetch_l2(sp + (size_t)c * N + rowBaseN, rowBytes);
                auto nbOfN = [&](auto L) { return nb_of(L, rowBaseN, true, wxpN, wxmN, wypN, wymN); };
                // 1: every neighbour row (doubles the L2 request traffic: each is wanted by up to z rows of this colour);
                // 2: the own row only - the one certain DRAM miss;  3: own row + the first link's row: one direction maps the
                // rows of this colour one-to-one onto rows of the other, so every neighbour row is asked for about once
                if constexpr (JIT_PF == 1 || JIT_PF == 3)
                cls.for_links(LinkCtx<real>{sp, (size_t)N, 0, Zd}, nbOfN, [&](auto L) {
                    if (JIT_PF == 3 && L.idx() != 0) return;
                    const int nb = nbOfN(L);
#pragma unroll
                    for (int c = 0; c < NC; c++) bulk_prefetch_l2(sp + (size_t)c * N + nb, rowBytes);
                }, [&]() {});
            }
        }
#endif
        const int xy = ((X * px + cls.ca() + xoff) * Ly + (Y * py + cls.cb())) * Lz;
        for (int zc = threadIdx.x; zc < Zc; zc += blockDim.x) {
            const int Z0 = zc * V;
#ifndef MCG_NO_F32X2
            if constexpr (NC >= 2 && !FULLJ && sizeof(real) == 4 && V == 4) {   // full tensors (DMI): measured 5 % slower packed (C4)
                // ---- XY / Heisenberg fp32: the four sites of the item as two packed pairs (sites 0,1 and 2,3) ----
                // Same formulas as the scalar branch below, with the plane normal taken as m = -n: the reflection
                // s' = s - 2 (s.m) m and its energy are even in the normal, so no sign has to be applied to the
                // sine/cosine pair and the move is bit-identical.
                F2 s2[3][2], H2[3][2], Hl2[3][2];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int p = 0; p < 2; p++) { s2[c][p] = F2{0.f, 0.f}; H2[c][p] = F2{0.f, 0.f}; Hl2[c][p] = F2{0.f, 0.f}; }
                const float *own = (const float *)sp + rowBase + Z0;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const float4 q4 = *reinterpret_cast<const float4 *>(own + (size_t)c * N);
                    s2[c][0] = F2{q4.x, q4.y}; s2[c][1] = F2{q4.z, q4.w};
                }
                auto nbOf = [&](auto L) { return nb_of(L, rowBase, edgeRow, wxp, wxm, wyp, wym); };
                cls.for_links(LinkCtx<float>{(const float *)sp, (size_t)N, Z0, Zd}, nbOf, [&](auto L) {
                    float t[3][4];
                    L.template fetch<float, 4, NC>((const float *)sp, nbOf(L), (size_t)N, Z0, Zd, t);
                    // aligned row, one diagonal exchange shared by every link of the class (its splat lives in one register
                    // pair): packed multiply-add straight from the float4.  Distinct tensors per link would each need their
                    // constant moved into a pair, where the scalar FFMA takes it as an immediate - measured slower (CrI3).
                    if (!FULLJ && cls.uniformJ() && L.cZ() == 0) {
#pragma unroll
                        for (int c = 0; c < NC; c++)
#pragma unroll
                            for (int p = 0; p < 2; p++) {
                                const F2 t2 = F2{t[c][2 * p], t[c][2 * p + 1]};
                                const F2 J2 = splat2((float)L.J(c));
                                H2[c][p] = fma2(J2, t2, H2[c][p]);
                            }
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            const float tx = t[0][v], ty = t[1][v], tz = NC == 3 ? t[2][v] : 0.f;
                            float hx, hy, hz = 0.f;
                            if (NC == 2) {
                                if (FULLJ) { hx = (float)L.J(0) * tx + (float)L.J(3) * ty; hy = (float)L.J(6) * tx + (float)L.J(1) * ty; }
                                else { hx = (float)L.J(0) * tx; hy = (float)L.J(1) * ty; }
                            } else if (FULLJ) {
                                hx = (float)L.J(0) * tx + (float)L.J(3) * ty + (float)L.J(4) * tz;
                                hy = (float)L.J(6) * tx + (float)L.J(1) * ty + (float)L.J(5) * tz;
                                hz = (float)L.J(7) * tx + (float)L.J(8) * ty + (float)L.J(2) * tz;
                            } else { hx = (float)L.J(0) * tx; hy = (float)L.J(1) * ty; hz = (float)L.J(2) * tz; }
                            float &Hx = (v & 1) ? H2[0][v >> 1].y : H2[0][v >> 1].x;
                            float &Hy = (v & 1) ? H2[1][v >> 1].y : H2[1][v >> 1].x;
                            Hx += hx; Hy += hy;
                            if (NC == 3) { float &Hz = (v & 1) ? H2[2][v >> 1].y : H2[2][v >> 1].x; Hz += hz; }
                        }
                    }
                }, [&]() {
                    // links are sorted with the lower-colour neighbours first: after the last of them the running sum IS the
                    // field of the final neighbours that the fused bond energy needs
                    if (MODE == 1 && lowmode == 2) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { Hl2[c][0] = H2[c][0]; Hl2[c][1] = H2[c][1]; }
                    }
                });
                const uint32_t id0 = (uint32_t)((xy + Z0 * pz + cls.cc()) * norb + cls.co());
                const float fbeta = (float)beta, fhf = (float)hf, fS = (float)S;
                F2 aM[3] = {F2{0.f, 0.f}, F2{0.f, 0.f}, F2{0.f, 0.f}}, aE = F2{0.f, 0.f};
                ItemWords<NC, 4> iw;   // the item's four sites share NC Philox blocks (rng.cuh)
                iw.begin(a.key, a.replica0 + r, sweep, id0, (uint32_t)idStrideZ, PARTIAL);
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    uint32_t wa[4], wb[4];
                    iw.need(a.key, a.replica0 + r, sweep, 2 * p + 2);
                    iw.lane(2 * p, PARTIAL, wa);
                    iw.lane(2 * p + 1, PARTIAL, wb);
                    const F2 sx = s2[0][p], sy = s2[1][p], sz = s2[2][p];
                    const F2 hx = H2[0][p], hy = H2[1][p], hz = H2[2][p];
                    // plane normal m = -n of random_dir: O(3) m = (r cos phi, r sin phi, -z), z = 2u-1; O(2) m = (cos phi, sin phi);
                    // phi = 2 pi (u' - 1/2) from the second (O(3)) / first (O(2)) Philox word
                    const F2 uphi = NC == 3 ? F2{u01<float>(wa[1]), u01<float>(wb[1])} : F2{u01<float>(wa[0]), u01<float>(wb[0])};
                    const F2 ph = mul2(add2(uphi, splat2(-0.5f)), splat2(6.283185307179586f));
                    float sa, ca, sb, cb;
                    __sincosf(ph.x, &sa, &ca);
                    __sincosf(ph.y, &sb, &cb);
                    F2 m0 = F2{ca, cb}, m1 = F2{sa, sb}, m2 = F2{0.f, 0.f};
                    if (NC == 3) {
                        const F2 u0 = F2{u01<float>(wa[0]), u01<float>(wb[0])};
                        m2 = fma2(u0, splat2(-2.f), splat2(1.f));
                        const F2 zp = fma2(u0, splat2(2.f), splat2(-1.f));
                        const F2 om = fma2(m2, zp, splat2(1.f));                     // 1 - z*z
                        const F2 rr = F2{r_sqrt<float>(om.x), r_sqrt<float>(om.y)};
                        m0 = mul2(rr, m0); m1 = mul2(rr, m1);
                    }
                    // heisenbergLib.c:451-456 with s1m = -2 (s.m): transSpin = s1m m, dE = s1m (beta m.H - hf m_axis) + on-site difference
                    F2 sm = fma2(sy, m1, mul2(sx, m0)), mH = fma2(m1, hy, mul2(m0, hx));
                    if (NC == 3) { sm = fma2(sz, m2, sm); mH = fma2(m2, hz, mH); }
                    const F2 s1m = mul2(sm, splat2(-2.f));
                    F2 dE = mul2(s1m, fma2(NC == 3 ? m2 : m0, splat2(-fhf), mul2(mH, splat2(fbeta))));
                    if (hasD) {
                        const F2 nx = fma2(s1m, m0, sx), ny = fma2(s1m, m1, sy);
                        const F2 neg1 = splat2(-1.f);
                        F2 dOn = mul2(splat2((float)D0), fma2(mul2(sx, sx), neg1, mul2(nx, nx)));
                        dOn = fma2(splat2((float)D1), fma2(mul2(sy, sy), neg1, mul2(ny, ny)), dOn);
                        if (NC == 3) {
                            const F2 nz = fma2(s1m, m2, sz);
                            dOn = fma2(splat2((float)D2), fma2(mul2(sz, sz), neg1, mul2(nz, nz)), dOn);
                        }
                        dE = fma2(splat2(fbeta), dOn, dE);
                    }
                    // heisenbergLib.c:461 accepts if dE <= 0 or exp(-dE) > u; u < 1 <= exp(-dE) for dE <= 0, so one test decides
                    const F2 ex = mul2(dE, splat2(-1.4426950408889634f));
                    float ea, eb_;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(ex.x));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb_) : "f"(ex.y));
                    const bool atta = PARTIAL ? (u01<float>(wa[3]) < (float)pAtt) : true;
                    const bool attb = PARTIAL ? (u01<float>(wb[3]) < (float)pAtt) : true;
                    const bool acca = atta & (ea > u01<float>(wa[2])), accb = attb & (eb_ > u01<float>(wb[2]));
                    const F2 ms = F2{acca ? s1m.x : 0.f, accb ? s1m.y : 0.f};   // rejected: s + 0*m = s exactly
                    F2 tx = fma2(ms, m0, sx), ty = fma2(ms, m1, sy), tz = NC == 3 ? fma2(ms, m2, sz) : sz;
                    if (renorm) {   // every site, accepted or not
                        F2 n2 = fma2(ty, ty, mul2(tx, tx));
                        if (NC == 3) n2 = fma2(tz, tz, n2);
                        const F2 f = mul2(F2{r_rsqrt<float>(n2.x), r_rsqrt<float>(n2.y)}, splat2(fS));
                        tx = mul2(tx, f); ty = mul2(ty, f);
                        if (NC == 3) tz = mul2(tz, f);
                    }
                    if (PARTIAL) natt += (atta ? 1 : 0) + (attb ? 1 : 0);
                    nacc += (acca ? 1 : 0) + (accb ? 1 : 0);
                    s2[0][p] = tx; s2[1][p] = ty; s2[2][p] = tz;
                    if (MODE == 1) {
                        aM[0] = add2(aM[0], tx); aM[1] = add2(aM[1], ty);
                        if (NC == 3) aM[2] = add2(aM[2], tz);
                        F2 e = fma2(NC == 3 ? tz : tx, splat2(-fhf), aE);
                        if (lowmode == 1) {
                            F2 eb2 = fma2(ty, hy, mul2(tx, hx));
                            if (NC == 3) eb2 = fma2(tz, hz, eb2);
                            e = fma2(splat2(fbeta), eb2, e);
                        } else if (lowmode == 2) {
                            F2 eb2 = fma2(ty, Hl2[1][p], mul2(tx, Hl2[0][p]));
                            if (NC == 3) eb2 = fma2(tz, Hl2[2][p], eb2);
                            e = fma2(splat2(fbeta), eb2, e);
                        }
                        if (hasD) {
                            F2 on = fma2(splat2((float)D1), mul2(ty, ty), mul2(splat2((float)D0), mul2(tx, tx)));
                            if (NC == 3) on = fma2(splat2((float)D2), mul2(tz, tz), on);
                            e = fma2(splat2(fbeta), on, e);
                        }
                        aE = e;
                    }
                }
                if (MODE == 1) {
                    acc_add((double)(aM[0].x + aM[0].y), (double)(aM[1].x + aM[1].y), NC == 3 ? (double)(aM[2].x + aM[2].y) : 0.0, (double)(aE.x + aE.y));
                }
                if (!PARTIAL) natt += V;
                float *ownw = (float *)sp + rowBase + Z0;
#pragma unroll
                for (int c = 0; c < NC; c++)
                    *reinterpret_cast<float4 *>(ownw + (size_t)c * N) = make_float4(s2[c][0].x, s2[c][0].y, s2[c][1].x, s2[c][1].y);
                continue;
            }
#endif
            real s[3][V], H[3][V], Hl[3][V];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int v = 0; v < V; v++) { s[c][v] = 0; H[c][v] = 0; Hl[c][v] = 0; }
            const real *own = sp + rowBase + Z0;
            cls.template own_early<real, V, NC>(own, (size_t)N, s);
            auto nbOf = [&](auto L) { return nb_of(L, rowBase, edgeRow, wxp, wxm, wyp, wym); };
            cls.for_links(LinkCtx<real>{sp, (size_t)N, Z0, Zd}, nbOf, [&](auto L) {
                real t[3][V];
                L.template fetch<real, V, NC>(sp, nbOf(L), (size_t)N, Z0, Zd, t);
                if constexpr (LinkPacked<decltype(L)>::value) {
                    // same multiply-add chain as the scalar full-tensor branch below, two sites per instruction: the LDS.128
                    // quads are aligned pairs already and the coefficients come from shared memory as (J, J)
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const F2 tx{(float)t[0][2 * p], (float)t[0][2 * p + 1]}, ty{(float)t[1][2 * p], (float)t[1][2 * p + 1]}, tz{(float)t[2][2 * p], (float)t[2][2 * p + 1]};
                        F2 h0{(float)H[0][2 * p], (float)H[0][2 * p + 1]}, h1{(float)H[1][2 * p], (float)H[1][2 * p + 1]}, h2{(float)H[2][2 * p], (float)H[2][2 * p + 1]};
                        h0 = fma2(L.J2(0), tx, fma2(L.J2(3), ty, fma2(L.J2(4), tz, h0)));
                        h1 = fma2(L.J2(6), tx, fma2(L.J2(1), ty, fma2(L.J2(5), tz, h1)));
                        h2 = fma2(L.J2(7), tx, fma2(L.J2(8), ty, fma2(L.J2(2), tz, h2)));
                        H[0][2 * p] = h0.x; H[0][2 * p + 1] = h0.y;
                        H[1][2 * p] = h1.x; H[1][2 * p + 1] = h1.y;
                        H[2][2 * p] = h2.x; H[2][2 * p + 1] = h2.y;
                    }
                } else {
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const real tx = t[0][v], ty = NC >= 2 ? t[1][v] : real(0), tz = NC == 3 ? t[2][v] : real(0);
                    real hx, hy = 0, hz = 0;
                    if (NC == 1) hx = L.J(0) * tx;
                    else if (NC == 2) {
                        if (FULLJ) { hx = L.J(0) * tx + L.J(3) * ty; hy = L.J(6) * tx + L.J(1) * ty; }
                        else { hx = L.J(0) * tx; hy = L.J(1) * ty; }
                    } else {
                        if (FULLJ) {
                            hx = L.J(0) * tx + L.J(3) * ty + L.J(4) * tz;
                            hy = L.J(6) * tx + L.J(1) * ty + L.J(5) * tz;
                            hz = L.J(7) * tx + L.J(8) * ty + L.J(2) * tz;
                        } else { hx = L.J(0) * tx; hy = L.J(1) * ty; hz = L.J(2) * tz; }
                    }
                    if (FULLJ) {
                        // full tensor: accumulate inside the multiply-add chain (three FFMA per component instead of FMUL + two
                        // FFMA + FADD)
                        if (NC == 2) {
                            H[0][v] = L.J(0) * tx + (L.J(3) * ty + H[0][v]);
                            H[1][v] = L.J(6) * tx + (L.J(1) * ty + H[1][v]);
                        } else {
                            H[0][v] = L.J(0) * tx + (L.J(3) * ty + (L.J(4) * tz + H[0][v]));
                            H[1][v] = L.J(6) * tx + (L.J(1) * ty + (L.J(5) * tz + H[1][v]));
                            H[2][v] = L.J(7) * tx + (L.J(8) * ty + (L.J(2) * tz + H[2][v]));
                        }
                    } else { H[0][v] += hx; H[1][v] += hy; H[2][v] += hz; }
                }
                }
            }, [&]() {
                // lower-colour neighbours come first in the link list: snapshot their field for the fused bond energy
                if (MODE == 1 && lowmode == 2) {
                    if constexpr (PARK) {
#pragma unroll
                        for (int c = 0; c < 3; c++) hls[c * 256 + tid] = make_float4((float)H[c][0], (float)H[c][1], (float)H[c][2], (float)H[c][3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; c++)
#pragma unroll
                            for (int v = 0; v < V; v++) Hl[c][v] = H[c][v];
                    }
                }
            });
            cls.template own_late<real, V, NC>(s);
            if constexpr (PARK) {
                if (lowmode == 2) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float4 h4 = hls[c * 256 + tid];
                        Hl[c][0] = (real)h4.x; Hl[c][1] = (real)h4.y; Hl[c][2] = (real)h4.z; Hl[c][3] = (real)h4.w;
                    }
                }
            }
            const uint32_t id0 = (uint32_t)((xy + Z0 * pz + cls.cc()) * norb + cls.co());
            ItemWords<NC, V> iw;   // V > 1: the item's sites share their Philox blocks (rng.cuh); V == 1: per-site streams
            if (V > 1) iw.begin(a.key, a.replica0 + r, sweep, id0, (uint32_t)idStrideZ, PARTIAL);
            real iM[3] = {0, 0, 0}, iE = 0;   // this item's contribution to the fused sums
#pragma unroll
            for (int v = 0; v < V; v++) {
                real sx = s[0][v], sy = s[1][v], sz = s[2][v];
                const real hx = H[0][v], hy = H[1][v], hz = H[2][v];
                uint32_t w[4];
                if (V > 1) { iw.need(a.key, a.replica0 + r, sweep, v + 1); iw.lane(v, PARTIAL, w); }
                else if (NC == 1) { IsingWords one; one.get(a.key, a.replica0 + r, sweep, id0, PARTIAL, w[2], w[3]); }
                else rng4(a.key, a.replica0 + r, STREAM_METRO, 0, sweep, id0, w);
                const bool att = PARTIAL ? (u01<real>(w[3]) < pAtt) : true;
                bool acc;
                if (NC == 1) {
                    const real corr = real(2) * (beta * sx * hx - hf * sx);                     // isingLib.c:242
                    // isingLib.c:244-252 accepts if corr >= 0 or exp(corr) > u; since u < 1 <= exp(corr) for corr >= 0 the
                    // second test alone decides identically
                    acc = att & metro_accept<real>(corr, w[2]);
                    sx = acc ? -sx : sx;
                } else {
                    real n[3];
                    random_dir<NC, real>(w[0], w[1], n);
                    // heisenbergLib.c:451-456: transSpin = s1n * n with s1n = -2 (s.n);  dE = transSpin . H + on-site difference.
                    // transSpin . H = s1n * (n . H): the proposal vector itself is only formed when the D terms need it.
                    const real s1n = real(-2) * (sx * n[0] + sy * n[1] + (NC == 3 ? sz * n[2] : real(0)));
                    const real nH = n[0] * hx + n[1] * hy + (NC == 3 ? n[2] * hz : real(0));
                    real nx = sx + s1n * n[0], ny = sy + s1n * n[1], nz = NC == 3 ? sz + s1n * n[2] : real(0);
                    real dE = s1n * (beta * nH - hf * (NC == 3 ? n[2] : n[0]));
                    if (hasD) {
                        real dOn = D0 * (nx * nx - sx * sx) + D1 * (ny * ny - sy * sy);
                        if (NC == 3) dOn += D2 * (nz * nz - sz * sz);
                        dE += beta * dOn;
                    }
                    // heisenbergLib.c:461 accepts if dE <= 0 or exp(-dE) > u; u < 1 <= exp(-dE) for dE <= 0, so one test decides
                    acc = att & metro_accept<real>(-dE, w[2]);
                    sx = acc ? nx : sx; sy = acc ? ny : sy; sz = acc ? nz : sz;
                    if (renorm) {   // every site, accepted or not
                        const real f = S * r_rsqrt<real>(sx * sx + sy * sy + sz * sz);
                        sx *= f; sy *= f; sz *= f;
                    }
                }
                if (PARTIAL) natt += att ? 1 : 0;
                nacc += acc ? 1 : 0;
                s[0][v] = sx; s[1][v] = sy; s[2][v] = sz;
                if (MODE == 1) {
                    iM[0] += sx; iM[1] += sy; iM[2] += sz;
                    real eb;
                    if (lowmode == 1) eb = sx * hx + sy * hy + sz * hz;
                    else if (lowmode == 2) eb = sx * Hl[0][v] + sy * Hl[1][v] + sz * Hl[2][v];
                    else eb = real(0);
                    real eon = -hf * (NC == 3 ? sz : sx);
                    if (hasD && NC > 1) eon += beta * (D0 * sx * sx + D1 * sy * sy + (NC == 3 ? D2 * sz * sz : real(0)));
                    iE += beta * eb + eon;
                }
            }
            if (MODE == 1) acc_add((double)iM[0], (double)iM[1], (double)iM[2], (double)iE);
            if (!PARTIAL) natt += V;
            real *ownw = sp + rowBase + Z0;
#pragma unroll
            for (int c = 0; c < NC; c++) vstore<real, V>(ownw + (size_t)c * N, s[c]);
        }
    }
    natt = __reduce_add_sync(0xffffffffu, natt);
    nacc = __reduce_add_sync(0xffffffffu, nacc);
    if ((tid & 31) == 0 && natt) {
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ATTEMPT, (unsigned long long)natt);
        atomicAdd(a.cnt + (size_t)r * NCNT + CNT_ACCEPT, (unsigned long long)nacc);
    }
    if (MODE == 1) {
        double v[4] = {(double)accM[0], (double)accM[1], (double)accM[2], (double)accE};
        const int lane = tid & 31, w = tid >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
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
                if (lane == 0 && sum != 0.0) atomicAdd(a.classSums + ((size_t)r * nclass + q) * 4 + i, sum);
            }
        }
    }
}

#ifndef MCG_JIT
// offline entry: tables as a __grid_constant__ parameter
// full 3x3 tensors on three components (DMI, dipole stencils: up to 32 links) need more than 64 registers to keep the
// four sites' field sums, the fused-energy snapshot and a neighbour's 12 values live: 3 resident blocks instead of 4
template <int NC, typename real, bool FULLJ, int MODE, int V, bool PARTIAL>
__global__ void __launch_bounds__(256, (FULLJ && NC == 3 && sizeof(real) == 4) ? 3 : 4)
k_struct_fast(const __grid_constant__ StructArgs a, const __grid_constant__ PassTable<real> T, int q0, int rowsPerBlock, int nrb,
              uint64_t sweep, real pAtt) {
    __shared__ double red[4 * 32];
    constexpr bool PARK = FULLJ && NC == 3 && MODE == 1 && V == 4 && sizeof(real) == 4;
    __shared__ float4 hls[PARK ? 3 * 256 : 1];
    const int nqc = T.nqc;
    const int bid = blockIdx.x;
    const int j = bid % nqc, tq = bid / nqc, rb = tq % nrb, r = tq / nrb;
    pass_body<NC, real, FULLJ, MODE, V, PARTIAL>(a, RtClass<real>{T, j}, q0 + j, r, rb, rowsPerBlock, sweep, pAtt, red, hls);
}
// full-tensor fp32 pass with the asynchronous link pipeline (AsClass); dynamic shared memory = ASYNC_D stages
template <int MODE, bool PARTIAL>
__global__ void __launch_bounds__(256, 3)
k_struct_async(const __grid_constant__ StructArgs a, const __grid_constant__ PassTable<float> T, const AsTab *__restrict__ gtab, int q0,
               int rowsPerBlock, int nrb, uint64_t sweep, float pAtt) {
    __shared__ double red[4 * 32];
    __shared__ float4 hls[MODE == 1 ? 3 * 256 : 1];
    extern __shared__ float4 async_stages[];
    const int nqc = T.nqc;
    const int bid = blockIdx.x;
    const int j = bid % nqc, tq = bid / nqc, rb = tq % nrb, r = tq / nrb;
    __shared__ AsTab tab[PT_MAXL + 1];
    AsClass cls{{T, j}, async_stages, tab};
    cls.stage_table(tab, gtab + (size_t)(q0 + j) * (PT_MAXL + 1));
    pass_body<3, float, true, MODE, 4, PARTIAL>(a, cls, q0 + j, r, rb, rowsPerBlock, sweep, pAtt, red, hls);
}
#else
// ---- JIT entry points: JIT_NC, jit_real, JIT_FULLJ, JIT_V, JIT_NQC, JIT_PARTIAL, JIT_MINB, CtLinkData<J,K>, CtClassData<J>
// come from the generated prologue ----
template <int JJ, int K> struct CtLink {
    typedef CtLinkData<JJ, K> D;
    __device__ __forceinline__ constexpr int delta() const { return D::delta; }
    __device__ __forceinline__ constexpr int mxp() const { return D::mxp; }
    __device__ __forceinline__ constexpr int mxm() const { return D::mxm; }
    __device__ __forceinline__ constexpr int myp() const { return D::myp; }
    __device__ __forceinline__ constexpr int mym() const { return D::mym; }
    __device__ __forceinline__ constexpr int cZ() const { return D::cZ; }
    __device__ __forceinline__ constexpr int low() const { return D::low; }
    __device__ __forceinline__ constexpr int idx() const { return K; }
    __device__ __forceinline__ constexpr jit_real J(int e) const { return D::J(e); }
    template <typename R, int V, int NC> __device__ __forceinline__ void fetch(const R *__restrict__ sp, int nb, size_t N, int Z0, int Zd, R (&t)[3][V]) const {
        load_shifted_nc<R, V, NC>(sp, N, nb, Z0, D::cZ, Zd, t);
    }
};
template <int JJ> struct CtClass {
    typedef CtClassData<JJ> C;
    __device__ __forceinline__ constexpr int nl() const { return C::nl; }
    __device__ __forceinline__ constexpr int ca() const { return C::ca; }
    __device__ __forceinline__ constexpr int cb() const { return C::cb; }
    __device__ __forceinline__ constexpr int cc() const { return C::cc; }
    __device__ __forceinline__ constexpr int co() const { return C::co; }
    __device__ __forceinline__ constexpr int lowmode() const { return C::lowmode; }
    __device__ __forceinline__ constexpr int nlow() const { return C::nlow; }
    __device__ __forceinline__ constexpr jit_real S() const { return C::S; }
    __device__ __forceinline__ constexpr jit_real D(int e) const { return C::D(e); }
    // the item's own spins: read up front (the asynchronous pipeline fetches them behind the last link instead)
    template <typename R, int V, int NC> __device__ __forceinline__ void own_early(const R *__restrict__ own, size_t N, R (&sv)[3][V]) const {
#pragma unroll
        for (int c = 0; c < NC; c++) vload<R, V>(own + (size_t)c * N, sv[c]);
    }
    template <typename R, int V, int NC> __device__ __forceinline__ void own_late(R (&)[3][V]) const {}
    template <int K> static __device__ __forceinline__ constexpr bool uj_from() {
        if constexpr (K >= C::nl) return true;
        else
            return CtLinkData<JJ, K>::J(0) == CtLinkData<JJ, 0>::J(0) && CtLinkData<JJ, K>::J(1) == CtLinkData<JJ, 0>::J(1) &&
                   CtLinkData<JJ, K>::J(2) == CtLinkData<JJ, 0>::J(2) && uj_from<K + 1>();
    }
    __device__ __forceinline__ constexpr bool uniformJ() const { return uj_from<0>(); }
    template <typename X, typename P, typename F, typename G> __device__ __forceinline__ void for_links(const X &, P &&, F &&f, G &&snap) const {
        if constexpr (C::nlow == 0) snap();
        ct_for<0, C::nl>([&](auto k) {
            f(CtLink<JJ, decltype(k)::value>{});
            if constexpr (decltype(k)::value == C::nlow - 1) snap();
        });
    }
};

template <int MODE, int J>
__device__ __forceinline__ void jit_case(const StructArgs &a, int j, int q0, int r, int rb, int rowsPerBlock, uint64_t sweep,
                                         jit_real pAtt, double *red) {
    if constexpr (J < JIT_NQC) {
        if (j == J) pass_body<JIT_NC, jit_real, JIT_FULLJ, MODE, JIT_V, JIT_PARTIAL>(a, CtClass<J>{}, q0 + J, r, rb, rowsPerBlock, sweep, pAtt, red);
        else jit_case<MODE, J + 1>(a, j, q0, r, rb, rowsPerBlock, sweep, pAtt, red);
    }
}

#define MCG_JIT_ENTRY(NAME, MODE)                                                                                              \
    extern "C" __global__ void __launch_bounds__(256, JIT_MINB)                                                                \
    NAME(const __grid_constant__ StructArgs a, int q0, int rowsPerBlock, int nrb, uint64_t sweep, jit_real pAtt) {             \
        __shared__ double red[4 * 32];                                                                                         \
        const int bid = blockIdx.x;                                                                                            \
        const int j = bid % JIT_NQC, tq = bid / JIT_NQC, rb = tq % nrb, r = tq / nrb;                                          \
        jit_case<MODE, 0>(a, j, q0, r, rb, rowsPerBlock, sweep, pAtt, red);                                                    \
    }
MCG_JIT_ENTRY(mcg_pass_m0, 0)
MCG_JIT_ENTRY(mcg_pass_m1, 1)
#endif

}  // namespace mcg
