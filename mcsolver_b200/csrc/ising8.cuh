// Ising spins as int8 (precision = 8): one byte per spin - 0 = up (+|S|), 1 = down (-|S|) - 16 sites per thread.
//
// Why: the Ising colour pass is the one workload whose state does not need floating point.  With fp32 planes an attempt
// moves 12 bytes (own read + own write + every neighbour-colour spin once); with one byte per spin it moves 3 (SURVEY 8d:
// "Ising int8 w = 1 -> 3 B/attempt"), and the per-site arithmetic collapses to integer byte tricks:
//   * a thread holds an ITEM of 16 consecutive sites = one 16-byte load per neighbour row;
//   * "number of DOWN neighbours" is summed for four sites at once by plain 32-bit adds of the neighbour words (bytes are
//     0 / 1, up to 32 rows cannot carry into the next byte), rows shifted by one cell are rebuilt with funnel shifts;
//   * every link carries the same exchange and every site the same |S| (checked when the system is created), so the
//     acceptance probability depends on (own spin, number of down neighbours) only: 2 (z + 1) values per replica, turned
//     into 32-bit integer thresholds once per block (fp64 exp) and kept in shared memory.  The test
//         exp(corr) > (w + 1/2) / 2^32      (isingLib.c:242-252 with the engine's fp64 uniform, rng.cuh u01<double>)
//     becomes  w < T,  T = ceil(exp(corr) 2^32 - 1/2)  - the same decision for every 32-bit word w, so the trajectory is the
//     one the oracle's fp64 restatement produces, bit for bit (only exception: a flip less probable than 2^-33 is
//     attempted with 2^-32, see the threshold table below);
//   * one Philox word per attempt, four Philox blocks per item (ItemWords<1, 16>);
//   * magnetisation and bond energy of the fused measurement are integer sums (popcount, dp4a) - exact.
// The pass is bound by the integer ALU pipe (half rate), not by HBM: per attempt ~10 Philox instructions (5 IMAD.WIDE on
// the FMA pipe, 5 LOP3), ~2 for the neighbour sums and 4 for the test (PRMT, LDS, ISETP, predicated LOP3).
//
// Reference: isingLib.c:238-254 (localUpdate), :121-127 (energy), :395-421 (per-sweep sums).
#pragma once
#include "struct_pass.cuh"

namespace mcg {

constexpr int I8_MAXZ = 32;

struct I8Class {          // what the pass needs to know about one class of the colour
    int nl, nlow, lowmode, ca, cb, cc, co;
    int delta[I8_MAXZ];   // neighbour row offset (cells), wrap corrections as in PLink
    signed char wx[I8_MAXZ], wy[I8_MAXZ], cz[I8_MAXZ];
};
struct I8Table {
    int nqc;
    double JS2, S;        // exchange * S^2 (table units), |S|
    I8Class c[PT_MAXC];
};

// neighbour row words shifted by one cell: o = bytes [Z0 + cz, Z0 + cz + 4 NW) of the row; e = the byte that enters
template <int NW> __device__ __forceinline__ void i8_shift(uint32_t (&t)[NW], int cz, uint32_t e) {
    if (cz < 0) {   // o[i] = row[Z0 - 1 + i]
#pragma unroll
        for (int i = NW - 1; i > 0; i--) t[i] = __funnelshift_l(t[i - 1], t[i], 8);
        t[0] = (t[0] << 8) | e;
    } else {        // o[i] = row[Z0 + 1 + i]
#pragma unroll
        for (int i = 0; i < NW - 1; i++) t[i] = __funnelshift_r(t[i], t[i + 1], 8);
        t[NW - 1] = (t[NW - 1] >> 8) | (e << 24);
    }
}

template <int NW> __device__ __forceinline__ void i8_load(const unsigned char *__restrict__ p, uint32_t (&o)[NW]) {
    if (NW == 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[NW - 1] = v.w;
    } else o[0] = *reinterpret_cast<const uint32_t *>(p);
}
template <int NW> __device__ __forceinline__ void i8_store(unsigned char *__restrict__ p, const uint32_t (&o)[NW]) {
    if (NW == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[NW - 1]);
    else *reinterpret_cast<uint32_t *>(p) = o[0];
}

// dims: runtime fields offline, literals under JIT (the generated prologue defines JIT_Xd ...)
#ifdef MCG_JIT_I8
#define I8_DIM(a, f) (JIT_##f)
#else
#define I8_DIM(a, f) ((a).f)
#endif

// runtime view of a class (offline build): link k's data comes from the table in the constant bank
struct RtI8Class {
    const I8Class &c;
    static constexpr int NLMAX = 8;            // link lists up to this length keep their row offsets in registers
    __device__ __forceinline__ int nl() const { return c.nl; }
    __device__ __forceinline__ int nlow() const { return c.nlow; }
    __device__ __forceinline__ int lowmode() const { return c.lowmode; }
    __device__ __forceinline__ int ca() const { return c.ca; }
    __device__ __forceinline__ int cb() const { return c.cb; }
    __device__ __forceinline__ int cc() const { return c.cc; }
    __device__ __forceinline__ int co() const { return c.co; }
    __device__ __forceinline__ int delta(int k) const { return c.delta[k]; }
    __device__ __forceinline__ int wx(int k) const { return c.wx[k]; }
    __device__ __forceinline__ int wy(int k) const { return c.wy[k]; }
    __device__ __forceinline__ int cz(int k) const { return c.cz[k]; }
};

// One colour pass over the rows [rb*rowsPerBlock, ...) of class q for replica r.
// MODE 0: update   1: update + fused measurement   2: measurement only (energy with the 1/2 of double counting)
// SMALL: the class's link list fits CLS::NLMAX (row offsets in registers, loop unrolled); compile-time classes always do.
template <int MODE, bool PARTIAL, int NW, bool SMALL, typename CLS>
__device__ __forceinline__ void i8_body(const StructArgs &a, const CLS cl, double JS2, double S, int q, int r, int rb, int rowsPerBlock, uint64_t sweep,
                                        double pAtt, uint32_t *lut, double *red) {
    constexpr int V = 4 * NW;
    constexpr int NLMAX = CLS::NLMAX;
    const int nl = cl.nl();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const double beta = a.beta[r], hf = a.beta[r] * a.field[r];
    if (MODE != 2 && tid < 2 * (nl + 1)) {
        const int nd = tid % (nl + 1), dn = tid / (nl + 1);
        const double sg = dn ? -1.0 : 1.0;
        const double corr = 2.0 * sg * (beta * JS2 * (double)(nl - 2 * nd) - hf * S);   // isingLib.c:242
        // accept  <=>  exp(corr) > (w + 1/2) / 2^32  <=>  w < Tn,  Tn = ceil(exp(corr) 2^32 - 1/2) in [0, 2^32]; stored as
        // Tn - 1 for a <= test, which covers Tn = 2^32 (every word accepted: all corr >= 0) exactly.  Tn = 0 (a flip
        // less likely than 2^-33, corr < -22.9) is stored as Tn = 1.
        const double thr = ceil(exp(corr) * 4294967296.0 - 0.5);
        lut[tid] = !(thr < 4294967296.0) ? 0xffffffffu : (thr >= 1.0 ? (uint32_t)thr - 1u : 0u);
    }
    __syncthreads();
    const uint32_t pthr = PARTIAL ? (uint32_t)fmin(4294967295.0, fmax(0.0, ceil(pAtt * 4294967296.0 - 0.5))) : 0u;   // (w + 1/2) / 2^32 < pAtt
    const int Xd = I8_DIM(a, Xd), Yd = I8_DIM(a, Yd), Zd = I8_DIM(a, Zd), Zc = I8_DIM(a, Zc), N = I8_DIM(a, N);
    const int px = I8_DIM(a, px), py = I8_DIM(a, py), pz = I8_DIM(a, pz), norb = I8_DIM(a, norb);
    const int Ly = I8_DIM(a, Ly), Lz = I8_DIM(a, Lz), nrows = I8_DIM(a, nrows), nclass = I8_DIM(a, nclass);
    unsigned char *sp = (unsigned char *)a.spin + (size_t)r * N;
    const int planeY = Yd * Zd, planeX = planeY * Xd;
    const int idStrideZ = pz * norb;
    const uint32_t idxMul = 4u * (uint32_t)(nl + 1);
    int natt = 0, nacc = 0;
    long long mDown = 0, eInt = 0;    // number of down spins; sum_i sigma_i (n_i - 2 down-neighbours_i) over the links that count
    int nSites = 0;
    const int rowEnd = min(nrows, (rb + 1) * rowsPerBlock);
    for (int row = rb * rowsPerBlock + threadIdx.y; row < rowEnd; row += blockDim.y) {
        const int X = row / Yd, Y = row - X * Yd;
        const int rowBase = ((q * Xd + X) * Yd + Y) * Zd;
        const int wxp = X == Xd - 1 ? -planeX : 0, wxm = X == 0 ? planeX : 0;
        const int wyp = Y == Yd - 1 ? -planeY : 0, wym = Y == 0 ? planeY : 0;
        const int xy = ((X * px + cl.ca()) * Ly + (Y * py + cl.cb())) * Lz;
        auto rowOf = [&](int k) {
            const int wx = cl.wx(k), wy = cl.wy(k);
            return rowBase + cl.delta(k) + (wx > 0 ? wxp : 0) + (wx < 0 ? wxm : 0) + (wy > 0 ? wyp : 0) + (wy < 0 ? wym : 0);
        };
        int nbr[SMALL ? NLMAX : 1];
        uint32_t czs = 0u;          // (cz + 1) of link k in bits 2k, 2k + 1
        if (SMALL) {
#pragma unroll
            for (int k = 0; k < NLMAX; k++)
                if (k < nl) { nbr[k] = rowOf(k); czs |= (uint32_t)(cl.cz(k) + 1) << (2 * k); }
        }
        for (int zc = threadIdx.x; zc < Zc; zc += blockDim.x) {
            const int Z0 = zc * V;
            uint32_t own[NW], nd[NW], ndl[NW];
            i8_load<NW>(sp + rowBase + Z0, own);
#pragma unroll
            for (int i = 0; i < NW; i++) { nd[i] = 0u; ndl[i] = 0u; }
            const int zl = Z0 == 0 ? Zd - 1 : Z0 - 1, zr = Z0 + V >= Zd ? 0 : Z0 + V;
            auto link = [&](int k, int nb, int cz) {
                uint32_t t[NW];
                i8_load<NW>(sp + nb + Z0, t);
                if (cz != 0) i8_shift<NW>(t, cz, (uint32_t)sp[nb + (cz < 0 ? zl : zr)]);
#pragma unroll
                for (int i = 0; i < NW; i++) nd[i] += t[i];        // bytes are 0 / 1: the sum of <= 32 rows cannot carry
                if (MODE == 1 && k == cl.nlow() - 1) {
#pragma unroll
                    for (int i = 0; i < NW; i++) ndl[i] = nd[i];
                }
            };
            if (SMALL) {
#pragma unroll
                for (int k = 0; k < NLMAX; k++)
                    if (k < nl) link(k, nbr[k], (int)((czs >> (2 * k)) & 3u) - 1);
            } else {
                for (int k = 0; k < nl; k++) link(k, rowOf(k), cl.cz(k));
            }
            if (MODE != 2) {
                const uint32_t id0 = (uint32_t)((xy + Z0 * pz + cl.cc()) * norb + cl.co());
                ItemWords<1, V> iw;
                iw.begin(a.key, a.replica0 + r, sweep, id0, (uint32_t)idStrideZ, PARTIAL);
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    iw.need(a.key, a.replica0 + r, sweep, 4 * i + 4);
                    const uint32_t idx4 = own[i] * idxMul + nd[i] * 4u;      // byte b: offset of the site's entry in lut[]
                    uint32_t flip = 0u;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint32_t off = b == 3 ? idx4 >> 24 : __byte_perm(idx4, 0u, 0x4440u + b);
                        const uint32_t thr = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(lut) + off);
                        bool acc = iw.c[i][b] <= thr;
                        if constexpr (PARTIAL) {
                            const bool att = iw.p[i][b] < pthr;
                            natt += att ? 1 : 0;
                            acc = acc & att;
                        }
                        if (acc) flip |= 1u << (8 * b);
                    }
                    own[i] ^= flip;
                    nacc += __popc(flip);
                }
                if (!PARTIAL) natt += V;
                i8_store<NW>(sp + rowBase + Z0, own);
            }
            if (MODE != 0) {
                // sum_i sigma_i (n - 2 d_i) with d_i the down neighbours that count:  sum(n - 2 d) - 2 sum_{i down}(n - 2 d_i)
                const int n = MODE == 2 ? nl : (cl.lowmode() == 1 ? nl : (cl.lowmode() == 2 ? cl.nlow() : 0));
                int down = 0;
                unsigned sd = 0u, sdd = 0u;
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    const uint32_t cnt = MODE == 2 || cl.lowmode() == 1 ? nd[i] : ndl[i];
                    down += __popc(own[i]);
                    sd = __dp4a(cnt, 0x01010101u, sd);
                    sdd = __dp4a(cnt, own[i], sdd);
                }
                mDown += down;
                nSites += V;
                if (n > 0) eInt += (long long)(V * n - 2 * (int)sd) - 2ll * (long long)(n * down - 2 * (int)sdd);
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
        // M = S (sites - 2 down);  E = beta J S^2 [1/2 in MODE 2] sum - hf S (sites - 2 down)     (isingLib.c:121-127, 232)
        const double msum = (double)((long long)nSites - 2 * mDown);
        double v[4] = {S * msum, 0.0, 0.0, beta * JS2 * (MODE == 2 ? 0.5 : 1.0) * (double)eInt - hf * S * msum};
        const int lane = tid & 31, w = tid >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
        for (int i = 0; i < 4; i += 3) {
            double sum = warp_sum(v[i]);
            if (lane == 0) red[i * 32 + w] = sum;
        }
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int i = 0; i < 4; i += 3) {
                double sum = lane < nw ? red[i * 32 + lane] : 0.0;
                sum = warp_sum(sum);
                if (lane == 0 && sum != 0.0) atomicAdd(a.classSums + ((size_t)r * nclass + q) * 4 + i, sum);
            }
        }
    }
}

#ifndef MCG_JIT_I8
// offline entry: link tables as a __grid_constant__ parameter
template <int MODE, bool PARTIAL, int NW, bool SMALL>
__global__ void __launch_bounds__(256, 4)
k_i8_pass(const __grid_constant__ StructArgs a, const __grid_constant__ I8Table T, int q0, int rowsPerBlock, int nrb, uint64_t sweep, double pAtt) {
    __shared__ uint32_t lut[2 * (I8_MAXZ + 1)];     // threshold - 1 by (own spin down ? z + 1 : 0) + down neighbours
    __shared__ double red[4 * 32];
    const int bid = blockIdx.x;
    const int j = bid % T.nqc, tq = bid / T.nqc, rb = tq % nrb, r = tq / nrb;
    i8_body<MODE, PARTIAL, NW, SMALL>(a, RtI8Class{T.c[j]}, T.JS2, T.S, q0 + j, r, rb, rowsPerBlock, sweep, pAtt, lut, red);
}
#else
// ---- JIT entry points: the lattice is the program.  JIT_NW, JIT_NQC, JIT_PARTIAL, JIT_JS2, JIT_S, the dims and I8Ct<J> come
// from the generated prologue ----
template <int JJ> struct CtI8Class {
    typedef I8Ct<JJ> C;
    static constexpr int NLMAX = C::nl;
    __device__ __forceinline__ constexpr int nl() const { return C::nl; }
    __device__ __forceinline__ constexpr int nlow() const { return C::nlow; }
    __device__ __forceinline__ constexpr int lowmode() const { return C::lowmode; }
    __device__ __forceinline__ constexpr int ca() const { return C::ca; }
    __device__ __forceinline__ constexpr int cb() const { return C::cb; }
    __device__ __forceinline__ constexpr int cc() const { return C::cc; }
    __device__ __forceinline__ constexpr int co() const { return C::co; }
    __device__ __forceinline__ constexpr int delta(int k) const { return C::delta(k); }
    __device__ __forceinline__ constexpr int wx(int k) const { return C::wx(k); }
    __device__ __forceinline__ constexpr int wy(int k) const { return C::wy(k); }
    __device__ __forceinline__ constexpr int cz(int k) const { return C::cz(k); }
};
template <int MODE, int J>
__device__ __forceinline__ void i8_case(const StructArgs &a, int j, int q0, int r, int rb, int rowsPerBlock, uint64_t sweep, double pAtt, uint32_t *lut,
                                        double *red) {
    if constexpr (J < JIT_NQC) {
        if (j == J) i8_body<MODE, JIT_PARTIAL, JIT_NW, true>(a, CtI8Class<J>{}, JIT_JS2, JIT_S, q0 + J, r, rb, rowsPerBlock, sweep, pAtt, lut, red);
        else i8_case<MODE, J + 1>(a, j, q0, r, rb, rowsPerBlock, sweep, pAtt, lut, red);
    }
}
#define MCG_I8_ENTRY(NAME, MODE)                                                                                               \
    extern "C" __global__ void __launch_bounds__(256, JIT_MINB)                                                                \
    NAME(const __grid_constant__ StructArgs a, int q0, int rowsPerBlock, int nrb, uint64_t sweep, double pAtt) {               \
        __shared__ uint32_t lut[2 * (I8_MAXZ + 1)];                                                                            \
        __shared__ double red[4 * 32];                                                                                         \
        const int bid = blockIdx.x;                                                                                            \
        const int j = bid % JIT_NQC, tq = bid / JIT_NQC, rb = tq % nrb, r = tq / nrb;                                          \
        i8_case<MODE, 0>(a, j, q0, r, rb, rowsPerBlock, sweep, pAtt, lut, red);                                                \
    }
MCG_I8_ENTRY(mcg_pass_m0, 0)
MCG_I8_ENTRY(mcg_pass_m1, 1)
#endif

// ---- small companions of the pass: the other kernels that touch the int8 planes ----
// site id <-> storage position are the structured path's (structured.cu); the kernels are instantiated there.

}  // namespace mcg
