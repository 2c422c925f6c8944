// Ising spins as int8 (precision = 8): one byte per spin, +1 / -1, 16 sites per thread.
//
// Why: the Ising colour pass is the one workload whose state does not need floating point.  With fp32 planes an attempt
// moves 12 bytes (own read + own write + every neighbour-colour spin once); with one byte per spin it moves 3 (SURVEY 8d:
// "Ising int8 w = 1 -> 3 B/attempt"), and the per-site arithmetic collapses to integer byte tricks:
//   * a thread holds an ITEM of 16 consecutive sites = one 16-byte load per neighbour row;
//   * "number of DOWN neighbours" is summed for four sites at once in the bytes of a 32-bit word (bit 1 of a spin byte is
//     set exactly for -1 = 0xFF), rows shifted by one cell are rebuilt with funnel shifts;
//   * every link carries the same exchange and every site the same |S| (checked when the system is created), so the
//     acceptance probability depends on (own spin, number of down neighbours) only: 2 (z + 1) values per replica, turned
//     into 32-bit integer thresholds once per block (fp64 exp) and kept in shared memory.  The test
//         exp(corr) > (w + 1/2) / 2^32      (isingLib.c:242-252 with the engine's fp64 uniform, rng.cuh u01<double>)
//     becomes  w < T,  T = ceil(exp(corr) 2^32 - 1/2)  - the same decision for every 32-bit word w, so the trajectory is the
//     one the oracle's fp64 restatement produces, bit for bit;
//   * one Philox word per attempt, four Philox blocks per item (ItemWords<1, 16>);
//   * magnetisation and bond energy of the fused measurement are integer sums (popcount, dp4a) - exact.
// Instruction budget per attempt: ~10 Philox + ~2 neighbour sums + 4-5 threshold test and flip.
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

__device__ __forceinline__ uint32_t i8_down(uint32_t w) { return (w >> 1) & 0x01010101u; }   // 1 where the byte is -1

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

template <int NW> __device__ __forceinline__ void i8_load(const signed char *__restrict__ p, uint32_t (&o)[NW]) {
    if (NW == 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[NW - 1] = v.w;
    } else o[0] = *reinterpret_cast<const uint32_t *>(p);
}
template <int NW> __device__ __forceinline__ void i8_store(signed char *__restrict__ p, const uint32_t (&o)[NW]) {
    if (NW == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[NW - 1]);
    else *reinterpret_cast<uint32_t *>(p) = o[0];
}

// MODE 0: update   1: update + fused measurement   2: measurement only (energy with the 1/2 of double counting)
template <int MODE, bool PARTIAL, int NW>
__global__ void __launch_bounds__(256, 4)
k_i8_pass(const __grid_constant__ StructArgs a, const __grid_constant__ I8Table T, int q0, int rowsPerBlock, int nrb, uint64_t sweep, double pAtt) {
    constexpr int V = 4 * NW;
    __shared__ uint2 lut[2 * (I8_MAXZ + 1)];     // (threshold, always) by (own spin down ? z + 1 : 0) + down neighbours
    __shared__ double red[4 * 32];
    const int bid = blockIdx.x;
    const int j = bid % T.nqc, tq = bid / T.nqc, rb = tq % nrb, r = tq / nrb;
    const I8Class &cl = T.c[j];
    const int q = q0 + j, nl = cl.nl;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const double beta = a.beta[r], hf = a.beta[r] * a.field[r];
    if (MODE != 2 && tid < 2 * (nl + 1)) {
        const int nd = tid % (nl + 1), dn = tid / (nl + 1);
        const double sg = dn ? -1.0 : 1.0;
        const double corr = 2.0 * sg * (beta * T.JS2 * (double)(nl - 2 * nd) - hf * T.S);   // isingLib.c:242
        const double E = exp(corr) * 4294967296.0;
        double thr = ceil(E - 0.5);
        uint2 e;
        if (!(thr < 4294967296.0)) e = make_uint2(0u, 1u);          // accepted whatever the word (corr >= 0 lands here)
        else e = make_uint2(thr > 0.0 ? (uint32_t)thr : 0u, 0u);
        lut[tid] = e;
    }
    __syncthreads();
    const uint32_t pthr = PARTIAL ? (uint32_t)fmin(4294967295.0, fmax(0.0, ceil(pAtt * 4294967296.0 - 0.5))) : 0u;   // (w + 1/2) / 2^32 < pAtt
    signed char *sp = (signed char *)a.spin + (size_t)r * a.N;
    const int Xd = a.Xd, Yd = a.Yd, Zd = a.Zd, Zc = a.Zc;
    const int planeY = Yd * Zd, planeX = planeY * Xd;
    const int idStrideZ = a.pz * a.norb;
    int natt = 0, nacc = 0;
    long long mDown = 0, eInt = 0;    // number of down spins; sum_i sigma_i (n_i - 2 down-neighbours_i) over the links that count
    int nSites = 0;
    const int rowEnd = min(a.nrows, (rb + 1) * rowsPerBlock);
    for (int row = rb * rowsPerBlock + threadIdx.y; row < rowEnd; row += blockDim.y) {
        const int X = row / Yd, Y = row - X * Yd;
        const int rowBase = ((q * Xd + X) * Yd + Y) * Zd;
        const int wxp = X == Xd - 1 ? -planeX : 0, wxm = X == 0 ? planeX : 0;
        const int wyp = Y == Yd - 1 ? -planeY : 0, wym = Y == 0 ? planeY : 0;
        const int xy = ((X * a.px + cl.ca) * a.Ly + (Y * a.py + cl.cb)) * a.Lz;
        for (int zc = threadIdx.x; zc < Zc; zc += blockDim.x) {
            const int Z0 = zc * V;
            uint32_t own[NW], nd[NW], ndl[NW];
            i8_load<NW>(sp + rowBase + Z0, own);
#pragma unroll
            for (int i = 0; i < NW; i++) { nd[i] = 0u; ndl[i] = 0u; }
            for (int k = 0; k < nl; k++) {
                int nb = rowBase + cl.delta[k];
                const int wx = cl.wx[k], wy = cl.wy[k], cz = cl.cz[k];
                nb += (wx > 0 ? wxp : 0) + (wx < 0 ? wxm : 0) + (wy > 0 ? wyp : 0) + (wy < 0 ? wym : 0);
                uint32_t t[NW];
                i8_load<NW>(sp + nb + Z0, t);
                if (cz != 0) {
                    const int ze = cz < 0 ? (Z0 == 0 ? Zd - 1 : Z0 - 1) : (Z0 + V >= Zd ? 0 : Z0 + V);
                    i8_shift<NW>(t, cz, (uint32_t)(unsigned char)sp[nb + ze]);
                }
#pragma unroll
                for (int i = 0; i < NW; i++) nd[i] += t[i] & 0x02020202u;      // bit 1 of a spin byte: set for -1 only
                if (MODE == 1 && k == cl.nlow - 1) {
#pragma unroll
                    for (int i = 0; i < NW; i++) ndl[i] = nd[i];
                }
            }
#pragma unroll
            for (int i = 0; i < NW; i++) { nd[i] >>= 1; ndl[i] >>= 1; }       // byte sums stay below 2 * 32: no carry between bytes
            if (MODE != 2) {
                const uint32_t id0 = (uint32_t)((xy + Z0 * a.pz + cl.cc) * a.norb + cl.co);
                ItemWords<1, V> iw;
                iw.begin(a.key, a.replica0 + r, sweep, id0, (uint32_t)idStrideZ, PARTIAL);
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    iw.need(a.key, a.replica0 + r, sweep, 4 * i + 4);
                    const uint32_t idx = nd[i] + i8_down(own[i]) * (uint32_t)(nl + 1);
                    uint32_t flip = 0u;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint2 e = lut[(idx >> (8 * b)) & 0xffu];
                        bool acc = (iw.c[i][b] < e.x) | (e.y != 0u);
                        if (PARTIAL) {
                            const bool att = iw.p[i][b] < pthr;
                            natt += att ? 1 : 0;
                            acc = acc & att;
                        }
                        flip |= acc ? (0xfeu << (8 * b)) : 0u;               // +1 = 0x01 <-> -1 = 0xff
                    }
                    own[i] ^= flip;
                    nacc += __popc(flip & 0x02020202u);
                }
                if (!PARTIAL) natt += V;
                i8_store<NW>(sp + rowBase + Z0, own);
            }
            if (MODE != 0) {
                // sum_i sigma_i (n - 2 d_i) with d_i the down neighbours that count:  sum(n - 2 d) - 2 sum_{i down}(n - 2 d_i)
                const int n = MODE == 2 ? nl : (cl.lowmode == 1 ? nl : (cl.lowmode == 2 ? cl.nlow : 0));
                int down = 0;
                unsigned sd = 0u, sdd = 0u;
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    const uint32_t dw = i8_down(own[i]);
                    const uint32_t cnt = MODE == 2 || cl.lowmode == 1 ? nd[i] : ndl[i];
                    down += __popc(dw);
                    sd = __dp4a(cnt, 0x01010101u, sd);
                    sdd = __dp4a(cnt, dw, sdd);
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
        double v[4] = {T.S * msum, 0.0, 0.0, beta * T.JS2 * (MODE == 2 ? 0.5 : 1.0) * (double)eInt - hf * T.S * msum};
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
                if (lane == 0 && sum != 0.0) atomicAdd(a.classSums + ((size_t)r * a.nclass + q) * 4 + i, sum);
            }
        }
    }
}

// ---- small companions of the pass: the other kernels that touch the int8 planes ----
// site id <-> storage position are the structured path's (structured.cu); the kernels are instantiated there.

}  // namespace mcg
