// Counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) and the engine's counter layout.
//
// Replaces the reference's process-global, never-seeded libc rand() (heisenbergLib.c:96-98, 443-445,
// 461).  One Philox call = the four 32-bit words one Metropolis attempt needs:
//   w0,w1 -> proposal direction, w2 -> acceptance uniform, w3 -> partial-sweep attempt mask.
// Counter = (site id in the REFERENCE numbering, sweep_lo, sweep_hi16 | sub<<16 | stream<<24, replica),
// key = 64-bit user seed.  Because the counter holds the reference site id (not the colour-major
// storage position) the stream of a site is independent of the colouring, the memory layout and
// the number of GPUs.  oracle/oracle.c (rng4) restates exactly this layout for the checker.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#else
#include <cstdint>
#endif

namespace mcg {

enum : uint32_t { STREAM_METRO = 0, STREAM_INIT = 1, STREAM_WBOND = 2, STREAM_WSEED = 3, STREAM_PT = 4 };

// The ten Philox round keys (k + r*W) of the 64-bit seed, computed once on the host: the kernels read
// them straight from the constant bank (kernel parameter) as LOP3 operands - zero instructions per
// attempt instead of 20 integer adds (profiles/r01).
struct RngKey {
    uint32_t rk[10][2];
};
__host__ __device__ __forceinline__ RngKey make_rng_key(uint64_t seed) {
    RngKey k;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { k.rk[r][0] = k0; k.rk[r][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return k;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RngKey &key,
                                                       uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        // one IMAD.WIDE.U32 per product on the device
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ key.rk[r][0], n2 = hi0 ^ c3 ^ key.rk[r][1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ void rng4(const RngKey &key, uint32_t replica, uint32_t stream, uint32_t sub, uint64_t sweep,
                                              uint32_t site, uint32_t (&out)[4]) {
    philox4x32_10(site, (uint32_t)sweep, (uint32_t)((sweep >> 32) & 0xFFFFu) | (sub << 16) | (stream << 24), replica,
                  key, out);
}

// Ising needs ONE uniform per attempt, a Philox block holds four: four consecutive reference site ids share the block
// whose first counter word is id >> 2 (sub-stream 0) and use word id & 3 of it; the attempt-probability uniform of
// partial sweeps (ninterval < N) comes from sub-stream 1 in the same way.  The mapping depends on the reference site id
// only, so it is the same on every path; a caller visiting several sites keeps the last block (the V = 4 item of the
// structured pass needs two Philox calls instead of four).
__host__ __device__ __forceinline__ uint32_t pick4(const uint32_t (&w)[4], uint32_t i) {
    const uint32_t lo = (i & 1u) ? w[1] : w[0], hi = (i & 1u) ? w[3] : w[2];
    return (i & 2u) ? hi : lo;
}
struct IsingWords {
    uint32_t blk = 0xffffffffu;
    uint32_t wa[4], wp[4];
    __host__ __device__ __forceinline__ void get(const RngKey &key, uint32_t replica, uint64_t sweep, uint32_t id, bool partial,
                                                 uint32_t &wAcc, uint32_t &wAtt) {
        if ((id >> 2) != blk) {
            blk = id >> 2;
            rng4(key, replica, STREAM_METRO, 0, sweep, blk, wa);
            if (partial) rng4(key, replica, STREAM_METRO, 1, sweep, blk, wp);
        }
        wAcc = pick4(wa, id & 3u);
        wAtt = partial ? pick4(wp, id & 3u) : 0u;
    }
};

// ---- grouped streams of the vectorised structured pass ----
// One thread of that pass updates an ITEM of V sites whose reference ids are id0 + v*S (S = id stride along the vector
// axis).  Per-site Philox blocks waste words (an O(3) attempt uses 3 of 4, O(2) 2, Ising 1), so the item draws its words
// from shared blocks instead: group index G = (id / (V*S))*S + id % S is the first counter word, member m = (id / S) % V,
// the t-th word of member m is word k = W*m + t of the group (W = words per attempt = number of spin components), i.e.
// word k & 3 of the block drawn with sub-stream k >> 2.  V = 4: 3 Philox calls per item instead of 4 (O(3)), 2 (O(2)),
// 1 (Ising).  The attempt-probability uniform of partial sweeps is word m & 3 of the block with sub-stream 7 + (m >> 2)
// (items of more than four sites: the int8 Ising pass, V = 16, W = 1 - its acceptance words use sub-streams 0..3).
// Still a pure function of (site id, S, V) - the oracle restates it (oracle.c: grouped_words).
template <int W, int V> struct ItemWords {
    static constexpr int NCALL = (W * V + 3) / 4;
    static constexpr int NPART = (V + 3) / 4;
    uint32_t c[NCALL][4];
    uint32_t p[NPART][4];
    uint32_t G;
    int have;   // blocks drawn so far (a compile-time constant after unrolling)
    __host__ __device__ __forceinline__ void begin(const RngKey &key, uint32_t replica, uint64_t sweep, uint32_t id0, uint32_t S, bool partial) {
        G = (id0 / (V * S)) * S + id0 % S;
        have = 0;
        if (partial) {
#pragma unroll
            for (int i = 0; i < NPART; i++) rng4(key, replica, STREAM_METRO, 7u + (uint32_t)i, sweep, G, p[i]);
        }
    }
    // draw the blocks that hold the words of members [0, mEnd): called right before those members are processed, so that
    // at most the words of the members in flight are live (drawing all blocks up front spills the 64-register pass kernel)
    __host__ __device__ __forceinline__ void need(const RngKey &key, uint32_t replica, uint64_t sweep, int mEnd) {
        const int last = (W * mEnd - 1) >> 2;
#pragma unroll
        for (int i = 0; i < NCALL; i++)
            if (i >= have && i <= last) rng4(key, replica, STREAM_METRO, (uint32_t)i, sweep, G, c[i]);
        if (last + 1 > have) have = last + 1;
    }
    // m and t are compile-time constants in the unrolled callers: static register indices
    __host__ __device__ __forceinline__ uint32_t word(int m, int t) const { return c[(W * m + t) >> 2][(W * m + t) & 3]; }
    // the words of member m in the slots the per-site code uses: [0],[1] direction, [2] acceptance, [3] attempt probability
    __host__ __device__ __forceinline__ void lane(int m, bool partial, uint32_t (&w)[4]) const {
        if (W == 3) { w[0] = word(m, 0); w[1] = word(m, 1); w[2] = word(m, 2); }
        else if (W == 2) { w[0] = word(m, 0); w[1] = 0u; w[2] = word(m, 1); }
        else { w[0] = 0u; w[1] = 0u; w[2] = word(m, 0); }
        w[3] = partial ? p[m >> 2][m & 3] : 0u;
    }
};

// uniforms strictly inside (0,1): fp64 uses all 32 bits: (r+0.5)/2^32; fp32 the top 23 bits:
// (k+0.5)/2^23 with k = r>>9, built without an int->float conversion: 1.mantissa - (1 - 2^-24), exact.
template <typename real> __host__ __device__ __forceinline__ real u01(uint32_t r);
template <> __host__ __device__ __forceinline__ double u01<double>(uint32_t r) {
    return ((double)r + 0.5) * (1.0 / 4294967296.0);
}
template <> __host__ __device__ __forceinline__ float u01<float>(uint32_t r) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(0x3f800000u | (r >> 9)) - 0.99999994f;   // 0.99999994f == 1 - 2^-24 exactly
#else
    return ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f);
#endif
}

}  // namespace mcg
