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
#include <cstdint>

namespace mcg {

enum : uint32_t { STREAM_METRO = 0, STREAM_INIT = 1, STREAM_WBOND = 2, STREAM_WSEED = 3, STREAM_PT = 4 };

struct RngKey {
    uint32_t k0, k1;   // seed
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;   // uniform across the grid: folded into immediates / uniform registers
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ void rng4(RngKey key, uint32_t replica, uint32_t stream, uint32_t sub, uint64_t sweep,
                                              uint32_t site, uint32_t (&out)[4]) {
    philox4x32_10(site, (uint32_t)sweep, (uint32_t)((sweep >> 32) & 0xFFFFu) | (sub << 16) | (stream << 24), replica,
                  key.k0, key.k1, out);
}

// uniforms strictly inside (0,1): fp64 uses all 32 bits, fp32 the top 24 (exactly representable)
template <typename real> __host__ __device__ __forceinline__ real u01(uint32_t r);
template <> __host__ __device__ __forceinline__ double u01<double>(uint32_t r) {
    return ((double)r + 0.5) * (1.0 / 4294967296.0);
}
template <> __host__ __device__ __forceinline__ float u01<float>(uint32_t r) {
    return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

}  // namespace mcg
