// Device-side maths shared by every kernel; free of host/std includes so NVRTC can compile it.
#pragma once
#include "rng.cuh"

namespace mcg {

// heisenbergLib.c:6 - the reference's truncated PI (Q = sum(area)/(4*PI) is off-integer by ~3e-11)
#define MCG_REF_PI 3.1415926535

// indices of the per-sweep raw sums (per replica)
enum { SUM_TOT = 0, SUM_E = 3, SUM_SI = 4, SUM_SJ = 7, SUM_SIJ = 10, SUM_AREA = 11, NSUM = 12 };
// accumulators over measured sweeps (per replica) - mirrors the locals of heisenbergLib.c:624-642
enum {
    ACC_SI = 0, ACC_SJ = 3, ACC_SIJ = 6, ACC_E = 7, ACC_E2 = 8, ACC_M2 = 9, ACC_M4 = 10, ACC_MTMP = 11, ACC_MDOTM = 12,
    ACC_MTOT = 13, ACC_SIZ = 14, ACC_SJZ = 15, ACC_STZ = 16, ACC_SIH = 17, ACC_SJH = 18, ACC_STH = 19, ACC_Q = 20,
    ACC_NMEAS = 21, ACC_STOT = 22, ACC_LASTE = 23, ACC_SIR = 24, ACC_SJR = 27, ACC_SIJR = 30, ACC_ER = 31, ACC_E2R = 32,
    NACC = 40
};
enum { CNT_ATTEMPT = 0, CNT_ACCEPT = 1, CNT_CLUSTER = 2, CNT_WSTEPS = 3, NCNT = 4 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV per-thread doubles over the block and atomically add the block totals to dst[0..NV).
// smem: at least NV*32 doubles.  All threads of the block must call it.
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double *dst, double *smem) {
    const int lt = threadIdx.y * blockDim.x + threadIdx.x;   // 1-D and 2-D blocks
    int lane = lt & 31, w = lt >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = warp_sum(v[i]);
        if (lane == 0) smem[i * 32 + w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double s = lane < nw ? smem[i * 32 + lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0 && s != 0.0) atomicAdd(dst + i, s);
        }
    }
    __syncthreads();
}

// Same reduction without atomics: thread 0 stores the block totals to dst[0..NV) (any memory it alone reads back, or
// shared memory followed by the trailing barrier).  For kernels in which ONE block owns the destination.
template <int NV>
__device__ __forceinline__ void block_reduce_to(double (&v)[NV], double *dst, double *smem) {
    const int lt = threadIdx.y * blockDim.x + threadIdx.x;   // 1-D and 2-D blocks
    int lane = lt & 31, w = lt >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = warp_sum(v[i]);
        if (lane == 0) smem[i * 32 + w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double s = lane < nw ? smem[i * 32 + lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) dst[i] = s;
        }
    }
    __syncthreads();
}

// ---- packed fp32 pairs: Blackwell's FFMA2 / FADD2 / FMUL2 do two independent IEEE fp32 operations per instruction
// (PTX fma/add/mul.rn.f32x2, sm_100+).  The colour pass works on V = 4 consecutive sites per thread, so its per-site
// arithmetic pairs up naturally; each lane rounds exactly like the scalar instruction.
struct __align__(8) F2 { float x, y; };
__device__ __forceinline__ unsigned long long f2_pack(F2 a) {
    unsigned long long u;
    asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(a.x), "f"(a.y));
    return u;
}
__device__ __forceinline__ F2 f2_unpack(unsigned long long u) {
    F2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(u));
    return r;
}
__device__ __forceinline__ F2 splat2(float v) { return F2{v, v}; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(d);
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}

template <typename real> __device__ __forceinline__ real r_sqrt(real x);
template <> __device__ __forceinline__ float r_sqrt<float>(float x) {
    float y;   // one MUFU.SQRT; sqrtf() expands to a guarded Newton sequence with a slow-path call
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <> __device__ __forceinline__ double r_sqrt<double>(double x) { return sqrt(x); }
template <typename real> __device__ __forceinline__ real r_rsqrt(real x);
template <> __device__ __forceinline__ float r_rsqrt<float>(float x) { return rsqrtf(x); }
template <> __device__ __forceinline__ double r_rsqrt<double>(double x) { return 1.0 / sqrt(x); }
// exp(-x) for the Metropolis test.  fp32: ex2.approx (rel. err ~2^-21); fp64: libdevice exp
template <typename real> __device__ __forceinline__ real r_exp(real x);
template <> __device__ __forceinline__ float r_exp<float>(float x) {
    float y;   // FMUL + MUFU.EX2 (flush-to-zero: exp(-dE) below 1e-38 compares as 0 against u >= 2^-24 anyway)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
template <> __device__ __forceinline__ double r_exp<double>(double x) { return exp(x); }
// sin/cos of 2*pi*u, u in (0,1)
template <typename real> __device__ __forceinline__ void r_sincos2pi(real u, real &s, real &c);
template <> __device__ __forceinline__ void r_sincos2pi<float>(float u, float &s, float &c) {
    // argument folded to (-pi,pi) where the MUFU approximations are accurate to ~5e-7 absolute
    __sincosf(6.283185307179586f * (u - 0.5f), &s, &c);
    s = -s; c = -c;
}
// fp64: written out instead of sincospi() (profiles/r02a: 69 of the 309 instructions per attempt of the fp64 pass, two thirds of
// them moves of polynomial coefficients into uniform registers and special-case handling).  2 pi u = (pi/2)(k + r) with
// k = rint(4u) in 0..4 and r in [-1/2, 1/2]; Taylor polynomials of sin/cos(pi r / 2) in r^2 (truncation < 2.3e-16), the
// coefficients in constant memory so that each DFMA reads its own as an operand; the quadrant is a swap and two sign flips.
static __constant__ double MCG_SINC[8] = {1.5707963267948966192, -0.64596409750624625366, 0.079692626246167045121, -0.0046817541353186881007,
                                          0.00016044118478735982187, -3.5988432352120853405e-6, 5.6921729219679268118e-8, -6.6880351098114672325e-10};
static __constant__ double MCG_COSC[9] = {1.0, -1.2337005501361698274, 0.25366950790104801364, -0.020863480763352960873, 0.00091926027483942658024,
                                          -0.000025202042373060605481, 4.7108747788181715037e-7, -6.3866030837918522411e-9, 6.5659631149794723622e-11};
template <> __device__ __forceinline__ void r_sincos2pi<double>(double u, double &s, double &c) {
    const double t = 4.0 * u;
    const double tm = t + 6755399441055744.0;          // 2^52 + 2^51: the low word of the sum is rint(t)
    const int k = __double2loint(tm);
    const double r = t - (tm - 6755399441055744.0);
    const double x = r * r;
    double ps = MCG_SINC[7], pc = MCG_COSC[8];
#pragma unroll
    for (int j = 6; j >= 0; j--) ps = fma(ps, x, MCG_SINC[j]);
#pragma unroll
    for (int j = 7; j >= 0; j--) pc = fma(pc, x, MCG_COSC[j]);
    ps *= r;
    // k: 0 (s, c)   1 (c, -s)   2 (-s, -c)   3 (-c, s)   4 (s, c)
    const double a = (k & 1) ? pc : ps, b = (k & 1) ? ps : pc;
    s = (k & 2) ? -a : a;
    c = ((k + 1) & 2) ? -b : b;
}

// Metropolis test  exp(x) > u  with u = u01<real>(word)  (heisenbergLib.c:461, isingLib.c:244-252; x <= 0 is -dE, x > 0 always accepts).
// fp64: exp() is 47 instructions of the fp64 pass.  The decision is taken from the fp32 approximation whenever that is certain -
// |e32 - u32| exceeds every error of the approximation (ex2.approx 2^-22, argument rounding < 1.1e-5 for |x log2 e| <= 126, and the
// 2^-23 between the 23-bit and the 32-bit uniform of one word) - and from the full-precision exp() in the remaining ~2e-5 of the
// attempts: the outcome is the one the fp64 comparison gives, always.
template <typename real> __device__ __forceinline__ bool metro_accept(real x, uint32_t word);
template <> __device__ __forceinline__ bool metro_accept<float>(float x, uint32_t word) { return r_exp<float>(x) > u01<float>(word); }
template <> __device__ __forceinline__ bool metro_accept<double>(double x, uint32_t word) {
    if (x >= 0.0) return true;
    const float e = r_exp<float>((float)x);
    const float d = e - u01<float>(word);
    if (fabsf(d) > fmaf(e, 2e-5f, 3e-7f)) return d > 0.f;
    return exp(x) > u01<double>(word);
}

// proposal direction from two Philox words: uniform on S^2 (NC=3) / S^1 (NC=2)
template <int NC, typename real>
__device__ __forceinline__ void random_dir(uint32_t w0, uint32_t w1, real (&n)[3]) {
    if (NC == 3) {
        real z = real(2) * u01<real>(w0) - real(1);
        real sn, cs;
        r_sincos2pi<real>(u01<real>(w1), sn, cs);
        // |z| < 1 strictly: u01 is (k+0.5)/2^23 (fp32) or (r+0.5)/2^32 (fp64), so 1 - z*z >= 2^-22 > 0, no clamp needed
        real rr = r_sqrt<real>(real(1) - z * z);
        n[0] = rr * cs; n[1] = rr * sn; n[2] = z;
    } else {
        real sn, cs;
        r_sincos2pi<real>(u01<real>(w0), sn, cs);
        n[0] = cs; n[1] = sn; n[2] = real(0);
    }
}

// local field of one link: H += J . s_nb   (J flat order xx,yy,zz,xy,xz,yz,yx,zx,zy; XY uses 0,1,3,6)
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ void add_field(real (&H)[3], const real *__restrict__ J, const real (&t)[3]) {
    if (NC == 1) {
        H[0] += J[0] * t[0];
    } else if (NC == 2) {
        if (FULLJ) {
            H[0] += J[0] * t[0] + J[3] * t[1];
            H[1] += J[6] * t[0] + J[1] * t[1];
        } else {
            H[0] += J[0] * t[0];
            H[1] += J[1] * t[1];
        }
    } else {
        if (FULLJ) {
            H[0] += J[0] * t[0] + J[3] * t[1] + J[4] * t[2];
            H[1] += J[6] * t[0] + J[1] * t[1] + J[5] * t[2];
            H[2] += J[7] * t[0] + J[8] * t[1] + J[2] * t[2];
        } else {
            H[0] += J[0] * t[0];
            H[1] += J[1] * t[1];
            H[2] += J[2] * t[2];
        }
    }
}

// on-site energy  sum_a D_a s_a^2 - h s_axis   (getOnsiteEnergy heisenbergLib.c:249-253, xyLib.c:200-204)
template <int NC, typename real>
__device__ __forceinline__ real onsite_energy(const real (&s)[3], const real *__restrict__ D, real beta, real hf) {
    if (NC == 1) return -hf * s[0];
    real e = D[0] * s[0] * s[0] + D[1] * s[1] * s[1];
    if (NC == 3) e += D[2] * s[2] * s[2];
    return beta * e - hf * (NC == 3 ? s[2] : s[0]);
}

// n.J.n for the Wolff bond weight (heisenbergLib.c:355: diagonalDot(ref,ref,J))
template <int NC, typename real, bool FULLJ>
__device__ __forceinline__ real quad_form(const real *__restrict__ J, const real (&a)[3], const real (&b)[3]) {
    real H[3] = {0, 0, 0};
    add_field<NC, real, FULLJ>(H, J, b);
    real r = a[0] * H[0];
    if (NC >= 2) r += a[1] * H[1];
    if (NC == 3) r += a[2] * H[2];
    return r;
}

}  // namespace mcg
