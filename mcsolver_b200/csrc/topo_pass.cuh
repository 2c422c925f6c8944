// Topological charge of the structured path: solid angles of the user's triangle circuits (calcSignedArea,
// heisenbergLib.c:114-127; loop :712-716), compiled twice like struct_pass.cuh:
//  * OFFLINE: k_struct_topo_cells in structured.cu (runtime vertex/triangle tables through shared memory);
//  * JIT (NVRTC, -DMCG_JIT_TOPO + generated prologue): vertices, triangles, dims and periods are literals, the vertex
//    table lives in registers with static indices (profiles/r01d: the runtime-table kernel spends 784 instructions per
//    cell, two thirds of them on table lookups, shared-memory staging and address arithmetic).
// Must stay free of host/std includes (NVRTC).
#pragma once
#include "struct_pass.cuh"

namespace mcg {

// One thread per CELL: the ncircuit triangles of a cell share vertices (the four circuits of
// samples/SkyrmionOnHexLattice touch 5 distinct sites), so the distinct vertices are loaded once and the triangles index
// into them.  fp32 engines evaluate the solid angle in fp32 (sum in fp64); fp64 engines keep the reference's double
// arithmetic (parity <= 1e-12).
template <typename T> __device__ __forceinline__ T tri_area(const T (&a)[3], const T (&b)[3], const T (&c)[3], T la, T lb, T lc) {
    T ab = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / la / lb;
    T bc = (b[0] * c[0] + b[1] * c[1] + b[2] * c[2]) / lb / lc;
    T ca = (c[0] * a[0] + c[1] * a[1] + c[2] * a[2]) / lc / la;
    T cx = b[1] * c[2] - b[2] * c[1], cy = b[2] * c[0] - b[0] * c[2], cz = b[0] * c[1] - b[1] * c[0];
    T re = T(1) + ab + bc + ca;
    T im = (a[0] * cx + a[1] * cy + a[2] * cz) / la / lb / lc;
    if (fabs(re) < T(1e-6)) return im > 0 ? T(MCG_REF_PI) : T(-MCG_REF_PI);
    return T(2) * atan(im / re);
}
// same quantity for vertices already normalised to unit length (fp32 engines)
template <typename T> __device__ __forceinline__ T tri_area_unit(const T (&a)[3], const T (&b)[3], const T (&c)[3]) {
    T cx = b[1] * c[2] - b[2] * c[1], cy = b[2] * c[0] - b[0] * c[2], cz = b[0] * c[1] - b[1] * c[0];
    T re = T(1) + (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + (b[0] * c[0] + b[1] * c[1] + b[2] * c[2]) + (c[0] * a[0] + c[1] * a[1] + c[2] * a[2]);
    T im = a[0] * cx + a[1] * cy + a[2] * cz;
    if (fabs(re) < T(1e-6)) return im > 0 ? T(MCG_REF_PI) : T(-MCG_REF_PI);
    if constexpr (sizeof(T) == 4) return T(2) * atanf(__fdividef(im, re));
    else return T(2) * atan(im / re);
}

// ---- two cells at once (fp32): the same solid angles on packed pairs, lane x = cell (X, Y, Z), lane y = cell (X, Y+1, Z) ----
// 2 atan(im/re) per lane, the reference's plain atan (result in (-pi, pi)) with its guard |re| < 1e-6 -> +-PI.  One reciprocal per
// lane serves the division and the range reduction (q = min(|im|,|re|) / max, atan = pi/2 - atan(q) when |im| > |re|); the odd
// polynomial atan(q) = q + q z P(z), z = q^2, is a weighted least-squares fit on [0, 1] with the linear term pinned to 1 (small
// angles keep their relative accuracy): |error| < 1.2e-7, the accuracy of atanf.  13 instructions per solid angle instead of 25.
__device__ __forceinline__ F2 solid_angle2(F2 im, F2 re) {
    const float ax = fabsf(im.x), bx = fabsf(re.x), ay = fabsf(im.y), by = fabsf(re.y);
    const F2 q = F2{__fdividef(fminf(ax, bx), fmaxf(ax, bx)), __fdividef(fminf(ay, by), fmaxf(ay, by))};
    const F2 z = mul2(q, q);
    F2 p = splat2(-4.355608956e-03f);
    p = fma2(p, z, splat2(2.304084672e-02f));
    p = fma2(p, z, splat2(-5.777456935e-02f));
    p = fma2(p, z, splat2(9.794301793e-02f));
    p = fma2(p, z, splat2(-1.397660593e-01f));
    p = fma2(p, z, splat2(1.996270798e-01f));
    p = fma2(p, z, splat2(-3.333165926e-01f));
    const F2 r = fma2(mul2(p, z), q, q);
    float rx = ax > bx ? 1.5707963267948966f - r.x : r.x, ry = ay > by ? 1.5707963267948966f - r.y : r.y;
    rx = __int_as_float(__float_as_int(rx) ^ ((__float_as_int(im.x) ^ __float_as_int(re.x)) & 0x80000000));
    ry = __int_as_float(__float_as_int(ry) ^ ((__float_as_int(im.y) ^ __float_as_int(re.y)) & 0x80000000));
    F2 a = F2{2.f * rx, 2.f * ry};
    if (bx < 1e-6f) a.x = im.x > 0.f ? (float)MCG_REF_PI : -(float)MCG_REF_PI;
    if (by < 1e-6f) a.y = im.y > 0.f ? (float)MCG_REF_PI : -(float)MCG_REF_PI;
    return a;
}
__device__ __forceinline__ F2 dot3_2(const F2 (&a)[3], const F2 (&b)[3]) { return fma2(a[2], b[2], fma2(a[1], b[1], mul2(a[0], b[0]))); }
__device__ __forceinline__ F2 neg2(F2 v) {   // sign flips on the integer pipe: the kernel is bound by the FMA pipe
    return F2{__int_as_float(__float_as_int(v.x) ^ 0x80000000), __int_as_float(__float_as_int(v.y) ^ 0x80000000)};
}
// a.(b x c) with the packed instructions' missing negation supplied by -c:  (b x c)_x = b1 c2 + b2 (-c1), ...
__device__ __forceinline__ F2 det3_2(const F2 (&a)[3], const F2 (&b)[3], const F2 (&c)[3]) {
    const F2 n0 = neg2(c[0]), n1 = neg2(c[1]), n2 = neg2(c[2]);
    const F2 cx = fma2(b[1], c[2], mul2(b[2], n1)), cy = fma2(b[2], c[0], mul2(b[0], n2)), cz = fma2(b[0], c[1], mul2(b[1], n0));
    return fma2(a[2], cz, fma2(a[1], cy, mul2(a[0], cx)));
}

// cells along X handled by one thread of the specialised kernel: the block reduction (two rounds of double-precision
// shuffles, a barrier and one atomic per block) was a fifth of the instructions per cell; it is paid once per TOPO_XPT cells
constexpr int TOPO_XPT = 4;
constexpr int TOPO_YPT = 8;   // cells along Y per thread where X offers fewer than TOPO_XPT

#ifdef MCG_JIT_TOPO
// ---- JIT entry: JT_NV, JT_NT, JT_NPAR, JT_Xd, JT_Yd, JT_Zd, JT_N, jit_real, CtVert<PAR,K>, CtTri<T> from the prologue ----
template <int PAR>
__device__ __forceinline__ double topo_cell(const jit_real *__restrict__ sp, int X, int Y, int Z) {
    jit_real s[JT_NV][3];
    ct_for<0, JT_NV>([&](auto kk) {
        constexpr int K = decltype(kk)::value;
        typedef CtVert<PAR, K> Vt;
        int Xn = X + Vt::cX, Yn = Y + Vt::cY, Zn = Z + Vt::cZ;
        if (Vt::cX != 0 && Xn >= JT_Xd) Xn -= JT_Xd;
        if (Vt::cY != 0 && Yn >= JT_Yd) Yn -= JT_Yd;
        if (Vt::cZ != 0 && Zn >= JT_Zd) Zn -= JT_Zd;
        const jit_real *q = sp + (Vt::base + (Xn * JT_Yd + Yn) * JT_Zd + Zn);
        s[K][0] = q[0]; s[K][1] = q[JT_N]; s[K][2] = q[2 * (size_t)JT_N];
        if (sizeof(jit_real) == 4 && Vt::len != jit_real(1)) {   // fp32: unit vectors once per vertex
            constexpr jit_real inv = jit_real(1) / Vt::len;
            s[K][0] *= inv; s[K][1] *= inv; s[K][2] *= inv;
        }
    });
    double acc = 0.0;
    float accf = 0.f;
    ct_for<0, JT_NT>([&](auto tt) {
        typedef CtTri<decltype(tt)::value> Tr;
        if (sizeof(jit_real) == 4) accf += (float)tri_area_unit<jit_real>(s[Tr::i0], s[Tr::i1], s[Tr::i2]);
        else acc += (double)tri_area<jit_real>(s[Tr::i0], s[Tr::i1], s[Tr::i2], CtVert<PAR, Tr::i0>::len, CtVert<PAR, Tr::i1>::len, CtVert<PAR, Tr::i2>::len);
    });
    return acc + (double)accf;
}
// cells (X, Y, Z) and (X, Y+1, Z) of a lattice whose colouring does not alternate along Y (JT_NPAR == 1), fp32
__device__ __forceinline__ float topo_cell_pair(const float *__restrict__ sp, int X, int Y, int Z) {
    F2 s[JT_NV][3];
    ct_for<0, JT_NV>([&](auto kk) {
        constexpr int K = decltype(kk)::value;
        typedef CtVert<0, K> Vt;
        int Xn = X + Vt::cX, Ya = Y + Vt::cY, Yb = Y + 1 + Vt::cY, Zn = Z + Vt::cZ;
        if (Vt::cX != 0 && Xn >= JT_Xd) Xn -= JT_Xd;
        if (Ya >= JT_Yd) Ya -= JT_Yd;
        if (Yb >= JT_Yd) Yb -= JT_Yd;
        if (Vt::cZ != 0 && Zn >= JT_Zd) Zn -= JT_Zd;
        const float *qa = sp + (Vt::base + (Xn * JT_Yd + Ya) * JT_Zd + Zn), *qb = sp + (Vt::base + (Xn * JT_Yd + Yb) * JT_Zd + Zn);
#pragma unroll
        for (int c = 0; c < 3; c++) s[K][c] = F2{qa[(size_t)c * JT_N], qb[(size_t)c * JT_N]};
        if (Vt::len != 1.f) {
            constexpr float inv = 1.f / (float)Vt::len;
#pragma unroll
            for (int c = 0; c < 3; c++) s[K][c] = mul2(s[K][c], splat2(inv));
        }
    });
    // unit vertices (calcSignedArea heisenbergLib.c:114-127): re = 1 + a.b + b.c + c.a, im = a.(b x c).  The dot product of an edge
    // is written with the lower vertex index first, so that triangles sharing the edge share the instructions
    F2 acc = F2{0.f, 0.f};
    ct_for<0, JT_NT>([&](auto tt) {
        typedef CtTri<decltype(tt)::value> Tr;
        constexpr int A = Tr::i0, B = Tr::i1, C = Tr::i2;
        const F2 dab = dot3_2(s[A < B ? A : B], s[A < B ? B : A]), dbc = dot3_2(s[B < C ? B : C], s[B < C ? C : B]);
        const F2 dca = dot3_2(s[C < A ? C : A], s[C < A ? A : C]);
        const F2 re = add2(add2(dab, dbc), add2(dca, splat2(1.f)));
        acc = add2(acc, solid_angle2(det3_2(s[A], s[B], s[C]), re));
    });
    return acc.x + acc.y;
}

template <int PAR>
__device__ __forceinline__ double topo_case(int par, const jit_real *__restrict__ sp, int X, int Y, int Z) {
    if constexpr (PAR < JT_NPAR) {
        if (par == PAR) return topo_cell<PAR>(sp, X, Y, Z);
        return topo_case<PAR + 1>(par, sp, X, Y, Z);
    } else return 0.0;
}
// Thread (tz, ty) of block (zc + nzc*(yc + nyc*parity), X / TOPO_XPT, r) handles coarse cells (X.., Y0.., zc*TZ + tz) of one sublattice
// parity - TOPO_XPT cells along X and JT_YPT along Y (2D lattices have Xd == 1: the cells per thread then come from Y) - so that every
// vertex lies in a block-uniform class at a literal coarse offset and the block index arithmetic and the reduction (86 of 367
// instructions per cell at one cell per thread, profiles/r02a) are paid once per thread.
extern "C" __global__ void __launch_bounds__(128) mcg_topo(const __grid_constant__ StructArgs a, int nzc, int nyc, double *sums) {
    __shared__ double smem[32];
    const int par = blockIdx.x / (nzc * nyc), rest = blockIdx.x - par * (nzc * nyc);
    const int yc = rest / nzc, zc = rest - yc * nzc;
    const int r = blockIdx.z;
    const int Y0 = (yc * blockDim.y + threadIdx.y) * JT_YPT, Z = zc * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (Z < JT_Zd) {
        const jit_real *sp = (const jit_real *)a.spin + (size_t)r * 3 * JT_N;
        int Y = Y0;
        if constexpr (sizeof(jit_real) == 4 && JT_NPAR == 1 && JT_YPT % 2 == 0) {
            // two cells along Y per trip on packed fp32 pairs (FFMA2): half the arithmetic instructions of the issue-bound kernel
            float accf = 0.f;
#pragma unroll 1
            for (; Y + 1 < JT_Yd && Y + 1 < Y0 + JT_YPT; Y += 2)
#pragma unroll 1
                for (int X = blockIdx.y * TOPO_XPT; X < JT_Xd && X < (int)(blockIdx.y + 1) * TOPO_XPT; X++)
                    accf += topo_cell_pair((const float *)sp, X, Y, Z);
            v[0] += (double)accf;
        }
#pragma unroll 1
        for (; Y < JT_Yd && Y < Y0 + JT_YPT; Y++)
#pragma unroll 1
            for (int X = blockIdx.y * TOPO_XPT; X < JT_Xd && X < (int)(blockIdx.y + 1) * TOPO_XPT; X++) v[0] += topo_case<0>(par, sp, X, Y, Z);
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_AREA, smem);
}
#endif

}  // namespace mcg
