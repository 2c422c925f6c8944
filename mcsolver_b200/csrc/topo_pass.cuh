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
#pragma unroll 1
        for (int Y = Y0; Y < JT_Yd && Y < Y0 + JT_YPT; Y++)
#pragma unroll 1
            for (int X = blockIdx.y * TOPO_XPT; X < JT_Xd && X < (int)(blockIdx.y + 1) * TOPO_XPT; X++) v[0] += topo_case<0>(par, sp, X, Y, Z);
    }
    block_accumulate<1>(v, sums + (size_t)r * NSUM + SUM_AREA, smem);
}
#endif

}  // namespace mcg
