// D3Q19 lattice description, cell classification, CSoA addressing and the arithmetic policy shared
// by all kernels of the collide-and-stream path.
//
// Semantics follow the reference (paths relative to the reference tree):
//   direction numbering / weights / opposites   kernels.cl:110-226
//   cell-type bits and predicates               common.h:7-66, kernels.cl:236-266
//   CSoA(stride) index                          kernels.cl:64
// The code itself is written for sm_100a: compile-time direction tables so that the 19 x VEC
// populations of a thread stay in registers, and an arithmetic policy that either reproduces the
// reference's IEEE operation order exactly (strict) or lets ptxas contract (fast, the -o switch).
#pragma once

// No standard-library headers: this file and lbm_kernels.cuh are also compiled at run time by NVRTC
// (lbm_nvrtc.cu, the -D specialised variant), which has none.
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace lbm {

constexpr int Q = 19;

// ---- cell-type bits (common.h:7-18) ----
enum : int {
    CT_NONE = 0x000,
    CT_FLUID = 0x001,
    CT_MOVING = 0x002,
    CT_CORNER = 0x004,
    CT_WALL = 0x008,
    CT_LEFT = 0x010,    // x == 1
    CT_RIGHT = 0x020,   // x == DIM-2
    CT_TOP = 0x040,     // y == DIM-2
    CT_BOTTOM = 0x080,  // y == 1
    CT_FRONT = 0x100,   // z == DIM-2   (the moving lid plane, common.h:20)
    CT_BACK = 0x200     // z == 1
};

// ---- compile-time direction tables (kernels.cl:131-226) ----
__host__ __device__ constexpr int ex(int q)
{
    return (q == 1 || q == 7 || q == 10 || q == 11 || q == 15) ? 1
         : (q == 3 || q == 8 || q == 9 || q == 13 || q == 17) ? -1 : 0;
}
__host__ __device__ constexpr int ey(int q)
{
    return (q == 2 || q == 7 || q == 8 || q == 12 || q == 16) ? 1
         : (q == 4 || q == 9 || q == 10 || q == 14 || q == 18) ? -1 : 0;
}
__host__ __device__ constexpr int ez(int q)
{
    return (q == 6 || q == 15 || q == 16 || q == 17 || q == 18) ? 1
         : (q == 5 || q == 11 || q == 12 || q == 13 || q == 14) ? -1 : 0;
}
// opposite direction: e(opp(q)) == -e(q)
__host__ __device__ constexpr int opp(int q)
{
    constexpr int t[Q] = { 0, 3, 4, 1, 2, 6, 5, 9, 10, 7, 8, 17, 18, 15, 16, 13, 14, 11, 12 };
    return t[q];
}
// weight class: 0 -> 1/3, 1 -> 1/18, 2 -> 1/36  (kernels.cl:110-128)
__host__ __device__ constexpr int wclass(int q) { return q == 0 ? 0 : (q <= 6 ? 1 : 2); }

static_assert(opp(1) == 3 && opp(3) == 1 && opp(2) == 4 && opp(4) == 2, "opp");
static_assert(opp(5) == 6 && opp(6) == 5, "opp");
static_assert(opp(7) == 9 && opp(9) == 7 && opp(8) == 10 && opp(10) == 8, "opp");
static_assert(opp(11) == 17 && opp(17) == 11 && opp(12) == 18 && opp(18) == 12, "opp");
static_assert(opp(13) == 15 && opp(15) == 13 && opp(14) == 16 && opp(16) == 14, "opp");

// compile-time loop: f(IntC<0>{}) ... f(IntC<N-1>{}), fully unrolled; `decltype(arg)::value` is the index
template <int V>
struct IntC {
    static constexpr int value = V;
};
template <int I, int N, typename F>
__device__ __forceinline__ void static_for_step(F &f)
{
    if constexpr (I < N) {
        f(IntC<I>{});
        static_for_step<I + 1, N>(f);
    }
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_step<0, N>(f);
}

// ---- cell classification from coordinates (kernels.cl:236-266); 0 B of memory traffic ----
// Bits that depend on (y, z) only; CT_WALL if the whole row is wall.
__host__ __device__ inline int row_bits(int y, int z, int dim)
{
    int t = CT_NONE;
    if (y == 1) t |= CT_BOTTOM;
    if (y == dim - 2) t |= CT_TOP;
    if (z == 1) t |= CT_BACK;
    if (z == dim - 2) t |= CT_FRONT;
    if (y == 0 || y == dim - 1 || z == 0 || z == dim - 1) t = CT_WALL;
    return t;
}
// Full cell type given the row bits.
__host__ __device__ inline int cell_type_from_row(int rowbits, int x, int dim)
{
    int t = rowbits;
    if (t != CT_WALL) {
        if (x == 1) t |= CT_LEFT;
        if (x == dim - 2) t |= CT_RIGHT;
        if (x == 0 || x == dim - 1) t = CT_WALL;
    }
    if (t == (CT_LEFT | CT_BACK | CT_BOTTOM) || t == (CT_RIGHT | CT_BACK | CT_BOTTOM) ||
        t == (CT_LEFT | CT_BACK | CT_TOP) || t == (CT_RIGHT | CT_BACK | CT_TOP))
        t = CT_CORNER;
    if (t == CT_FRONT) t |= CT_MOVING;
    if (t == CT_NONE) t = CT_FLUID;
    return t;
}
__host__ __device__ inline int cell_type(int x, int y, int z, int dim)
{
    return cell_type_from_row(row_bits(y, z, dim), x, dim);
}
// common.h:23-66
__host__ __device__ inline bool is_collision(int t) { return t == CT_FLUID || (t & CT_MOVING); }
__host__ __device__ inline bool is_bounceback(int t)
{
    return (t & (CT_LEFT | CT_RIGHT | CT_BOTTOM | CT_TOP | CT_BACK | CT_FRONT)) && !(t & CT_MOVING);
}
// "is_moving_init": the cell starts with u = (U, 0, 0) (common.h:23-26; kernels.cl:292-295).
// True on the whole z == DIM-2 plane except its wall ring.
__host__ __device__ inline bool has_front_bit(int x, int y, int z, int dim)
{
    return z == dim - 2 && x >= 1 && x <= dim - 2 && y >= 1 && y <= dim - 2;
}

// ---- CSoA(stride) addressing over a slab's local cell ids (kernels.cl:64) ----
// Cells are grouped in blocks of `stride` consecutive ids; inside a block the 19 populations are 19
// contiguous runs of `stride` values:  index(id, q) = index(id, 0) + q * stride.
struct Layout {
    int sdiv;          // log2(stride)
    long long smod;    // stride - 1
    __host__ __device__ __forceinline__ long long base(long long id) const
    {
        return (((id >> sdiv) * Q) << sdiv) + (id & smod);
    }
    __host__ __device__ __forceinline__ long long qpitch() const { return smod + 1; }
};

// ---- arithmetic policy ----
// Strict: every operation is an individually rounded IEEE operation in the reference's order
// (the *_rn intrinsics are never contracted into FMAs).  Fast: plain operators (ptxas contracts)
// and, in fp32, the approximate division — the analogue of -cl-fast-relaxed-math.
template <typename T, bool FAST>
struct Arith;

template <>
struct Arith<float, false> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <>
struct Arith<double, false> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <>
struct Arith<float, true> {
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
};
template <>
struct Arith<double, true> {
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
};

// Constants the reference compiles into the kernel as folded literals (kernels.cl:56-62, 110-128),
// evaluated once in T arithmetic on the host and passed by value.
template <typename T>
struct Consts {
    T u_lid;     // VELOCITY literal
    T inv_tau;   // 1 / (3*VISCOSITY + 0.5)
    T w[3];      // 1/3, 1/18, 1/36
};

// e_q . u with the products by 0 / +-1 folded away (they are exact), leaving the reference's single
// rounded addition:  (ux*Ex + uy*Ey) + uz*Ez,  kernels.cl:415.
template <typename A, int q, typename T>
__device__ __forceinline__ T e_dot_u(T ux, T uy, T uz)
{
    constexpr int X = ex(q), Y = ey(q), Z = ez(q);
    T s = T(0);
    if constexpr (X != 0 && Y != 0) s = A::add(X > 0 ? ux : -ux, Y > 0 ? uy : -uy);
    else if constexpr (X != 0) s = X > 0 ? ux : -ux;
    else if constexpr (Y != 0) s = Y > 0 ? uy : -uy;
    if constexpr (Z != 0) {
        if constexpr (X != 0 || Y != 0) s = A::add(s, Z > 0 ? uz : -uz);
        else s = Z > 0 ? uz : -uz;
    }
    return s;
}

// (1 + 3*eu + 4.5*eu*eu - 1.5*u2) evaluated as ((1 + 3eu) + (4.5eu)eu) - c15u2, kernels.cl:416
template <typename A, typename T>
__device__ __forceinline__ T eq_poly(T eu, T c15u2)
{
    return A::sub(A::add(A::add(T(1), A::mul(T(3), eu)), A::mul(A::mul(T(4.5), eu), eu)), c15u2);
}

}  // namespace lbm
