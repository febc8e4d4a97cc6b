/* TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
 *
 * CPU restatement of the reference's D3Q19 BGK collide-and-stream path, written from the
 * algorithm (not translated from the macro-unrolled source) with run-time DIM / stride / nu / U.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this file
 *   (1) bit-for-bit (rho, u, map and the full f lattices) against oracle/_ref, i.e. the reference's
 *       unmodified kernels.cl compiled as host C++ (oracle/ref_shim.cpp), and
 *   (2) against the reference's golden fixtures target_results/8 and target_results/32
 *       (tests/golden/target{8,32}.npz) at the magnitudes recorded in SURVEY.md Appendix A.
 *
 * Every function names the reference lines it follows (paths relative to /root/reference).
 * Build: see oracle/Makefile (-ffp-contract=off, no fast-math: SURVEY F10/F11).
 */
#ifdef _OPENMP
#include <omp.h>
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- cell-type bits: common.h:7-18 ---- */
enum {
    CT_NONE = 0x0, CT_FLUID = 0x1, CT_MOVING = 0x2, CT_CORNER = 0x4, CT_WALL = 0x8,
    CT_LEFT = 0x10, CT_RIGHT = 0x20, CT_TOP = 0x40, CT_BOTTOM = 0x80, CT_FRONT = 0x100, CT_BACK = 0x200
};
#define CT_ANY_SIDE (CT_LEFT | CT_RIGHT | CT_BOTTOM | CT_TOP | CT_BACK | CT_FRONT)

/* ---- D3Q19 velocity set, kernels.cl:131-205; opposite directions, kernels.cl:208-226 ---- */
static const int EX[19] = { 0, 1, 0, -1, 0, 0, 0, 1, -1, -1, 1, 1, 0, -1, 0, 1, 0, -1, 0 };
static const int EY[19] = { 0, 0, 1, 0, -1, 0, 0, 1, 1, -1, -1, 0, 1, 0, -1, 0, 1, 0, -1 };
static const int EZ[19] = { 0, 0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, -1, -1, -1, 1, 1, 1, 1 };
static const int OPP[19] = { 0, 3, 4, 1, 2, 6, 5, 9, 10, 7, 8, 17, 18, 15, 16, 13, 14, 11, 12 };
/* weight class: 0 -> 1/3, 1 -> 1/18, 2 -> 1/36 (kernels.cl:110-128) */
static const int WCLASS[19] = { 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2 };

/* kernels.cl:236-266 (get_cell_type) */
static int cell_type(int x, int y, int z, int dim)
{
    int t = CT_NONE;
    if (x == 1) t |= CT_LEFT;
    if (x == dim - 2) t |= CT_RIGHT;
    if (y == 1) t |= CT_BOTTOM;
    if (y == dim - 2) t |= CT_TOP;
    if (z == 1) t |= CT_BACK;
    if (z == dim - 2) t |= CT_FRONT;
    if (x == 0 || x == dim - 1 || y == 0 || y == dim - 1 || z == 0 || z == dim - 1) t = CT_WALL;
    if (t == (CT_LEFT | CT_BACK | CT_BOTTOM) || t == (CT_RIGHT | CT_BACK | CT_BOTTOM) ||
        t == (CT_LEFT | CT_BACK | CT_TOP) || t == (CT_RIGHT | CT_BACK | CT_TOP))
        t = CT_CORNER;
    if (t == CT_FRONT) t |= CT_MOVING; /* MOVING_BOUNDARY == FRONT, common.h:20 */
    if (t == CT_NONE) t = CT_FLUID;
    return t;
}

/* predicates, common.h:23-66 */
static int t_moving_init(int t) { return t & CT_FRONT; }
static int t_fluid(int t) { return t == CT_FLUID; }
static int t_wall(int t) { return t == CT_WALL; }
static int t_moving(int t) { return t & CT_MOVING; }
static int t_collide(int t) { return t_fluid(t) || t_moving(t); } /* == is_store_macro */
static int t_bounceback(int t) { return (t & CT_ANY_SIDE) && !(t & CT_MOVING); }

static int ilog2(uint64_t v)
{
    int n = 0;
    while (v > 1) { v >>= 1; ++n; }
    return n;
}

/* lbmcl.hpp:131-156: VISCOSITY / VELOCITY reach the kernel as text printed by operator<< at the
 * default precision (6 significant digits, %g) from a T-typed member; the kernel compiler then
 * reads that text back as a T literal. */
static double text_roundtrip_f32(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)(float)v);
    return (double)strtof(buf, NULL);
}
static double text_roundtrip_f64(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", v);
    return strtod(buf, NULL);
}

#define REAL float
#define SUF(name) name##_f32
#define ROUNDTRIP text_roundtrip_f32
#include "lbm_oracle_body.inc"
#undef REAL
#undef SUF
#undef ROUNDTRIP

#define REAL double
#define SUF(name) name##_f64
#define ROUNDTRIP text_roundtrip_f64
#include "lbm_oracle_body.inc"
#undef REAL
#undef SUF
#undef ROUNDTRIP

/* Thread count of the z-parallel loops (launchers such as torchrun export OMP_NUM_THREADS=1). */
int lbm_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* cell-type map only (kernels.cl:290) */
void lbm_oracle_map(int dim, int *map)
{
    for (int z = 0; z < dim; ++z)
        for (int y = 0; y < dim; ++y)
            for (int x = 0; x < dim; ++x)
                map[(size_t)x + (size_t)y * dim + (size_t)z * dim * dim] = cell_type(x, y, z, dim);
}
