// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Builds the reference's OWN kernel source (kernels.cl + common.h, unmodified, included from
// $LBMCL_REF where it lies; nothing is copied into this repository) as host C++ so that the
// reference's arithmetic can be executed on a CPU.  SURVEY.md F8 / §8(c): the reference cannot
// be built as shipped (no OpenCL headers/ICD in this image), but its device code compiles as C++
// once the OpenCL address-space qualifiers are defined away.
//
// The driver below restates the host schedule of lbmcl.hpp:435-467 (ping-pong binding,
// update_macro flag) and lbmcl.hpp:490-521 (initialize, then one compute launch per iteration).
//
// Compile-time configuration is exactly what lbmcl.hpp:131-156 (kernelOptionsStr) would emit:
//   -DDIM= -DLWS= -DSTRIDE_DIV= -DSTRIDE_MOD= -DVISCOSITY= -DVELOCITY= -DFP_SINGLE|-DFP_DOUBLE
// fp32 additionally needs -fsingle-precision-constant (the host analogue of
// -cl-single-precision-constant, lbmcl.hpp:146).  Never build with -ffast-math (F11).
#include <cmath>
#include <cstddef>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#define __kernel
#define __global
#define __local static
#define restrict __restrict__
#define CLK_LOCAL_MEM_FENCE 0
static inline void barrier(int) {}

static thread_local int g_gid[3];
static inline int get_global_id(int d) { return g_gid[d]; }
static inline int get_local_id(int d) { return g_gid[d] % LWS; }

#include "kernels.cl"

extern "C" {

int ref_dim(void) { return DIM; }
int ref_stride(void) { return STRIDE_MOD + 1; }
int ref_sizeof_real(void) { return (int)sizeof(real_t); }
double ref_viscosity(void) { return (double)(real_t)VISCOSITY; }
double ref_velocity(void) { return (double)(real_t)VELOCITY; }

// Thread count of the z-parallel loops below.  Launchers such as torchrun export OMP_NUM_THREADS=1; the
// CPU-baseline legs of bench.py ask for the cores explicitly and report what they really got.
int ref_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// kernels.cl:277-318 over the NDRange (DIM,DIM,DIM) of lbmcl.hpp:371,495
void ref_initialize(real_t *f_stream, real_t *f_collide, real_t *density, real_t *u, int *map)
{
#pragma omp parallel for schedule(static)
    for (int z = 0; z < DIM; ++z)
        for (int y = 0; y < DIM; ++y)
            for (int x = 0; x < DIM; ++x) {
                g_gid[0] = x; g_gid[1] = y; g_gid[2] = z;
                initialize(f_stream, f_collide, density, u, map);
            }
}

// kernels.cl:321-425 over the same NDRange (one launch, lbmcl.hpp:505-511).  The push scheme has
// exactly one writer per (cell,q) slot, so the z-parallel loop is race free.
void ref_compute(real_t *dst, const real_t *src, real_t *density, real_t *u, const int *map,
                 int update_macro)
{
#pragma omp parallel for schedule(static)
    for (int z = 0; z < DIM; ++z)
        for (int y = 0; y < DIM; ++y)
            for (int x = 0; x < DIM; ++x) {
                g_gid[0] = x; g_gid[1] = y; g_gid[2] = z;
                compute(dst, src, density, u, map, update_macro);
            }
}

// lbmcl.hpp:435-467 + 490-521: iteration `it` (1-based) reads f_collide when it is odd and f_stream
// when it is even; update_macro = every != 0 && it % every == 0.  Snapshots of rho/u are taken at
// it = 0 and after every flagged iteration (what storeData() would write), packed one after the
// other into snap_rho[n_snap][N] and snap_u[n_snap][3N]; either may be NULL.  Returns the number of
// snapshots written.
int ref_run(real_t *f_stream, real_t *f_collide, real_t *density, real_t *u, int *map,
            int iterations, int every, real_t *snap_rho, real_t *snap_u)
{
    const size_t n = (size_t)DIM * DIM * DIM;
    int n_snap = 0;
    ref_initialize(f_stream, f_collide, density, u, map);
    if (every != 0) {
        if (snap_rho) std::memcpy(snap_rho + n_snap * n, density, n * sizeof(real_t));
        if (snap_u) std::memcpy(snap_u + n_snap * 3 * n, u, 3 * n * sizeof(real_t));
        ++n_snap;
    }
    for (int it = 1; it <= iterations; ++it) {
        const int flag = (every != 0 && it % every == 0) ? 1 : 0;
        const bool swap = (it % 2 == 0);
        ref_compute(swap ? f_collide : f_stream, swap ? f_stream : f_collide, density, u, map, flag);
        if (flag) {
            if (snap_rho) std::memcpy(snap_rho + n_snap * n, density, n * sizeof(real_t));
            if (snap_u) std::memcpy(snap_u + n_snap * 3 * n, u, 3 * n * sizeof(real_t));
            ++n_snap;
        }
    }
    return n_snap;
}

} // extern "C"
