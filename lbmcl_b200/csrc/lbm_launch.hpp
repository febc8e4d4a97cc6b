// Launch interface between the C ABI (lbm_capi.cu, host logic only) and the translation units that
// instantiate the kernels.  The kernels are split over several .cu files so that the library builds in
// parallel (`make -j`): one object per (precision, cells-per-thread) of the pull kernel, one for the
// in-place AA kernels, one for the TMA-fed kernels, one for everything that runs once (initialize, map,
// dump views, dense halos, slab flag helpers).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "lbm_kernels.cuh"

namespace lbm {

struct LaunchCfg {
    dim3 block;       // CUDA block of the step kernels
    int dim;
    int lm;           // LM_* addressing mode
    bool fast;        // -o
};

// grid of a pull / AA launch over `n_planes` planes
inline dim3 step_grid(const LaunchCfg &k, int vec, int n_planes)
{
    return dim3((unsigned)(k.dim / ((int)k.block.x * vec)), (unsigned)(k.dim / (int)k.block.y),
                (unsigned)((n_planes + (int)k.block.z - 1) / (int)k.block.z));
}
template <typename T>
inline int planes_of(const StepArgs<T> &a)
{
    return a.zmap_n + (a.z_end > a.z_begin ? a.z_end - a.z_begin : 0);
}

// two-lattice pull kernel, VEC cells per thread (lbm_launch_pull.cu, compiled once per (T, VEC))
cudaError_t launch_pull_f32_v1(const LaunchCfg &, const StepArgs<float> &, bool macro, int peer, cudaStream_t);
cudaError_t launch_pull_f32_v2(const LaunchCfg &, const StepArgs<float> &, bool macro, int peer, cudaStream_t);
cudaError_t launch_pull_f32_v4(const LaunchCfg &, const StepArgs<float> &, bool macro, int peer, cudaStream_t);
cudaError_t launch_pull_f64_v1(const LaunchCfg &, const StepArgs<double> &, bool macro, int peer, cudaStream_t);
cudaError_t launch_pull_f64_v2(const LaunchCfg &, const StepArgs<double> &, bool macro, int peer, cudaStream_t);

// in-place AA kernels (lbm_launch_aa.cu); shift = SHIFT step (even iterations), else LOCAL
cudaError_t launch_aa_f32(const LaunchCfg &, const StepArgs<float> &, bool macro, bool shift, cudaStream_t);
cudaError_t launch_aa_f64(const LaunchCfg &, const StepArgs<double> &, bool macro, bool shift, cudaStream_t);

// TMA-fed kernels (lbm_launch_tma.cu)
struct TmaCfg {
    const CUtensorMap *map_src;   // load map of the lattice read
    const CUtensorMap *map_dst;   // store map of the lattice written
    int tx;            // cells per row tile = consumer threads per CTA (the CTA has tx + 32 threads)
    int grid;          // persistent CTAs
    size_t smem;       // dynamic shared memory per CTA
    int osdiv;         // log2(min(stride, 32)): layout of a warp's output tile
    int *error;        // device flag: an mbarrier wait ran into its limit
    bool fast;
};
bool tma_direct_store();              // results leave through plain stores (else: per-warp bulk-tensor stores)
cudaError_t tma_prepare(int device);  // once per device: opt in to > 48 KB dynamic shared memory
int tma_resident_ctas(bool f64, int tx, size_t smem, bool fast);  // CTAs per SM by registers and shared memory
cudaError_t launch_tma_f32(const TmaCfg &, const StepArgs<float> &, int ns, bool macro, cudaStream_t);
cudaError_t launch_tma_f64(const TmaCfg &, const StepArgs<double> &, int ns, bool macro, cudaStream_t);

// run-once kernels (lbm_launch_misc.cu)
cudaError_t launch_init_f32(const InitArgs<float> &, bool aa, cudaStream_t);
cudaError_t launch_init_f64(const InitArgs<double> &, bool aa, cudaStream_t);
cudaError_t launch_stale_f32(const Consts<float> &, float *out, cudaStream_t);
cudaError_t launch_stale_f64(const Consts<double> &, double *out, cudaStream_t);
cudaError_t launch_map(int *map, int dim, cudaStream_t);
struct ViewCfg {
    int dim, zs0, nz_local, z_begin, z_end;
    Layout lay_local, lay_global;
    int pristine;
    int aa_swapped;   // AA only
};
cudaError_t launch_view_f32(const float *g, float *out, const ViewCfg &, const Consts<float> &, bool aa, cudaStream_t);
cudaError_t launch_view_f64(const double *g, double *out, const ViewCfg &, const Consts<double> &, bool aa, cudaStream_t);
cudaError_t launch_halo_f32(float *lattice, float *dense, int dim, long long plane_local, Layout, int dir_up, bool pack,
                            cudaStream_t);
cudaError_t launch_halo_f64(double *lattice, double *dense, int dim, long long plane_local, Layout, int dir_up, bool pack,
                            cudaStream_t);
cudaError_t launch_slab_wait(const SlabSync &, bool has_lo, bool has_hi, cudaStream_t);
cudaError_t launch_slab_signal(const SlabSync &, bool has_lo, bool has_hi, cudaStream_t);

}  // namespace lbm
