// Instantiations of the TMA-fed step kernels (lbm_tma.cuh) and their launch glue.
#include "lbm_launch.hpp"
#include "lbm_tma.cuh"

namespace lbm {

namespace {

constexpr int TMA_SMEM_MAX = 200 * 1024;

template <typename T>
cudaError_t prepare_t()
{
    cudaError_t e = cudaSuccess;
#define LBM_TMA_ATTR(F, M)                                                                               \
    if (e == cudaSuccess)                                                                                \
        e = cudaFuncSetAttribute(step_tma_kernel<T, F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                 TMA_SMEM_MAX)
    LBM_TMA_ATTR(false, false);
    LBM_TMA_ATTR(false, true);
    LBM_TMA_ATTR(true, false);
    LBM_TMA_ATTR(true, true);
#undef LBM_TMA_ATTR
    return e;
}

template <typename T>
void go(const TmaCfg &k, const TmaArgs<T> &a, bool macro, int grid, cudaStream_t s)
{
    const CUtensorMap &ms = *k.map_src, &md = *k.map_dst;
    const int tx = k.tx + 32;  // consumers + the producer warp
    if (k.fast) {
        if (macro) step_tma_kernel<T, true, true><<<grid, tx, k.smem, s>>>(ms, md, a, k.error);
        else step_tma_kernel<T, true, false><<<grid, tx, k.smem, s>>>(ms, md, a, k.error);
    } else {
        if (macro) step_tma_kernel<T, false, true><<<grid, tx, k.smem, s>>>(ms, md, a, k.error);
        else step_tma_kernel<T, false, false><<<grid, tx, k.smem, s>>>(ms, md, a, k.error);
    }
}

template <typename T>
cudaError_t launch_tma(const TmaCfg &k, const StepArgs<T> &sa, int ns, bool macro, cudaStream_t s)
{
    // live planes of this launch (the TMA kernels take one contiguous range)
    const int zf = sa.z_begin < 1 ? 1 : sa.z_begin;
    const int zl = sa.z_end > sa.dim - 1 ? sa.dim - 1 : sa.z_end;
    if (zl <= zf) return cudaSuccess;
    TmaArgs<T> a{};
    a.src = sa.src;
    a.dst = sa.dst;
    a.rho = sa.rho;
    a.u = sa.u;
    a.dim = sa.dim;
    a.zs0 = sa.zs0;
    a.z_first = zf;
    a.n_xseg = sa.dim / k.tx;
    a.n_tiles = (zl - zf) * (sa.dim - 2) * a.n_xseg;
    a.ns = ns;
    a.osdiv = k.osdiv;
    a.n_local = sa.n_local;
    a.lay = sa.lay;
    a.c = sa.c;
    for (int i = 0; i < 2; ++i)
        for (int q = 0; q < Q; ++q) a.stale[i][q] = sa.stale[i][q];
    const int grid = a.n_tiles < k.grid ? a.n_tiles : k.grid;
    go<T>(k, a, macro, grid, s);
    return cudaGetLastError();
}

template <typename T>
int resident_ctas(int tx, size_t smem, bool fast)
{
    int n = 0, m = 0;
    cudaError_t e;
    if (fast) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, step_tma_kernel<T, true, false>, tx + 32, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, step_tma_kernel<T, true, true>, tx + 32, smem);
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, step_tma_kernel<T, false, false>, tx + 32, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, step_tma_kernel<T, false, true>, tx + 32, smem);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    n = n < m ? n : m;
    return n < 1 ? 1 : n;
}

}  // namespace

// Whether this build's TMA kernels store with plain coalesced stores (no output tiles in shared memory).
bool tma_direct_store() { return LBM_TMA_DIRECT_STORE != 0; }

// How many CTAs of the TMA kernels fit one SM (registers and shared memory): the persistent grid is exactly that
// many per SM, so that it runs as ONE wave.
int tma_resident_ctas(bool f64, int tx, size_t smem, bool fast)
{
    return f64 ? resident_ctas<double>(tx, smem, fast) : resident_ctas<float>(tx, smem, fast);
}

// The opt-in to more than 48 KB of dynamic shared memory is per function and per device; done once per
// device when the first TMA context is created there (not per launch: launches may be inside a stream capture).
cudaError_t tma_prepare(int device)
{
    static bool done[64] = {};
    if (device >= 0 && device < 64 && done[device]) return cudaSuccess;
    cudaError_t e = prepare_t<float>();
    if (e == cudaSuccess) e = prepare_t<double>();
    if (e == cudaSuccess && device >= 0 && device < 64) done[device] = true;
    return e;
}

cudaError_t launch_tma_f32(const TmaCfg &k, const StepArgs<float> &a, int ns, bool macro, cudaStream_t s)
{
    return launch_tma<float>(k, a, ns, macro, s);
}
cudaError_t launch_tma_f64(const TmaCfg &k, const StepArgs<double> &a, int ns, bool macro, cudaStream_t s)
{
    return launch_tma<double>(k, a, ns, macro, s);
}

}  // namespace lbm
