// Instantiations of the in-place AA-pattern kernels (one lattice), both precisions.
#include "lbm_launch.hpp"

namespace lbm {

namespace {

template <typename T, bool FAST, bool MACRO, bool SHIFT>
void by_lm(const LaunchCfg &k, const StepArgs<T> &a, cudaStream_t s)
{
    const dim3 g = step_grid(k, 1, a.z_end - a.z_begin), b = k.block;
    switch (k.lm) {
        case LM_ROWS: step_aa_kernel<T, FAST, MACRO, SHIFT, LM_ROWS><<<g, b, 0, s>>>(a); break;
        case LM_SOA: step_aa_kernel<T, FAST, MACRO, SHIFT, LM_SOA><<<g, b, 0, s>>>(a); break;
        case LM_BLOCKROWS: step_aa_kernel<T, FAST, MACRO, SHIFT, LM_BLOCKROWS><<<g, b, 0, s>>>(a); break;
        default: step_aa_kernel<T, FAST, MACRO, SHIFT, LM_GENERIC><<<g, b, 0, s>>>(a); break;
    }
}

template <typename T>
cudaError_t launch_aa(const LaunchCfg &k, const StepArgs<T> &a, bool macro, bool shift, cudaStream_t s)
{
    if (a.z_end <= a.z_begin) return cudaSuccess;
#define LBM_AA(F, M)                                  \
    do {                                              \
        if (shift) by_lm<T, F, M, true>(k, a, s);     \
        else by_lm<T, F, M, false>(k, a, s);          \
    } while (0)
    if (k.fast) { if (macro) LBM_AA(true, true); else LBM_AA(true, false); }
    else        { if (macro) LBM_AA(false, true); else LBM_AA(false, false); }
#undef LBM_AA
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_aa_f32(const LaunchCfg &k, const StepArgs<float> &a, bool macro, bool shift, cudaStream_t s)
{
    return launch_aa<float>(k, a, macro, shift, s);
}
cudaError_t launch_aa_f64(const LaunchCfg &k, const StepArgs<double> &a, bool macro, bool shift, cudaStream_t s)
{
    return launch_aa<double>(k, a, macro, shift, s);
}

}  // namespace lbm
