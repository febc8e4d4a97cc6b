// Instantiations of the two-lattice pull kernel for ONE (precision, cells-per-thread) pair; the Makefile
// compiles this file five times (-DLBM_PULL_T=float|double -DLBM_PULL_VEC=1|2|4 -DLBM_PULL_FN=name).
#include "lbm_launch.hpp"

#if !defined(LBM_PULL_T) || !defined(LBM_PULL_VEC) || !defined(LBM_PULL_FN)
#error "compile with -DLBM_PULL_T=<type> -DLBM_PULL_VEC=<n> -DLBM_PULL_FN=<function name>"
#endif

namespace lbm {

namespace {

using T = LBM_PULL_T;
constexpr int VEC = LBM_PULL_VEC;

template <bool FAST, bool MACRO, int PEER, int LM>
void go(const LaunchCfg &k, const StepArgs<T> &a, cudaStream_t s)
{
    // PEER_FLAGS: blocks are one plane thick (the flag protocol counts blocks per plane)
    dim3 b = k.block;
    if (PEER == PEER_FLAGS) b.z = 1;
    LaunchCfg kk = k;
    kk.block = b;
    step_pull_kernel<T, VEC, FAST, MACRO, PEER, LM><<<step_grid(kk, VEC, planes_of(a)), b, 0, s>>>(a);
}

template <bool FAST, bool MACRO, int PEER>
void by_lm(const LaunchCfg &k, const StepArgs<T> &a, cudaStream_t s)
{
    switch (k.lm) {
        case LM_ROWS: go<FAST, MACRO, PEER, LM_ROWS>(k, a, s); break;
        case LM_SOA: go<FAST, MACRO, PEER, LM_SOA>(k, a, s); break;
        case LM_BLOCKROWS: go<FAST, MACRO, PEER, LM_BLOCKROWS>(k, a, s); break;
        default:
            // the generic addressing (a cross-check) exists for launches without neighbours only
            if constexpr (PEER == PEER_NONE) go<FAST, MACRO, PEER_NONE, LM_GENERIC>(k, a, s);
            break;
    }
}

template <bool FAST, bool MACRO>
void by_peer(const LaunchCfg &k, const StepArgs<T> &a, int peer, cudaStream_t s)
{
    if (peer == PEER_FLAGS) by_lm<FAST, MACRO, PEER_FLAGS>(k, a, s);
    else if (peer == PEER_STORE) by_lm<FAST, MACRO, PEER_STORE>(k, a, s);
    else by_lm<FAST, MACRO, PEER_NONE>(k, a, s);
}

}  // namespace

cudaError_t LBM_PULL_FN(const LaunchCfg &k, const StepArgs<T> &a, bool macro, int peer, cudaStream_t s)
{
    if (planes_of(a) <= 0) return cudaSuccess;
    if (peer != PEER_NONE && k.lm == LM_GENERIC) return cudaErrorInvalidValue;
    if (k.fast) { if (macro) by_peer<true, true>(k, a, peer, s); else by_peer<true, false>(k, a, peer, s); }
    else        { if (macro) by_peer<false, true>(k, a, peer, s); else by_peer<false, false>(k, a, peer, s); }
    return cudaGetLastError();
}

}  // namespace lbm
