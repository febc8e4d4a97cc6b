// The context behind the opaque lbm_ctx of include/lbm_b200.h, shared by the translation units that
// implement the C ABI (lbm_capi.cu: single context, z-slab transports; lbm_group.inl: same-process group;
// lbm_nvrtc.cu: run-time specialised kernels).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>  // types only: the functions are resolved with dlopen (no link-time dependency)

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm_launch.hpp"

struct LbmEventPair {
    cudaEvent_t start = nullptr;
    cudaEvent_t stop = nullptr;
    bool stop_recorded = false;
};

// how a slab context orders itself with its neighbours
enum LbmSlabSync : int {
    LBM_SYNC_NONE = 0,    // no neighbours, or ordering is the caller's business (lbm_group: events)
    LBM_SYNC_NCCL = 1,    // library-driven NCCL: dense halos, or peer stores + a one-word token
    LBM_SYNC_FLAGS = 2    // peer stores + in-kernel epoch flags, one launch per iteration
};

struct LbmNvrtcKernel;  // lbm_nvrtc.cu

struct lbm_ctx {
    lbm_params p{};
    int device = 0;
    std::string device_name;
    std::string error;

    // geometry of the slab
    int dim = 0;
    int z_begin = 0, z_end = 0;  // owned planes
    int zs0 = 0;                 // global z of stored plane 0
    int nz_local = 0;            // stored planes (owned + halos)
    long long n_local = 0;       // stored cells
    long long n_alloc = 0;       // stored cells rounded up to a multiple of the stride
    lbm::Layout lay{};
    int layout_mode = lbm::LM_GENERIC;     // what non-peer launches use (test hook: may be LM_GENERIC)
    int layout_natural = lbm::LM_GENERIC;  // what stride and DIM call for
    bool aa = false;             // in-place AA variant: only f[0] exists
    int swz_y = -1, swz_z = -1;  // block order of the step kernels (block_yz): log2 tile extents, < 0 = grid order
    bool swz_shift_only = false; // ... applied to the in-place variant's SHIFT launches only
    bool tma = false;            // TMA-fed variant: tensor maps of the two lattices
    CUtensorMap tmap[2];         // loads: box = one direction of one row tile
    CUtensorMap tmap_st[2];      // stores: box = one warp's 32 cells x 19 directions
    int tma_osdiv = 5;           // log2(min(stride, 32))
    int tma_tx = 0, tma_ns = 0, tma_grid = 0;
    size_t tma_smem = 0;
    int *tma_error = nullptr;    // device flag set by a kernel whose mbarrier wait timed out
    LbmNvrtcKernel *spec = nullptr;  // run-time specialised step kernel (LBM_VARIANT_NVRTC)
    int vec = 1;
    dim3 block{1, 1, 1};
    size_t esize = 4;

    // effective constants (after the reference's text round trip)
    double eff_viscosity = 0, eff_velocity = 0, eff_inv_tau = 0;
    lbm::Consts<float> cf{};
    lbm::Consts<double> cd{};
    float stale_f[2][lbm::Q]{};
    double stale_d[2][lbm::Q]{};

    // device memory
    void *f[2] = {nullptr, nullptr};  // the two lattices; f[0] carries the slab flag words behind the lattice
    size_t f_bytes = 0;               // bytes of one lattice
    size_t flag_off = 0;              // byte offset of the two incoming flag words inside the f[0] allocation
    void *rho = nullptr;
    void *u = nullptr;
    void *halo_send[2] = {nullptr, nullptr};
    void *halo_recv[2] = {nullptr, nullptr};
    int64_t device_bytes = 0;
    int cur = 0;  // index of the lattice the NEXT iteration reads

    // neighbours whose halo planes this context's boundary kernels write directly (peer stores):
    // same-process contexts (lbm_group, lbm_peer_attach) or lattices of other processes opened through
    // CUDA IPC (lbm_ipc_attach).  peer_f[face][lattice], peer_zs0[face] = global z of the neighbour's plane 0,
    // peer_flag[face] = the neighbour's flag word that belongs to me.
    lbm_ctx *peer[2] = {nullptr, nullptr};
    void *peer_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int peer_zs0[2] = {0, 0};
    unsigned *peer_flag[2] = {nullptr, nullptr};
    bool peer_ipc[2] = {false, false};

    // slab ordering
    int sync_mode = LBM_SYNC_NONE;
    bool fused = false;          // crossing populations travel as peer stores (else: dense halos over NCCL)
    unsigned phase = 0;          // LBM_SYNC_FLAGS: phases (initialisations + iterations) completed
    unsigned *sync_local = nullptr;  // device: count[2], error
    unsigned long long sync_timeout_ns = 20ull * 1000 * 1000 * 1000;

    // one process per device: NCCL communicator over the slabs (lbm_comm_init)
    ncclComm_t comm = nullptr;
    int comm_rank = -1, comm_world = 0;
    cudaStream_t bstream = nullptr;           // high-priority stream: boundary planes + exchange
    cudaEvent_t ev_bk[2] = {nullptr, nullptr};  // boundary kernels of iteration parity p done
    cudaEvent_t ev_in[2] = {nullptr, nullptr};  // interior kernel of iteration parity p done
    cudaEvent_t ev_join = nullptr;
    int *token = nullptr;                     // 3 ints: sent token, received from above, received from below

    // execution
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int64_t iteration = 0;
    int64_t launches = 0;
    bool initialised = false;

    // asynchronous rho/u read-back (lbm_read_macros_async): copies run on their own stream; the next
    // kernel that overwrites rho/u waits for them
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_ready = nullptr;  // compute stream: everything the copy must see is done
    cudaEvent_t ev_copy_done = nullptr;   // copy stream: rho/u have left the device buffers
    bool copy_pending = false;

    // CUDA graphs of LBM_GRAPH_CHUNK unflagged iterations for launch-bound (small) lattices,
    // one per starting parity
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    cudaStream_t graph_stream = nullptr;  // the stream the graphs were captured on

    // profiling (the reference's event list, lbmcl.hpp:74)
    cudaEvent_t ev_init_start = nullptr;
    cudaEvent_t ev_last = nullptr;
    std::vector<LbmEventPair> compute_events;
    double kernels_ms_accum = 0.0;  // folded-in pairs
    std::vector<float> launch_ms;   // duration of every folded pair, in enqueue order (bounded)
};

int lbm_fail(lbm_ctx *ctx, int code, const char *fmt, ...);

#define LBM_CUDA(ctx, ...)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (__VA_ARGS__);                                                                \
        if (e__ != cudaSuccess)                                                                         \
            return lbm_fail((ctx), e__ == cudaErrorMemoryAllocation ? LBM_ERR_OOM : LBM_ERR_CUDA,       \
                            "%s:%d %s(%d) - %s", __FILE__, __LINE__, #__VA_ARGS__, (int)e__, cudaGetErrorName(e__)); \
    } while (0)

// lbmcl.hpp:140-141 + kernels.cl:61-62: the value printed with 6 significant digits by operator<<
// from a T-typed member is what the kernel compiler parses back as a T literal.
template <typename T>
T text_roundtrip(double v);
template <>
inline float text_roundtrip<float>(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)(float)v);
    return strtof(buf, nullptr);
}
template <>
inline double text_roundtrip<double>(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", v);
    return strtod(buf, nullptr);
}

// Folded in T arithmetic exactly like the reference's compile-time constants.  `volatile` keeps the
// host compiler from contracting 3*nu + 0.5 into an FMA whatever its flags are.
template <typename T>
lbm::Consts<T> make_consts(double nu, double u_lid, double *eff)
{
    lbm::Consts<T> c;
    volatile T visc = text_roundtrip<T>(nu);
    volatile T three_nu = T(3.0) * visc;
    volatile T tau = three_nu + T(0.5);
    c.u_lid = text_roundtrip<T>(u_lid);
    c.inv_tau = T(1.0) / tau;
    c.w[0] = T(1.0) / T(3.0);
    c.w[1] = T(1.0) / T(18.0);
    c.w[2] = T(1.0) / T(36.0);
    eff[0] = (double)visc;
    eff[1] = (double)c.u_lid;
    eff[2] = (double)c.inv_tau;
    return c;
}

// lbm_nvrtc.cu: compile / launch the specialised step kernel of a context
int lbm_nvrtc_build(lbm_ctx *c);
void lbm_nvrtc_destroy(lbm_ctx *c);
cudaError_t lbm_nvrtc_launch(lbm_ctx *c, const void *step_args, bool macro, int n_planes, cudaStream_t s);
