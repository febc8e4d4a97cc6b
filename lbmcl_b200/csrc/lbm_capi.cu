// C ABI of the B200-native D3Q19 collide-and-stream path (see include/lbm_b200.h).
//
// This file replaces what the reference's lbmcl.hpp obtains from libs/CLUtil.hpp + cl.hpp: device
// selection, buffers, kernel launches with the ping-pong binding, readbacks and event timing.
// There is no CPU fallback anywhere in here: without a CUDA device lbm_create fails.
#include "../../include/lbm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the functions are resolved with dlopen (no link-time dependency)

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "lbm_kernels.cuh"
#include "lbm_tma.cuh"

using namespace lbm;

namespace {

thread_local std::string g_create_error;

struct EventPair {
    cudaEvent_t start = nullptr;
    cudaEvent_t stop = nullptr;
};

}  // namespace

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
    Layout lay{};
    int layout_mode = LM_GENERIC;
    bool aa = false;             // in-place AA variant: only f[0] exists
    bool tma = false;            // TMA-fed variant: tensor maps of the two lattices
    CUtensorMap tmap[2];
    int tma_tx = 0, tma_ns = 0, tma_grid = 0;
    size_t tma_smem = 0;
    int *tma_error = nullptr;    // device flag set by a kernel whose mbarrier wait timed out
    int vec = 1;
    dim3 block{1, 1, 1};
    size_t esize = 4;

    // effective constants (after the reference's text round trip)
    double eff_viscosity = 0, eff_velocity = 0, eff_inv_tau = 0;
    Consts<float> cf{};
    Consts<double> cd{};
    float stale_f[2][Q]{};
    double stale_d[2][Q]{};

    // device memory
    void *f[2] = {nullptr, nullptr};  // the two lattices
    void *rho = nullptr;
    void *u = nullptr;
    void *halo_send[2] = {nullptr, nullptr};
    void *halo_recv[2] = {nullptr, nullptr};
    int64_t device_bytes = 0;
    int cur = 0;  // index of the lattice the NEXT iteration reads

    // neighbours whose halo planes this context's boundary kernels write directly (peer stores):
    // same-process group members (lbm_group) or lattices of other processes opened through CUDA IPC
    // (lbm_ipc_attach).  peer_f[face][lattice], peer_zs0[face] = global z of the neighbour's plane 0.
    lbm_ctx *peer[2] = {nullptr, nullptr};
    void *peer_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int peer_zs0[2] = {0, 0};
    bool peer_ipc[2] = {false, false};

    // one process per device: NCCL communicator over the slabs (lbm_comm_init)
    ncclComm_t comm = nullptr;
    int comm_rank = -1, comm_world = 0;
    cudaStream_t bstream = nullptr;           // high-priority stream: boundary planes + exchange
    cudaEvent_t ev_bk[2] = {nullptr, nullptr};  // boundary kernels of iteration parity p done
    cudaEvent_t ev_in[2] = {nullptr, nullptr};  // interior kernel of iteration parity p done
    cudaEvent_t ev_join = nullptr;
    int *token = nullptr;                     // 3 ints: sent token, received from above, received from below
    bool fused = false;                       // lbm_comm_fused(): peer stores + token instead of dense halos

    // execution
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int64_t iteration = 0;
    int64_t launches = 0;
    bool initialised = false;

    // CUDA graphs of LBM_GRAPH_CHUNK unflagged iterations for launch-bound (small) lattices,
    // one per starting parity
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    cudaStream_t graph_stream = nullptr;  // the stream the graphs were captured on

    // profiling (the reference's event list, lbmcl.hpp:74)
    cudaEvent_t ev_init_start = nullptr;
    cudaEvent_t ev_last = nullptr;
    std::vector<EventPair> compute_events;
    double kernels_ms_accum = 0.0;  // folded-in pairs
    std::vector<float> launch_ms;   // duration of every folded pair, in enqueue order (bounded)
};

static void lbm_nccl_destroy(ncclComm_t comm);
static int setup_tma(lbm_ctx *c, int n_sm);

namespace {

int fail(lbm_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_create_error = buf;
    return code;
}

#define LBM_CUDA(ctx, ...)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (__VA_ARGS__);                                                                \
        if (e__ != cudaSuccess)                                                                         \
            return fail((ctx), e__ == cudaErrorMemoryAllocation ? LBM_ERR_OOM : LBM_ERR_CUDA,           \
                        "%s:%d %s(%d) - %s", __FILE__, __LINE__, #__VA_ARGS__, (int)e__, cudaGetErrorName(e__)); \
    } while (0)

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2(long long v)
{
    int n = 0;
    while (v > 1) { v >>= 1; ++n; }
    return n;
}
int floor_pow2(int v)
{
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// lbmcl.hpp:140-141 + kernels.cl:61-62: the value printed with 6 significant digits by operator<<
// from a T-typed member is what the kernel compiler parses back as a T literal.
template <typename T>
T text_roundtrip(double v);
template <>
float text_roundtrip<float>(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", (double)(float)v);
    return strtof(buf, nullptr);
}
template <>
double text_roundtrip<double>(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", v);
    return strtod(buf, nullptr);
}

// Folded in T arithmetic exactly like the reference's compile-time constants.  `volatile` keeps the
// host compiler from contracting 3*nu + 0.5 into an FMA whatever its flags are.
template <typename T>
Consts<T> make_consts(double nu, double u_lid, double *eff)
{
    Consts<T> c;
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

template <typename T>
StepArgs<T> make_step_args(lbm_ctx *c, int z_begin, int z_end, const Consts<T> &k, const T (&stale)[2][Q])
{
    StepArgs<T> a{};
    a.dst = static_cast<T *>(c->f[c->cur ^ 1]);
    a.src = static_cast<const T *>(c->f[c->cur]);
    a.rho = static_cast<T *>(c->rho);
    a.u = static_cast<T *>(c->u);
    a.peer_lo = nullptr;
    a.peer_hi = nullptr;
    // my first / last owned plane is the neighbour's high / low halo plane, in the lattice it reads
    // next (all slabs advance in lock step, so the neighbour's lattice index equals mine)
    if (c->peer_f[0][0]) {
        a.peer_lo = static_cast<T *>(c->peer_f[0][c->cur ^ 1]);
        a.peer_lo_plane = c->z_begin - c->peer_zs0[0];
    }
    if (c->peer_f[1][0]) {
        a.peer_hi = static_cast<T *>(c->peer_f[1][c->cur ^ 1]);
        a.peer_hi_plane = (c->z_end - 1) - c->peer_zs0[1];
    }
    a.z_own_begin = c->z_begin;
    a.z_own_end = c->z_end;
    a.dim = c->dim;
    a.zs0 = c->zs0;
    a.z_begin = z_begin;
    a.z_end = z_end;
    a.n_local = c->n_local;
    a.lay = c->lay;
    a.c = k;
    for (int i = 0; i < 2; ++i)
        for (int q = 0; q < Q; ++q) a.stale[i][q] = stale[i][q];
    const long long S = c->lay.qpitch(), dim = c->dim, plane = dim * dim, es = (long long)sizeof(T);
    for (int q = 0; q < Q; ++q) {
        a.soff[q] = q * S * es;
        const long long dcell = (long long)ey(q) * dim + (long long)ez(q) * plane;  // cells between the two rows
        const long long rowmul = c->layout_mode == LM_ROWS ? Q : 1;                  // CSoA rows hold Q values per cell
        if (c->layout_mode == LM_GENERIC) {
            a.goff[q] = a.poff[q] = 0;
        } else if (c->aa) {
            a.goff[q] = (opp(q) * S - rowmul * dcell) * es;  // SHIFT step: read (c - e_q, opp(q))
            a.poff[q] = (q * S + rowmul * dcell) * es;       //             write (c + e_q, q)
        } else {
            a.goff[q] = (q * S - rowmul * dcell) * es;
            a.poff[q] = 0;
        }
    }
    if (c->aa) {
        a.dst = static_cast<T *>(c->f[0]);
        a.src = static_cast<const T *>(c->f[0]);
    }
    return a;
}

template <typename T, int VEC>
cudaError_t launch_step_t(lbm_ctx *c, const StepArgs<T> &a, bool macro, bool peer, cudaStream_t s)
{
    const int nz = a.z_end - a.z_begin;
    if (nz <= 0) return cudaSuccess;
    const dim3 b = c->block;
    const dim3 g((unsigned)(c->dim / (b.x * VEC)), (unsigned)(c->dim / b.y), (unsigned)((nz + b.z - 1) / b.z));
    const bool fast = c->p.fast_math != 0;
#define LBM_LAUNCH_LM(F, M, P)                                                                   \
    do {                                                                                          \
        if (c->layout_mode == LM_ROWS) step_pull_kernel<T, VEC, F, M, P, LM_ROWS><<<g, b, 0, s>>>(a);     \
        else if (c->layout_mode == LM_SOA) step_pull_kernel<T, VEC, F, M, P, LM_SOA><<<g, b, 0, s>>>(a);  \
        else step_pull_kernel<T, VEC, F, M, P, LM_GENERIC><<<g, b, 0, s>>>(a);                    \
    } while (0)
#define LBM_LAUNCH(F, M, P) LBM_LAUNCH_LM(F, M, P)
    if (peer) {
        if (fast) { if (macro) LBM_LAUNCH(true, true, true); else LBM_LAUNCH(true, false, true); }
        else      { if (macro) LBM_LAUNCH(false, true, true); else LBM_LAUNCH(false, false, true); }
    } else {
        if (fast) { if (macro) LBM_LAUNCH(true, true, false); else LBM_LAUNCH(true, false, false); }
        else      { if (macro) LBM_LAUNCH(false, true, false); else LBM_LAUNCH(false, false, false); }
    }
#undef LBM_LAUNCH
#undef LBM_LAUNCH_LM
    c->launches += 1;
    return cudaGetLastError();
}

// AA variant: iteration `it` (1-based) is a LOCAL step when odd, a SHIFT step when even.
template <typename T>
cudaError_t launch_aa_t(lbm_ctx *c, const StepArgs<T> &a, bool macro, cudaStream_t s)
{
    const int nz = a.z_end - a.z_begin;
    if (nz <= 0) return cudaSuccess;
    const dim3 b = c->block;
    const dim3 g((unsigned)(c->dim / b.x), (unsigned)(c->dim / b.y), (unsigned)((nz + b.z - 1) / b.z));
    const bool fast = c->p.fast_math != 0;
    const bool shift = ((c->iteration + 1) % 2) == 0;
#define LBM_AA_LM(F, M, SH)                                                                       \
    do {                                                                                          \
        if (c->layout_mode == LM_ROWS) step_aa_kernel<T, F, M, SH, LM_ROWS><<<g, b, 0, s>>>(a);   \
        else if (c->layout_mode == LM_SOA) step_aa_kernel<T, F, M, SH, LM_SOA><<<g, b, 0, s>>>(a); \
        else step_aa_kernel<T, F, M, SH, LM_GENERIC><<<g, b, 0, s>>>(a);                          \
    } while (0)
#define LBM_AA(F, M)                           \
    do {                                       \
        if (shift) LBM_AA_LM(F, M, true);      \
        else LBM_AA_LM(F, M, false);           \
    } while (0)
    if (fast) { if (macro) LBM_AA(true, true); else LBM_AA(true, false); }
    else      { if (macro) LBM_AA(false, true); else LBM_AA(false, false); }
#undef LBM_AA
#undef LBM_AA_LM
    c->launches += 1;
    return cudaGetLastError();
}

// TMA-fed variant: persistent CTAs over the live rows of [z_begin, z_end).
template <typename T, int TX>
cudaError_t launch_tma_tx(lbm_ctx *c, const StepArgs<T> &sa, const TmaArgs<T> &a, bool macro, cudaStream_t s, int grid)
{
    const bool fast = c->p.fast_math != 0;
    const CUtensorMap &ms = c->tmap[c->cur], &md = c->tmap[c->cur ^ 1];
    (void)sa;
    // the opt-in to > 48 KB of dynamic shared memory is per function AND per device: set it on every
    // launch (a host-side table update) rather than caching it per process
#define LBM_TMA_LAUNCH(F, M)                                                                                    \
    do {                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(step_tma_kernel<T, F, M, TX>,                                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);         \
        if (e != cudaSuccess) return e;                                                                         \
        step_tma_kernel<T, F, M, TX><<<grid, TX, c->tma_smem, s>>>(ms, md, a, c->tma_error);                   \
    } while (0)
    if (fast) { if (macro) LBM_TMA_LAUNCH(true, true); else LBM_TMA_LAUNCH(true, false); }
    else      { if (macro) LBM_TMA_LAUNCH(false, true); else LBM_TMA_LAUNCH(false, false); }
#undef LBM_TMA_LAUNCH
    c->launches += 1;
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_tma_t(lbm_ctx *c, const StepArgs<T> &sa, bool macro, cudaStream_t s)
{
    // live planes of this launch
    const int zf = sa.z_begin < 1 ? 1 : sa.z_begin;
    const int zl = sa.z_end > c->dim - 1 ? c->dim - 1 : sa.z_end;
    if (zl <= zf) return cudaSuccess;
    TmaArgs<T> a{};
    a.src = sa.src;
    a.rho = sa.rho;
    a.u = sa.u;
    a.dim = c->dim;
    a.zs0 = c->zs0;
    a.z_first = zf;
    a.n_xseg = c->dim / c->tma_tx;
    a.n_tiles = (zl - zf) * (c->dim - 2) * a.n_xseg;
    a.ns = c->tma_ns;
    a.n_local = c->n_local;
    a.lay = c->lay;
    a.c = sa.c;
    for (int i = 0; i < 2; ++i)
        for (int q = 0; q < Q; ++q) a.stale[i][q] = sa.stale[i][q];
    const int grid = a.n_tiles < c->tma_grid ? a.n_tiles : c->tma_grid;
    switch (c->tma_tx) {
        case 256: return launch_tma_tx<T, 256>(c, sa, a, macro, s, grid);
        case 128: return launch_tma_tx<T, 128>(c, sa, a, macro, s, grid);
        case 64: return launch_tma_tx<T, 64>(c, sa, a, macro, s, grid);
        default: return launch_tma_tx<T, 32>(c, sa, a, macro, s, grid);
    }
}

template <typename T>
cudaError_t launch_step_p(lbm_ctx *c, int z_begin, int z_end, bool macro, cudaStream_t s,
                          const Consts<T> &k, const T (&stale)[2][Q])
{
    const StepArgs<T> a = make_step_args<T>(c, z_begin, z_end, k, stale);
    if (c->aa) return launch_aa_t<T>(c, a, macro, s);
    const bool peer0 = (a.peer_lo != nullptr && z_begin <= c->z_begin && c->z_begin < z_end) ||
                       (a.peer_hi != nullptr && z_begin <= c->z_end - 1 && c->z_end - 1 < z_end);
    if (c->tma && !peer0) return launch_tma_t<T>(c, a, macro, s);
    const bool peer = (a.peer_lo != nullptr && z_begin <= c->z_begin && c->z_begin < z_end) ||
                      (a.peer_hi != nullptr && z_begin <= c->z_end - 1 && c->z_end - 1 < z_end);
    switch (c->vec) {
        case 4:
            if constexpr (sizeof(T) == 4) return launch_step_t<T, 4>(c, a, macro, peer, s);
            else return cudaErrorInvalidValue;
        case 2: return launch_step_t<T, 2>(c, a, macro, peer, s);
        default: return launch_step_t<T, 1>(c, a, macro, peer, s);
    }
}

// one launch of the step kernel over global planes [z_begin, z_end) (clipped to the computed range)
cudaError_t launch_step(lbm_ctx *c, int z_begin, int z_end, bool macro, cudaStream_t s)
{
    if (c->p.precision == LBM_F32) return launch_step_p<float>(c, z_begin, z_end, macro, s, c->cf, c->stale_f);
    return launch_step_p<double>(c, z_begin, z_end, macro, s, c->cd, c->stale_d);
}

template <typename T>
cudaError_t launch_init_t(lbm_ctx *c, const Consts<T> &k, cudaStream_t s)
{
    InitArgs<T> a{};
    a.f0 = static_cast<T *>(c->f[0]);
    a.f1 = static_cast<T *>(c->f[1]);
    a.rho = static_cast<T *>(c->rho);
    a.u = static_cast<T *>(c->u);
    a.dim = c->dim;
    a.zs0 = c->zs0;
    a.nz_local = c->nz_local;
    a.n_local = c->n_local;
    a.lay = c->lay;
    a.c = k;
    const int bx = c->dim < 64 ? c->dim : 64;
    const int by = (256 / bx) < c->dim ? (256 / bx) : c->dim;
    const dim3 b(bx, by, 1);
    const dim3 g(c->dim / bx, c->dim / by, c->nz_local);
    if (c->aa) init_aa_kernel<T><<<g, b, 0, s>>>(a);
    else init_kernel<T><<<g, b, 0, s>>>(a);
    return cudaGetLastError();
}

template <typename T>
int compute_stale(lbm_ctx *c, const Consts<T> &k, T (&stale)[2][Q])
{
    T *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, sizeof(T) * 2 * Q));
    stale_kernel<T><<<1, 1, 0, c->stream>>>(k, d);
    LBM_CUDA(c, cudaGetLastError());
    LBM_CUDA(c, cudaMemcpyAsync(&stale[0][0], d, sizeof(T) * 2 * Q, cudaMemcpyDeviceToHost, c->stream));
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    LBM_CUDA(c, cudaFree(d));
    return LBM_OK;
}

// CUDA block of the step kernel.  Constraints: powers of two, bx*VEC | DIM, by | DIM, at most 256
// threads (the kernels are compiled with __launch_bounds__(256)).
//   default            x-major rows: bx = min(DIM/VEC, 256), the rest of the 256 threads in y, then z.
//                      Measured on B200 (profiles/r01_ncu_summary.md, sweeps): whole x-rows per block are never worse
//                      than any other shape, and shapes with a short x extent (the reference default
//                      -w 8,8,8) lose coalescing.  The requested work-group size is therefore a hint
//                      that does not change the shape -- it never changes the results either.
//   reserved[1] == 1   honour the requested shape as far as the constraints allow (sweeps, tests).
void choose_block(lbm_ctx *c)
{
    const int dim = c->dim;
    int bx, by, bz;
    if (c->p.reserved[1] == 1) {
        bx = floor_pow2(c->p.block_x > 0 ? c->p.block_x : 1) / c->vec;
        if (bx < 1) bx = 1;
        if (bx > dim / c->vec) bx = dim / c->vec;
        by = floor_pow2(c->p.block_y > 0 ? c->p.block_y : 1);
        if (by > dim) by = dim;
        bz = floor_pow2(c->p.block_z > 0 ? c->p.block_z : 1);
        if (bz > dim) bz = dim;
        if (bz > 64) bz = 64;
        while (bx * by * bz > 256) {
            if (bz > 1) bz /= 2;
            else if (by > 1) by /= 2;
            else bx /= 2;
        }
    } else {
        bx = dim / c->vec < 256 ? dim / c->vec : 256;
        by = 256 / bx < dim ? 256 / bx : dim;
        bz = 256 / (bx * by) < dim ? 256 / (bx * by) : dim;
        if (bz > 64) bz = 64;
        if (bz < 1) bz = 1;
    }
    c->block = dim3(bx, by, bz);
}

int record_last(lbm_ctx *c)
{
    LBM_CUDA(c, cudaEventRecord(c->ev_last, c->stream));
    return LBM_OK;
}

int use_device(lbm_ctx *c)
{
    LBM_CUDA(c, cudaSetDevice(c->device));
    return LBM_OK;
}

}  // namespace

// ---- CUDA graphs for launch-bound lattices ----
namespace {

constexpr int LBM_GRAPH_CHUNK = 16;     // even: lattice parity and AA step type are restored after a chunk
constexpr int LBM_GRAPH_MAX_DIM = 64;   // above this one launch takes longer than its enqueue
constexpr int LBM_GRAPH_MIN_CHUNKS = 8; // replays needed to amortise capture + instantiation

// Enqueue LBM_GRAPH_CHUNK iterations without macro store as ONE graph launch.  The graph is captured
// on first use for the current parity (source lattice / AA step type) and stream, then replayed.
int launch_graph_chunk(lbm_ctx *c)
{
    const int par = c->aa ? (int)(c->iteration & 1) : c->cur;
    if (c->graph_stream != c->stream) {  // captured on another stream: rebuild
        for (int i = 0; i < 2; ++i)
            if (c->graph_exec[i]) {
                cudaGraphExecDestroy(c->graph_exec[i]);
                c->graph_exec[i] = nullptr;
            }
        c->graph_stream = c->stream;
    }
    const int cur0 = c->cur;
    const int64_t it0 = c->iteration, launches0 = c->launches;
    if (!c->graph_exec[par]) {
        cudaGraph_t graph = nullptr;
        LBM_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < LBM_GRAPH_CHUNK && e == cudaSuccess; ++i) {
            e = launch_step(c, c->z_begin, c->z_end, false, c->stream);
            c->cur ^= 1;
            c->iteration += 1;
        }
        const cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
        c->cur = cur0;
        c->iteration = it0;
        c->launches = launches0;
        if (e != cudaSuccess || e2 != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            LBM_CUDA(c, e != cudaSuccess ? e : e2);
        }
        const cudaError_t e3 = cudaGraphInstantiate(&c->graph_exec[par], graph, 0);
        cudaGraphDestroy(graph);
        LBM_CUDA(c, e3);
    }
    LBM_CUDA(c, cudaGraphLaunch(c->graph_exec[par], c->stream));
    c->iteration += LBM_GRAPH_CHUNK;  // even chunk: c->cur unchanged
    c->launches += LBM_GRAPH_CHUNK;
    return LBM_OK;
}

}  // namespace

// ---- NCCL, resolved at run time ----
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // RTLD_NOLOAD first: under torchrun the process already holds torch's libnccl.so.2
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return api;
    }
    api.handle = h;
#define LBM_NCCL_SYM(field, name)                                             \
    do {                                                                      \
        *(void **)(&api.field) = dlsym(h, name);                              \
        if (!api.field) api.error = std::string("libnccl lacks ") + name;     \
    } while (0)
    LBM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    LBM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    LBM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    LBM_NCCL_SYM(Send, "ncclSend");
    LBM_NCCL_SYM(Recv, "ncclRecv");
    LBM_NCCL_SYM(GroupStart, "ncclGroupStart");
    LBM_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    LBM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef LBM_NCCL_SYM
    return api;
}

#define LBM_NCCL(ctx, call)                                                                             \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess)                                                                         \
            return fail((ctx), LBM_ERR_CUDA, "%s:%d %s(%d) - %s", __FILE__, __LINE__, #call, (int)r__,  \
                        nccl().GetErrorString ? nccl().GetErrorString(r__) : "?");                      \
    } while (0)

// The overlapped z-slab schedule of one rank (see include/lbm_b200.h, transport 2b).
//   bstream: wait interior(k-1) -> boundary planes(k) -> [advance] pack -> send/recv -> unpack
//   stream : wait boundary kernels(k-1) -> interior planes(k)
// The events alternate with the iteration parity; everything is enqueued without host synchronisation.
int run_slab_with_comm(lbm_ctx *c, int n_iterations, int every)
{
    NcclApi &n = nccl();
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    const ncclDataType_t dt = c->p.precision == LBM_F32 ? ncclFloat32 : ncclFloat64;
    const size_t count = (size_t)5 * c->dim * c->dim;
    cudaStream_t S = c->stream, B = c->bstream;
    // fused transport: every interior face has an IPC-attached neighbour
    const bool fused = c->fused;
    // the boundary stream starts after whatever the main stream was asked to do before
    LBM_CUDA(c, cudaEventRecord(c->ev_join, S));
    LBM_CUDA(c, cudaStreamWaitEvent(B, c->ev_join, 0));
    // one word to / from each neighbour, stream-ordered on B
    auto exchange_tokens = [&]() -> ncclResult_t {
        ncclResult_t r = n.GroupStart();
        if (r == ncclSuccess && has_hi) r = n.Send(c->token, 1, ncclInt32, c->comm_rank + 1, c->comm, B);
        if (r == ncclSuccess && has_hi) r = n.Recv(c->token + 1, 1, ncclInt32, c->comm_rank + 1, c->comm, B);
        if (r == ncclSuccess && has_lo) r = n.Send(c->token, 1, ncclInt32, c->comm_rank - 1, c->comm, B);
        if (r == ncclSuccess && has_lo) r = n.Recv(c->token + 2, 1, ncclInt32, c->comm_rank - 1, c->comm, B);
        const ncclResult_t e = n.GroupEnd();
        return r == ncclSuccess ? e : r;
    };
    if (fused && c->iteration == 0) {
        // first iteration after lbm_init: my boundary kernel will store into the neighbours' halo planes,
        // which their `initialize` kernels also write -- wait until the neighbours' initialisation is done
        const ncclResult_t r = exchange_tokens();
        if (r != ncclSuccess)
            return fail(c, LBM_ERR_CUDA, "NCCL token exchange failed (%d) - %s", (int)r, n.GetErrorString(r));
    }
    for (int i = 0; i < n_iterations; ++i) {
        const int64_t it = c->iteration + 1;
        const bool macro = every != 0 && (it % every) == 0;
        const int par = (int)(it & 1), prev = par ^ 1;
        const int zlo = c->z_begin, zhi = c->z_end - 1;
        LBM_CUDA(c, cudaStreamWaitEvent(B, c->ev_in[prev], 0));
        if (has_lo) LBM_CUDA(c, launch_step(c, zlo, zlo + 1, macro, B));
        if (has_hi && !(has_lo && zhi == zlo)) LBM_CUDA(c, launch_step(c, zhi, zhi + 1, macro, B));
        LBM_CUDA(c, cudaEventRecord(c->ev_bk[par], B));

        LBM_CUDA(c, cudaStreamWaitEvent(S, c->ev_bk[prev], 0));
        LBM_CUDA(c, launch_step(c, zlo + (has_lo ? 1 : 0), c->z_end - (has_hi ? 1 : 0), macro, S));
        LBM_CUDA(c, cudaEventRecord(c->ev_in[par], S));

        c->cur ^= 1;
        c->iteration = it;

        if (fused) {
            // the boundary kernels have already stored the crossing populations in the neighbours' halo
            // planes (IPC-mapped peer memory); only a stream-ordered token travels through NCCL: the
            // neighbour's next boundary kernel starts after my boundary kernel has completed
            const ncclResult_t r = exchange_tokens();
            if (r != ncclSuccess)
                return fail(c, LBM_ERR_CUDA, "NCCL token exchange failed (%d) - %s", (int)r, n.GetErrorString(r));
            continue;
        }
        // exchange on the boundary stream (pack / unpack use c->stream: point it at B for a moment)
        c->stream = B;
        int rc = lbm_halo_pack(c);
        if (rc == LBM_OK) {
            ncclResult_t r = n.GroupStart();
            if (r == ncclSuccess && has_hi) r = n.Send(c->halo_send[1], count, dt, c->comm_rank + 1, c->comm, B);
            if (r == ncclSuccess && has_hi) r = n.Recv(c->halo_recv[1], count, dt, c->comm_rank + 1, c->comm, B);
            if (r == ncclSuccess && has_lo) r = n.Send(c->halo_send[0], count, dt, c->comm_rank - 1, c->comm, B);
            if (r == ncclSuccess && has_lo) r = n.Recv(c->halo_recv[0], count, dt, c->comm_rank - 1, c->comm, B);
            const ncclResult_t e = n.GroupEnd();
            if (r == ncclSuccess) r = e;
            if (r != ncclSuccess) {
                c->stream = S;
                return fail(c, LBM_ERR_CUDA, "NCCL halo exchange failed (%d) - %s", (int)r, n.GetErrorString(r));
            }
            rc = lbm_halo_unpack(c);
        }
        c->stream = S;
        if (rc != LBM_OK) return rc;
    }
    // the main stream's timeline ends after the last exchange
    LBM_CUDA(c, cudaEventRecord(c->ev_join, B));
    LBM_CUDA(c, cudaStreamWaitEvent(S, c->ev_join, 0));
    return LBM_OK;
}

}  // namespace

// Tensor maps of the two lattices for the TMA-fed variant: the CSoA lattice as the 3-D tensor
// [block][q][stride]; box = one x-row segment of one direction.
static int setup_tma(lbm_ctx *c, int n_sm)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LBM_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(c, LBM_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    EncodeFn encode = reinterpret_cast<EncodeFn>(fn);

    const long long S = c->lay.qpitch();
    c->tma_tx = c->dim < 256 ? c->dim : 256;
    if (const char *e = std::getenv("LBM_TMA_TX")) {  // test hook: narrower tiles => several segments per row
        const int v = std::atoi(e);
        if ((v == 32 || v == 64 || v == 128 || v == 256) && v <= c->dim) c->tma_tx = v;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)S, (cuuint64_t)Q, (cuuint64_t)(c->n_alloc / S)};
    const cuuint64_t gstride[2] = {(cuuint64_t)(S * c->esize), (cuuint64_t)(Q * S * c->esize)};
    const cuuint32_t b0 = (cuuint32_t)(S < c->tma_tx ? S : c->tma_tx);
    const cuuint32_t box[3] = {b0, 1u, (cuuint32_t)(c->tma_tx / b0)};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = c->p.precision == LBM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    for (int i = 0; i < 2; ++i) {
        const CUresult r = encode(&c->tmap[i], dt, 3, c->f[i], gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, LBM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    // ring depth / resident CTAs (profiles/r01_tma_experiment.md): the kernel is latency-bound per CTA,
    // so the shallowest ring with the most resident CTAs wins: 2 stages, as many CTAs as fit in 200 KB
    const size_t stage = (size_t)Q * c->tma_tx * c->esize;
    int ns = 2;
    int ctas = (int)((200 * 1024) / (2 * stage + 1024));
    if (ctas > 5) ctas = 5;
    if (const char *e = std::getenv("LBM_TMA_NS")) ns = std::atoi(e);
    if (const char *e = std::getenv("LBM_TMA_CTAS")) ctas = std::atoi(e);
    if (ns < 2) ns = 2;
    if (ctas < 1) ctas = 1;
    while (ns > 2 && (size_t)ns * stage + 64 > 200 * 1024) --ns;
    c->tma_ns = ns;
    c->tma_smem = (size_t)ns * stage + (size_t)ns * sizeof(uint64_t);
    if (c->tma_smem > 200 * 1024) return fail(c, LBM_ERR_INVALID, "TMA variant: tile ring does not fit shared memory");
    c->tma_grid = ctas * n_sm;
    LBM_CUDA(c, cudaMalloc(&c->tma_error, sizeof(int)));
    LBM_CUDA(c, cudaMemset(c->tma_error, 0, sizeof(int)));
    return LBM_OK;
}

static void lbm_nccl_destroy(ncclComm_t comm)
{
    if (comm && nccl().CommDestroy) nccl().CommDestroy(comm);
}

extern "C" {

int lbm_comm_unique_id(uint8_t id[LBM_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == LBM_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return LBM_ERR_INVALID;
    NcclApi &n = nccl();
    if (!n.error.empty() || !n.GetUniqueId) return fail(nullptr, LBM_ERR_CUDA, "lbm_comm_unique_id: %s", n.error.c_str());
    ncclUniqueId u;
    const ncclResult_t r = n.GetUniqueId(&u);
    if (r != ncclSuccess) return fail(nullptr, LBM_ERR_CUDA, "ncclGetUniqueId(%d) - %s", (int)r, n.GetErrorString(r));
    std::memcpy(id, &u, sizeof u);
    return LBM_OK;
}

int lbm_comm_init(lbm_ctx *c, const uint8_t id[LBM_COMM_ID_BYTES], int rank, int world)
{
    if (!c || !id) return LBM_ERR_INVALID;
    if (c->comm) return fail(c, LBM_ERR_STATE, "lbm_comm_init: already initialised");
    if (c->aa) return fail(c, LBM_ERR_INVALID, "lbm_comm_init: the AA variant is single-device");
    if (world < 1 || rank < 0 || rank >= world) return fail(c, LBM_ERR_INVALID, "lbm_comm_init: rank %d of %d", rank, world);
    // the slabs must tile the cube in rank order
    if ((rank == 0) != (c->z_begin == 0) || (rank == world - 1) != (c->z_end == c->dim))
        return fail(c, LBM_ERR_INVALID, "lbm_comm_init: planes [%d, %d) do not fit rank %d of %d", c->z_begin,
                    c->z_end, rank, world);
    NcclApi &n = nccl();
    if (!n.error.empty()) return fail(c, LBM_ERR_CUDA, "lbm_comm_init: %s", n.error.c_str());
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    int lo = 0, hi = 0;
    LBM_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LBM_CUDA(c, cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; ++i) {
        LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_bk[i], cudaEventDisableTiming));
        LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    }
    LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    LBM_NCCL(c, n.CommInitRank(&c->comm, world, u, rank));
    c->comm_rank = rank;
    c->comm_world = world;
    LBM_CUDA(c, cudaMalloc(&c->token, 3 * sizeof(int)));
    LBM_CUDA(c, cudaMemset(c->token, 0, 3 * sizeof(int)));
    return LBM_OK;
}

int lbm_comm_fused(lbm_ctx *c, int enable)
{
    if (!c) return LBM_ERR_INVALID;
    if (!enable) {
        c->fused = false;
        return LBM_OK;
    }
    if (!c->comm) return fail(c, LBM_ERR_STATE, "lbm_comm_fused: no communicator (lbm_comm_init)");
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    if ((has_lo && !c->peer_f[0][0]) || (has_hi && !c->peer_f[1][0]))
        return fail(c, LBM_ERR_STATE, "lbm_comm_fused: an interior face has no attached neighbour (lbm_ipc_attach)");
    c->fused = has_lo || has_hi;
    return LBM_OK;
}

// ---- CUDA IPC: let a neighbouring PROCESS's boundary kernel store straight into this lattice ----
namespace {
struct IpcBlob {
    cudaIpcMemHandle_t f[2];
    int32_t zs0, nz_local, dim, precision;
    int64_t stride, n_alloc;
};
static_assert(sizeof(IpcBlob) <= LBM_IPC_HANDLE_BYTES, "IpcBlob must fit LBM_IPC_HANDLE_BYTES");
}  // namespace

int lbm_ipc_export(lbm_ctx *c, uint8_t out[LBM_IPC_HANDLE_BYTES])
{
    if (!c || !out) return LBM_ERR_INVALID;
    if (c->aa) return fail(c, LBM_ERR_INVALID, "lbm_ipc_export: the AA variant is single-device");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    IpcBlob b{};
    for (int i = 0; i < 2; ++i) LBM_CUDA(c, cudaIpcGetMemHandle(&b.f[i], c->f[i]));
    b.zs0 = c->zs0;
    b.nz_local = c->nz_local;
    b.dim = c->dim;
    b.precision = c->p.precision;
    b.stride = c->p.stride;
    b.n_alloc = c->n_alloc;
    std::memset(out, 0, LBM_IPC_HANDLE_BYTES);
    std::memcpy(out, &b, sizeof b);
    return LBM_OK;
}

int lbm_ipc_attach(lbm_ctx *c, int face, const uint8_t in[LBM_IPC_HANDLE_BYTES])
{
    if (!c || !in || (face != 0 && face != 1)) return LBM_ERR_INVALID;
    if (c->peer_f[face][0]) return fail(c, LBM_ERR_STATE, "lbm_ipc_attach: face %d already has a neighbour", face);
    if ((face == 0 && c->z_begin == 0) || (face == 1 && c->z_end == c->dim))
        return fail(c, LBM_ERR_INVALID, "lbm_ipc_attach: face %d lies on the cube boundary", face);
    IpcBlob b;
    std::memcpy(&b, in, sizeof b);
    if (b.dim != c->dim || b.precision != c->p.precision || b.stride != c->p.stride)
        return fail(c, LBM_ERR_INVALID, "lbm_ipc_attach: the neighbour runs a different configuration");
    // the neighbour must store the plane I write: its halo plane next to my boundary plane
    const int my_plane = face == 0 ? c->z_begin : c->z_end - 1;
    if (my_plane < b.zs0 || my_plane >= b.zs0 + b.nz_local)
        return fail(c, LBM_ERR_INVALID, "lbm_ipc_attach: the neighbour does not store plane %d", my_plane);
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    void *p[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        const cudaError_t e = cudaIpcOpenMemHandle(&p[i], b.f[i], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            if (i == 1) cudaIpcCloseMemHandle(p[0]);
            cudaGetLastError();
            return fail(c, LBM_ERR_CUDA, "cudaIpcOpenMemHandle(%d) - %s", (int)e, cudaGetErrorName(e));
        }
    }
    c->peer_f[face][0] = p[0];
    c->peer_f[face][1] = p[1];
    c->peer_zs0[face] = b.zs0;
    c->peer_ipc[face] = true;
    return LBM_OK;
}

void lbm_default_params(lbm_params *p)
{
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->abi_version = LBM_B200_ABI_VERSION;
    p->dim = 8;            // lbm_options.hpp:35
    p->precision = LBM_F32;
    p->fast_math = 0;      // lbm_options.hpp:45
    p->viscosity = 0.0089; // lbm_options.hpp:36
    p->velocity = 0.05;    // lbm_options.hpp:37
    p->stride = 32;        // lbm_options.hpp:43
    p->block_x = 8;        // lbm_options.hpp:40-42
    p->block_y = 8;
    p->block_z = 8;
    p->device = -1;
    p->variant = LBM_VARIANT_AUTO;
    p->z_begin = 0;
    p->z_end = 0;          // 0 = whole cube
}

const char *lbm_last_error(const lbm_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

void lbm_destroy(lbm_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    for (auto &e : c->compute_events) {
        if (e.start) cudaEventDestroy(e.start);
        if (e.stop) cudaEventDestroy(e.stop);
    }
    for (int face = 0; face < 2; ++face)
        if (c->peer_ipc[face])
            for (int i = 0; i < 2; ++i)
                if (c->peer_f[face][i]) cudaIpcCloseMemHandle(c->peer_f[face][i]);
    for (int i = 0; i < 2; ++i)
        if (c->graph_exec[i]) cudaGraphExecDestroy(c->graph_exec[i]);
    if (c->bstream) cudaStreamSynchronize(c->bstream);
    if (c->comm) lbm_nccl_destroy(c->comm);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_bk[i]) cudaEventDestroy(c->ev_bk[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    }
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->bstream) cudaStreamDestroy(c->bstream);
    if (c->ev_init_start) cudaEventDestroy(c->ev_init_start);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    for (int i = 0; i < 2; ++i) {
        if (c->f[i]) cudaFree(c->f[i]);
        if (c->halo_send[i]) cudaFree(c->halo_send[i]);
        if (c->halo_recv[i]) cudaFree(c->halo_recv[i]);
    }
    if (c->rho) cudaFree(c->rho);
    if (c->u) cudaFree(c->u);
    if (c->tma_error) cudaFree(c->tma_error);
    if (c->token) cudaFree(c->token);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int lbm_create(const lbm_params *p, lbm_ctx **out)
{
    if (!out) return fail(nullptr, LBM_ERR_INVALID, "lbm_create: out is NULL");
    *out = nullptr;
    if (!p) return fail(nullptr, LBM_ERR_INVALID, "lbm_create: params is NULL");
    if (p->abi_version != LBM_B200_ABI_VERSION)
        return fail(nullptr, LBM_ERR_INVALID, "lbm_create: abi_version %d, library is %d", p->abi_version,
                    LBM_B200_ABI_VERSION);
    if (p->dim < 4 || !is_pow2(p->dim) || p->dim > 2048)
        return fail(nullptr, LBM_ERR_INVALID, "lbm_create: dim %d must be a power of two in [4, 2048]", p->dim);
    if (p->precision != LBM_F32 && p->precision != LBM_F64)
        return fail(nullptr, LBM_ERR_INVALID, "lbm_create: unknown precision %d", p->precision);
    const long long n_cube = (long long)p->dim * p->dim * p->dim;
    if (!is_pow2(p->stride) || p->stride > n_cube)
        return fail(nullptr, LBM_ERR_INVALID,
                    "lbm_create: stride %lld must be a power of two in [1, dim^3] (the reference overruns its "
                    "buffers beyond that, kernels.cl:64)", (long long)p->stride);
    if (!(p->viscosity >= 0.0) || !std::isfinite(p->viscosity) || !std::isfinite(p->velocity))
        return fail(nullptr, LBM_ERR_INVALID, "lbm_create: viscosity/velocity must be finite, viscosity >= 0");
    int zb = p->z_begin, ze = p->z_end;
    if (zb == 0 && ze == 0) ze = p->dim;
    if (zb < 0 || ze > p->dim || zb >= ze)
        return fail(nullptr, LBM_ERR_INVALID, "lbm_create: bad z range [%d, %d)", zb, ze);
    switch (p->variant) {
        case LBM_VARIANT_AUTO: case LBM_VARIANT_SCALAR: case LBM_VARIANT_VEC2: case LBM_VARIANT_VEC4: break;
        case LBM_VARIANT_TMA: break;
        case LBM_VARIANT_AA:
            if (zb != 0 || ze != p->dim)
                return fail(nullptr, LBM_ERR_INVALID, "lbm_create: the AA variant needs the whole cube on one device");
            break;
        default: return fail(nullptr, LBM_ERR_INVALID, "lbm_create: unknown variant %d", p->variant);
    }

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, LBM_ERR_NO_DEVICE, "lbm_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorName(e));

    lbm_ctx *c = new (std::nothrow) lbm_ctx();
    if (!c) return fail(nullptr, LBM_ERR_OOM, "lbm_create: host allocation failed");
    c->p = *p;
    int rc = LBM_OK;
    auto bail = [&](int code) {
        g_create_error = c->error;
        lbm_destroy(c);
        return code;
    };

    if (p->device >= 0) c->device = p->device;
    else if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
    if (c->device >= n_dev) {
        fail(c, LBM_ERR_NO_DEVICE, "lbm_create: device %d requested, %d present", c->device, n_dev);
        return bail(LBM_ERR_NO_DEVICE);
    }
    if ((rc = use_device(c)) != LBM_OK) return bail(rc);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) {
        fail(c, LBM_ERR_CUDA, "lbm_create: cudaGetDeviceProperties failed");
        return bail(LBM_ERR_CUDA);
    }
    c->device_name = prop.name;

    c->dim = p->dim;
    c->z_begin = zb;
    c->z_end = ze;
    c->zs0 = zb > 0 ? zb - 1 : 0;
    const int zs1 = ze < p->dim ? ze + 1 : p->dim;
    c->nz_local = zs1 - c->zs0;
    c->n_local = (long long)c->nz_local * p->dim * p->dim;
    c->lay.sdiv = ilog2(p->stride);
    c->lay.smod = p->stride - 1;
    c->n_alloc = ((c->n_local + p->stride - 1) / p->stride) * p->stride;
    c->esize = p->precision == LBM_F32 ? 4 : 8;
    if (p->stride <= p->dim) c->layout_mode = LM_ROWS;
    else if (p->stride >= c->n_alloc) c->layout_mode = LM_SOA;
    else c->layout_mode = LM_GENERIC;
    if (p->reserved[0] == 1) c->layout_mode = LM_GENERIC;  // test hook: force the generic addressing

    // AUTO = one cell per thread.  Measured on B200 (profiles/): with 40-48 registers the scalar kernel
    // keeps 40+ warps per SM in flight and every warp request is already a full 128-byte line, so it
    // sustains 6.5-6.9 TB/s; the 2- and 4-cell variants (64- and 128-bit accesses, x shifts by warp
    // shuffle) need 64-150 registers, drop to 16-24 warps per SM and are 3-8 % slower.  They stay
    // selectable.  Vector width is limited to 16-byte accesses inside one CSoA run.
    const int vmax = p->precision == LBM_F32 ? 4 : 2;
    int vec = (p->variant == LBM_VARIANT_AUTO || p->variant == LBM_VARIANT_SCALAR) ? 1 : p->variant;
    c->aa = p->variant == LBM_VARIANT_AA;
    if (c->aa) vec = 1;
    if (p->variant == LBM_VARIANT_TMA) {
        vec = 1;
        // eligibility; otherwise the scalar kernel is used
        c->tma = p->stride <= p->dim && p->stride * (long long)(p->precision == LBM_F32 ? 4 : 8) >= 16 && p->dim >= 32;
    }
    if (vec > vmax) vec = vmax;
    while (vec > 1 && (p->stride % vec != 0 || p->dim % vec != 0)) vec /= 2;
    c->vec = vec;
    choose_block(c);

    double eff[3];
    c->cf = make_consts<float>(p->viscosity, p->velocity, eff);
    if (p->precision == LBM_F32) { c->eff_viscosity = eff[0]; c->eff_velocity = eff[1]; c->eff_inv_tau = eff[2]; }
    c->cd = make_consts<double>(p->viscosity, p->velocity, eff);
    if (p->precision == LBM_F64) { c->eff_viscosity = eff[0]; c->eff_velocity = eff[1]; c->eff_inv_tau = eff[2]; }

    auto cuda_or_bail = [&](cudaError_t err, const char *what) {
        if (err == cudaSuccess) return false;
        fail(c, err == cudaErrorMemoryAllocation ? LBM_ERR_OOM : LBM_ERR_CUDA, "lbm_create: %s(%d) - %s", what,
             (int)err, cudaGetErrorName(err));
        return true;
    };
    if (cuda_or_bail(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        return bail(LBM_ERR_CUDA);
    c->stream = c->own_stream;
    if (cuda_or_bail(cudaEventCreate(&c->ev_init_start), "cudaEventCreate")) return bail(LBM_ERR_CUDA);
    if (cuda_or_bail(cudaEventCreate(&c->ev_last), "cudaEventCreate")) return bail(LBM_ERR_CUDA);

    const size_t f_bytes = (size_t)c->n_alloc * Q * c->esize;
    const size_t m_bytes = (size_t)c->n_local * c->esize;
    for (int i = 0; i < (c->aa ? 1 : 2); ++i) {
        cudaError_t err = cudaMalloc(&c->f[i], f_bytes);
        if (err == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            fail(c, LBM_ERR_OOM,
                 "lbm_create: cudaMalloc of lattice %d (%.1f GB) failed - cudaErrorMemoryAllocation; %s", i,
                 (double)f_bytes / 1e9,
                 c->aa ? "split the cube over more GPUs"
                       : "the in-place variant (-A, LBM_VARIANT_AA) needs one lattice instead of two, or split the "
                         "cube over more GPUs (-G N)");
            return bail(LBM_ERR_OOM);
        }
        if (cuda_or_bail(err, "cudaMalloc(f)")) return bail(LBM_ERR_CUDA);
        c->device_bytes += (int64_t)f_bytes;
    }
    {
        cudaError_t err = cudaMalloc(&c->rho, m_bytes);
        if (cuda_or_bail(err, "cudaMalloc(rho)")) return bail(err == cudaErrorMemoryAllocation ? LBM_ERR_OOM : LBM_ERR_CUDA);
        err = cudaMalloc(&c->u, 3 * m_bytes);
        if (cuda_or_bail(err, "cudaMalloc(u)")) return bail(err == cudaErrorMemoryAllocation ? LBM_ERR_OOM : LBM_ERR_CUDA);
        c->device_bytes += (int64_t)(4 * m_bytes);
    }
    const size_t h_bytes = (size_t)5 * p->dim * p->dim * c->esize;
    const bool has_face[2] = { zb > 0, ze < p->dim };
    for (int fidx = 0; fidx < 2; ++fidx) {
        if (!has_face[fidx]) continue;
        if (cuda_or_bail(cudaMalloc(&c->halo_send[fidx], h_bytes), "cudaMalloc(halo)")) return bail(LBM_ERR_OOM);
        if (cuda_or_bail(cudaMalloc(&c->halo_recv[fidx], h_bytes), "cudaMalloc(halo)")) return bail(LBM_ERR_OOM);
        c->device_bytes += (int64_t)(2 * h_bytes);
    }

    if (p->precision == LBM_F32) rc = compute_stale<float>(c, c->cf, c->stale_f);
    else rc = compute_stale<double>(c, c->cd, c->stale_d);
    if (rc != LBM_OK) return bail(rc);

    if (c->tma) {
        rc = setup_tma(c, prop.multiProcessorCount);
        if (rc != LBM_OK) return bail(rc);
    }

    *out = c;
    return LBM_OK;
}

int lbm_init(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    // a fresh profile, like a fresh LBMCL object (lbmcl.hpp:74)
    for (auto &e : c->compute_events) {
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    c->compute_events.clear();
    c->kernels_ms_accum = 0.0;
    c->launch_ms.clear();
    LBM_CUDA(c, cudaEventRecord(c->ev_init_start, c->stream));
    if (c->p.precision == LBM_F32) LBM_CUDA(c, launch_init_t<float>(c, c->cf, c->stream));
    else LBM_CUDA(c, launch_init_t<double>(c, c->cd, c->stream));
    c->cur = 0;
    c->iteration = 0;
    c->launches = 0;
    c->initialised = true;
    return record_last(c);
}

static int push_pair(lbm_ctx *c, EventPair *out)
{
    EventPair ep;
    LBM_CUDA(c, cudaEventCreate(&ep.start));
    cudaError_t e = cudaEventCreate(&ep.stop);
    if (e != cudaSuccess) {
        cudaEventDestroy(ep.start);
        return fail(c, LBM_ERR_CUDA, "cudaEventCreate(%d) - %s", (int)e, cudaGetErrorName(e));
    }
    *out = ep;
    return LBM_OK;
}

// Fold finished event pairs into the accumulator so that very long runs driven by lbm_step do not
// hold an unbounded number of CUDA events.
static int fold_events(lbm_ctx *c, bool all)
{
    if (!all && c->compute_events.size() < 4096) return LBM_OK;
    if (!c->compute_events.empty()) LBM_CUDA(c, cudaEventSynchronize(c->compute_events.back().stop));
    for (auto &e : c->compute_events) {
        float ms = 0.f;
        LBM_CUDA(c, cudaEventElapsedTime(&ms, e.start, e.stop));
        c->kernels_ms_accum += ms;
        if (c->launch_ms.size() < (1u << 20)) c->launch_ms.push_back(ms);
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    c->compute_events.clear();
    return LBM_OK;
}

int lbm_step(lbm_ctx *c, int update_macro)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_step before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    if ((rc = fold_events(c, false)) != LBM_OK) return rc;
    EventPair ep;
    if ((rc = push_pair(c, &ep)) != LBM_OK) return rc;
    c->compute_events.push_back(ep);
    LBM_CUDA(c, cudaEventRecord(ep.start, c->stream));
    if (c->comm) {
        // a slab with a communicator always goes through the exchange schedule (one iteration of it)
        if ((rc = run_slab_with_comm(c, 1, update_macro ? 1 : 0)) != LBM_OK) return rc;
    } else {
        if (c->peer_f[0][0] || c->peer_f[1][0])
            return fail(c, LBM_ERR_STATE, "lbm_step: this slab has peer neighbours; drive it through lbm_group_run");
        LBM_CUDA(c, launch_step(c, c->z_begin, c->z_end, update_macro != 0, c->stream));
        c->cur ^= 1;
        c->iteration += 1;
    }
    LBM_CUDA(c, cudaEventRecord(ep.stop, c->stream));
    return record_last(c);
}

int lbm_run(lbm_ctx *c, int n_iterations, int every)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_run before lbm_init");
    if (n_iterations < 0 || every < 0) return fail(c, LBM_ERR_INVALID, "lbm_run: negative argument");
    if (n_iterations == 0) return LBM_OK;
    if (!c->comm && (c->peer_f[0][0] || c->peer_f[1][0]))
        return fail(c, LBM_ERR_STATE, "lbm_run: this slab has peer neighbours; drive it through lbm_group_run "
                                      "(same process) or give it a communicator (lbm_comm_init)");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    if ((rc = fold_events(c, false)) != LBM_OK) return rc;
    EventPair ep;
    if ((rc = push_pair(c, &ep)) != LBM_OK) return rc;
    c->compute_events.push_back(ep);
    LBM_CUDA(c, cudaEventRecord(ep.start, c->stream));
    if (c->comm) {
        if ((rc = run_slab_with_comm(c, n_iterations, every)) != LBM_OK) return rc;
    } else {
        int left = n_iterations;
        while (left > 0) {
            const int64_t it = c->iteration + 1;  // 1-based like lbmcl.hpp:435
            // launch-bound lattices: replay a captured chunk of unflagged iterations as one graph
            // (capturing + instantiating a chunk costs a few hundred microseconds of host time: only
            // worth it when at least LBM_GRAPH_MIN_CHUNKS replays follow, or when the graph exists already)
            const int gpar = c->aa ? (int)(c->iteration & 1) : c->cur;
            const bool have_graph = c->graph_exec[gpar] != nullptr && c->graph_stream == c->stream;
            if (c->dim <= LBM_GRAPH_MAX_DIM && left >= LBM_GRAPH_CHUNK &&
                (have_graph || left >= LBM_GRAPH_MIN_CHUNKS * LBM_GRAPH_CHUNK) && c->peer_f[0][0] == nullptr &&
                c->peer_f[1][0] == nullptr) {
                const int64_t last = it + LBM_GRAPH_CHUNK - 1;
                const bool flagged = every != 0 && (last / every) != ((it - 1) / every);
                if (!flagged) {
                    if ((rc = launch_graph_chunk(c)) != LBM_OK) return rc;
                    left -= LBM_GRAPH_CHUNK;
                    continue;
                }
            }
            const bool macro = every != 0 && (it % every) == 0;
            LBM_CUDA(c, launch_step(c, c->z_begin, c->z_end, macro, c->stream));
            c->cur ^= 1;
            c->iteration = it;
            --left;
        }
    }
    LBM_CUDA(c, cudaEventRecord(ep.stop, c->stream));
    return record_last(c);
}

int lbm_sync(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->tma_error) {
        int flag = 0;
        LBM_CUDA(c, cudaMemcpy(&flag, c->tma_error, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) return fail(c, LBM_ERR_CUDA, "TMA variant: an mbarrier wait timed out (bulk copy never completed)");
    }
    return LBM_OK;
}

int lbm_read_macros(lbm_ctx *c, void *rho_host, void *u_host)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_read_macros before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long plane = (long long)c->dim * c->dim;
    const long long n_cube = plane * c->dim;
    const size_t off_local = (size_t)(c->z_begin - c->zs0) * plane * c->esize;
    const size_t off_global = (size_t)c->z_begin * plane * c->esize;
    const size_t bytes = (size_t)(c->z_end - c->z_begin) * plane * c->esize;
    if (rho_host)
        LBM_CUDA(c, cudaMemcpyAsync((char *)rho_host + off_global, (const char *)c->rho + off_local, bytes,
                                    cudaMemcpyDeviceToHost, c->stream));
    if (u_host)
        for (int k = 0; k < 3; ++k)
            LBM_CUDA(c, cudaMemcpyAsync((char *)u_host + (size_t)k * n_cube * c->esize + off_global,
                                        (const char *)c->u + (size_t)k * c->n_local * c->esize + off_local, bytes,
                                        cudaMemcpyDeviceToHost, c->stream));
    if ((rc = record_last(c)) != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    return LBM_OK;
}

int lbm_read_macros_slab(lbm_ctx *c, void *rho_slab, void *u_slab)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_read_macros_slab before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long plane = (long long)c->dim * c->dim;
    const size_t off_local = (size_t)(c->z_begin - c->zs0) * plane * c->esize;
    const size_t bytes = (size_t)(c->z_end - c->z_begin) * plane * c->esize;
    if (rho_slab)
        LBM_CUDA(c, cudaMemcpyAsync(rho_slab, (const char *)c->rho + off_local, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (u_slab)
        for (int k = 0; k < 3; ++k)
            LBM_CUDA(c, cudaMemcpyAsync((char *)u_slab + (size_t)k * bytes,
                                        (const char *)c->u + (size_t)k * c->n_local * c->esize + off_local, bytes,
                                        cudaMemcpyDeviceToHost, c->stream));
    if ((rc = record_last(c)) != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    return LBM_OK;
}

int lbm_read_map(lbm_ctx *c, int32_t *map_host)
{
    if (!c || !map_host) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long n_cube = (long long)c->dim * c->dim * c->dim;
    int *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, (size_t)n_cube * sizeof(int)));
    const int bx = c->dim < 64 ? c->dim : 64;
    const int by = (256 / bx) < c->dim ? (256 / bx) : c->dim;
    map_kernel<<<dim3(c->dim / bx, c->dim / by, c->dim), dim3(bx, by, 1), 0, c->stream>>>(d, c->dim);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(map_host, d, (size_t)n_cube * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    LBM_CUDA(c, e);
    return record_last(c);
}

int lbm_read_f(lbm_ctx *c, void *f_host)
{
    if (!c || !f_host) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_read_f before lbm_init");
    if (c->z_begin != 0 || c->z_end != c->dim)
        return fail(c, LBM_ERR_INVALID, "lbm_read_f: only on a context that owns the whole cube");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long n_cube = (long long)c->dim * c->dim * c->dim;
    const size_t bytes = (size_t)n_cube * Q * c->esize;
    void *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, bytes));
    const int bx = c->dim < 64 ? c->dim : 64;
    const int by = (256 / bx) < c->dim ? (256 / bx) : c->dim;
    const dim3 b(bx, by, 1), g(c->dim / bx, c->dim / by, c->dim);
    const int pristine = c->iteration == 0 ? 1 : 0;
    if (c->aa) {
        const int swapped = (c->iteration % 2) == 1 ? 1 : 0;
        if (c->p.precision == LBM_F32)
            reference_view_aa_kernel<float><<<g, b, 0, c->stream>>>((const float *)c->f[0], (float *)d, c->dim, c->lay,
                                                                    c->cf, swapped, pristine);
        else
            reference_view_aa_kernel<double><<<g, b, 0, c->stream>>>((const double *)c->f[0], (double *)d, c->dim,
                                                                     c->lay, c->cd, swapped, pristine);
    } else if (c->p.precision == LBM_F32)
        reference_view_kernel<float><<<g, b, 0, c->stream>>>((const float *)c->f[c->cur], (float *)d, c->dim, c->zs0,
                                                             c->nz_local, 0, c->dim, c->lay, c->lay, c->cf, pristine);
    else
        reference_view_kernel<double><<<g, b, 0, c->stream>>>((const double *)c->f[c->cur], (double *)d, c->dim,
                                                              c->zs0, c->nz_local, 0, c->dim, c->lay, c->lay, c->cd,
                                                              pristine);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(f_host, d, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    LBM_CUDA(c, e);
    return record_last(c);
}

int lbm_time_ms(lbm_ctx *c, double *total_ms, double *kernels_ms)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_time_ms before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    LBM_CUDA(c, cudaEventSynchronize(c->ev_last));
    if (total_ms) {
        float ms = 0.f;
        LBM_CUDA(c, cudaEventElapsedTime(&ms, c->ev_init_start, c->ev_last));
        *total_ms = ms;
    }
    if (kernels_ms) {
        if ((rc = fold_events(c, true)) != LBM_OK) return rc;
        *kernels_ms = c->kernels_ms_accum;
    }
    return LBM_OK;
}

int lbm_launch_times_ms(lbm_ctx *c, double *out, int64_t capacity, int64_t *count)
{
    if (!c || !count) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_launch_times_ms before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    if ((rc = fold_events(c, true)) != LBM_OK) return rc;
    *count = (int64_t)c->launch_ms.size();
    if (out)
        for (int64_t i = 0; i < capacity && i < *count; ++i) out[i] = c->launch_ms[(size_t)i];
    return LBM_OK;
}

int lbm_device_name(const lbm_ctx *c, char *buf, size_t buflen)
{
    if (!c || !buf || buflen == 0) return LBM_ERR_INVALID;
    snprintf(buf, buflen, "%s", c->device_name.c_str());
    return LBM_OK;
}

int lbm_effective_params(const lbm_ctx *c, double out[3])
{
    if (!c || !out) return LBM_ERR_INVALID;
    out[0] = c->eff_viscosity;
    out[1] = c->eff_velocity;
    out[2] = c->eff_inv_tau;
    return LBM_OK;
}

int lbm_block_shape(const lbm_ctx *c, int32_t block[3], int32_t *cells_per_thread)
{
    if (!c) return LBM_ERR_INVALID;
    if (block) {
        block[0] = (int32_t)c->block.x;
        block[1] = (int32_t)c->block.y;
        block[2] = (int32_t)c->block.z;
    }
    if (cells_per_thread) *cells_per_thread = c->vec;
    return LBM_OK;
}

int64_t lbm_device_bytes(const lbm_ctx *c) { return c ? c->device_bytes : 0; }
int64_t lbm_launch_count(const lbm_ctx *c) { return c ? c->launches : 0; }
int64_t lbm_iteration(const lbm_ctx *c) { return c ? c->iteration : 0; }

int lbm_set_stream(lbm_ctx *c, void *cuda_stream)
{
    if (!c) return LBM_ERR_INVALID;
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return LBM_OK;
}

int lbm_step_planes(lbm_ctx *c, int z_begin, int z_end, int update_macro)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_step_planes before lbm_init");
    if (c->aa) return fail(c, LBM_ERR_INVALID, "lbm_step_planes: not available with the AA variant");
    if (z_begin < c->z_begin || z_end > c->z_end || z_begin > z_end)
        return fail(c, LBM_ERR_INVALID, "lbm_step_planes: [%d, %d) outside the owned planes [%d, %d)", z_begin,
                    z_end, c->z_begin, c->z_end);
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, launch_step(c, z_begin, z_end, update_macro != 0, c->stream));
    return LBM_OK;
}

int lbm_advance(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_advance before lbm_init");
    c->cur ^= 1;
    c->iteration += 1;
    return LBM_OK;
}

int lbm_z_range(const lbm_ctx *c, int32_t *z_begin, int32_t *z_end)
{
    if (!c) return LBM_ERR_INVALID;
    if (z_begin) *z_begin = c->z_begin;
    if (z_end) *z_end = c->z_end;
    return LBM_OK;
}

// ---- dense halo transport (one process per device) ----

int64_t lbm_halo_elems(const lbm_ctx *c) { return c ? (int64_t)5 * c->dim * c->dim : 0; }
void *lbm_halo_send_buffer(lbm_ctx *c, int face) { return (c && (face == 0 || face == 1)) ? c->halo_send[face] : nullptr; }
void *lbm_halo_recv_buffer(lbm_ctx *c, int face) { return (c && (face == 0 || face == 1)) ? c->halo_recv[face] : nullptr; }

}  // extern "C"

template <typename T, bool PACK>
static cudaError_t halo_launch(lbm_ctx *c, void *lattice, void *dense, long long plane_local, int dir_up)
{
    const int bx = c->dim < 256 ? c->dim : 256;
    const dim3 b(bx, 1, 1), g(c->dim / bx, c->dim, 5);
    halo_kernel<T, PACK><<<g, b, 0, c->stream>>>((T *)lattice, (T *)dense, c->dim, plane_local, c->lay, dir_up);
    c->launches += 1;
    return cudaGetLastError();
}

extern "C" {

// Packs, from the lattice the NEXT iteration reads (i.e. the one just written), what the neighbours
// will gather: low face -> populations with e_z = -1 of plane z_begin; high face -> e_z = +1 of plane
// z_end - 1.
int lbm_halo_pack(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    void *lat = c->f[c->cur];
    const bool f32 = c->p.precision == LBM_F32;
    if (c->halo_send[0]) {
        const long long pl = c->z_begin - c->zs0;
        LBM_CUDA(c, f32 ? halo_launch<float, true>(c, lat, c->halo_send[0], pl, 0)
                        : halo_launch<double, true>(c, lat, c->halo_send[0], pl, 0));
    }
    if (c->halo_send[1]) {
        const long long pl = (c->z_end - 1) - c->zs0;
        LBM_CUDA(c, f32 ? halo_launch<float, true>(c, lat, c->halo_send[1], pl, 1)
                        : halo_launch<double, true>(c, lat, c->halo_send[1], pl, 1));
    }
    return LBM_OK;
}

// Scatters what the neighbours packed into the halo planes of the lattice the NEXT iteration reads:
// low halo plane (z_begin - 1) receives e_z = +1 populations, high halo plane (z_end) e_z = -1.
int lbm_halo_unpack(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    void *lat = c->f[c->cur];
    const bool f32 = c->p.precision == LBM_F32;
    if (c->halo_recv[0]) {
        const long long pl = (c->z_begin - 1) - c->zs0;
        LBM_CUDA(c, f32 ? halo_launch<float, false>(c, lat, c->halo_recv[0], pl, 1)
                        : halo_launch<double, false>(c, lat, c->halo_recv[0], pl, 1));
    }
    if (c->halo_recv[1]) {
        const long long pl = c->z_end - c->zs0;
        LBM_CUDA(c, f32 ? halo_launch<float, false>(c, lat, c->halo_recv[1], pl, 0)
                        : halo_launch<double, false>(c, lat, c->halo_recv[1], pl, 0));
    }
    return LBM_OK;
}

}  // extern "C"

#include "lbm_group.inl"
