// C ABI of the B200-native D3Q19 collide-and-stream path (see include/lbm_b200.h).
//
// This file replaces what the reference's lbmcl.hpp obtains from libs/CLUtil.hpp + cl.hpp: device
// selection, buffers, kernel launches with the ping-pong binding, readbacks and event timing -- plus the
// z-slab transports (new functionality).  Host logic only: the kernels are instantiated in the
// lbm_launch_*.cu translation units.  There is no CPU fallback anywhere in here: without a CUDA device
// lbm_create fails.
#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>

#include "lbm_ctx.hpp"

using namespace lbm;

namespace {

thread_local std::string g_create_error;
using EventPair = LbmEventPair;

}  // namespace

int lbm_fail(lbm_ctx *ctx, int code, const char *fmt, ...)
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

static void lbm_nccl_destroy(ncclComm_t comm);
static int setup_tma(lbm_ctx *c, int n_sm);

namespace {

#define fail lbm_fail

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

// planes of one launch, in grid order: `n_named` individually named planes, then [z_begin, z_end)
struct Planes {
    int n_named = 0;
    int named[2] = {0, 0};
    int z_begin = 0, z_end = 0;
    static Planes range(int zb, int ze)
    {
        Planes p;
        p.z_begin = zb;
        p.z_end = ze > zb ? ze : zb;
        return p;
    }
    int count() const { return n_named + (z_end - z_begin); }
};

// the boundary planes of a slab (those with a neighbour), deduplicated, and its interior range
Planes boundary_planes(const lbm_ctx *c)
{
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    const int zlo = c->z_begin, zhi = c->z_end - 1;
    Planes p;
    if (has_lo) p.named[p.n_named++] = zlo;
    if (has_hi && !(has_lo && zhi == zlo)) p.named[p.n_named++] = zhi;
    return p;
}
Planes interior_planes(const lbm_ctx *c)
{
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    return Planes::range(c->z_begin + (has_lo ? 1 : 0), c->z_end - (has_hi ? 1 : 0));
}

SlabSync make_sync(const lbm_ctx *c, unsigned wait_epoch, unsigned signal_epoch)
{
    SlabSync y{};
    y.flag_in = reinterpret_cast<unsigned *>(static_cast<char *>(c->f[0]) + c->flag_off);
    y.flag_out[0] = c->peer_flag[0];
    y.flag_out[1] = c->peer_flag[1];
    y.count = c->sync_local;
    y.error = reinterpret_cast<int *>(c->sync_local + 2);
    y.wait_epoch = wait_epoch;
    y.signal_epoch = signal_epoch;
    y.timeout_ns = c->sync_timeout_ns;
    return y;
}

template <typename T>
StepArgs<T> make_step_args(lbm_ctx *c, const Planes &pl, int peer_mode, const Consts<T> &k, const T (&stale)[2][Q])
{
    StepArgs<T> a{};
    a.dst = static_cast<T *>(c->f[c->cur ^ 1]);
    a.src = static_cast<const T *>(c->f[c->cur]);
    a.rho = static_cast<T *>(c->rho);
    a.u = static_cast<T *>(c->u);
    a.peer_lo = nullptr;
    a.peer_hi = nullptr;
    // my first / last owned plane is the neighbour's high / low halo plane, in the lattice it reads
    // next (all slabs advance in lock step, so the neighbour's lattice index equals mine).  Launches
    // that do not hand populations to a neighbour (interior planes, dense-halo transports) get none.
    if (peer_mode != PEER_NONE) {
        if (c->peer_f[0][0]) {
            a.peer_lo = static_cast<T *>(c->peer_f[0][c->cur ^ 1]);
            a.peer_lo_plane = c->z_begin - c->peer_zs0[0];
        }
        if (c->peer_f[1][0]) {
            a.peer_hi = static_cast<T *>(c->peer_f[1][c->cur ^ 1]);
            a.peer_hi_plane = (c->z_end - 1) - c->peer_zs0[1];
        }
    }
    a.z_own_begin = c->z_begin;
    a.z_own_end = c->z_end;
    a.dim = c->dim;
    a.zs0 = c->zs0;
    a.zmap_n = pl.n_named;
    a.zmap0 = pl.named[0];
    a.zmap1 = pl.named[1];
    a.z_begin = pl.z_begin;
    a.z_end = pl.z_end;
    a.n_local = c->n_local;
    a.lay = c->lay;
    // tile order of the blocks: only for launches over one contiguous plane range whose grid the tiles divide
    a.swz_y = a.swz_z = -1;
    a.swz_nty = 0;
    const bool shift_step = c->aa && ((c->iteration + 1) % 2) == 0;
    if (c->swz_z >= 0 && pl.n_named == 0 && (!c->swz_shift_only || shift_step)) {
        const int ny = c->dim / (int)c->block.y, planes = pl.z_end - pl.z_begin;
        const int nzb = planes / (int)c->block.z;
        if (planes % (int)c->block.z == 0 && ny % (1 << c->swz_y) == 0 && nzb > 0 && nzb % (1 << c->swz_z) == 0) {
            a.swz_y = c->swz_y;
            a.swz_z = c->swz_z;
            a.swz_nty = ilog2(ny >> c->swz_y);
        }
    }
    a.row_shift = c->lay.sdiv > ilog2(c->dim) ? c->lay.sdiv - ilog2(c->dim) : 0;
    a.row_mask = (1 << a.row_shift) - 1;
    a.blk18 = 18ll * c->lay.qpitch() * (long long)sizeof(T);
    a.c = k;
    for (int i = 0; i < 2; ++i)
        for (int q = 0; q < Q; ++q) a.stale[i][q] = stale[i][q];
    const int lm = peer_mode != PEER_NONE ? c->layout_natural : c->layout_mode;
    const long long S = c->lay.qpitch(), dim = c->dim, plane = dim * dim, es = (long long)sizeof(T);
    for (int q = 0; q < Q; ++q) {
        a.soff[q] = q * S * es;
        const long long dcell = (long long)ey(q) * dim + (long long)ez(q) * plane;  // cells between the two rows
        const long long rowmul = lm == LM_ROWS ? Q : 1;                             // CSoA rows hold Q values per cell
        if (lm == LM_GENERIC) {
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

LaunchCfg launch_cfg(const lbm_ctx *c, int peer_mode)
{
    LaunchCfg k;
    k.block = c->block;
    k.dim = c->dim;
    k.lm = peer_mode != PEER_NONE ? c->layout_natural : c->layout_mode;
    k.fast = c->p.fast_math != 0;
    return k;
}

cudaError_t launch_pull(lbm_ctx *c, const StepArgs<float> &a, bool macro, int peer, cudaStream_t s)
{
    const LaunchCfg k = launch_cfg(c, peer);
    switch (c->vec) {
        case 4: return launch_pull_f32_v4(k, a, macro, peer, s);
        case 2: return launch_pull_f32_v2(k, a, macro, peer, s);
        default: return launch_pull_f32_v1(k, a, macro, peer, s);
    }
}
cudaError_t launch_pull(lbm_ctx *c, const StepArgs<double> &a, bool macro, int peer, cudaStream_t s)
{
    const LaunchCfg k = launch_cfg(c, peer);
    switch (c->vec) {
        case 2: return launch_pull_f64_v2(k, a, macro, peer, s);
        default: return launch_pull_f64_v1(k, a, macro, peer, s);
    }
}
cudaError_t launch_aa(lbm_ctx *c, const StepArgs<float> &a, bool macro, bool shift, cudaStream_t s)
{
    return launch_aa_f32(launch_cfg(c, PEER_NONE), a, macro, shift, s);
}
cudaError_t launch_aa(lbm_ctx *c, const StepArgs<double> &a, bool macro, bool shift, cudaStream_t s)
{
    return launch_aa_f64(launch_cfg(c, PEER_NONE), a, macro, shift, s);
}
cudaError_t launch_tma(const TmaCfg &k, const StepArgs<float> &a, int ns, bool macro, cudaStream_t s)
{
    return launch_tma_f32(k, a, ns, macro, s);
}
cudaError_t launch_tma(const TmaCfg &k, const StepArgs<double> &a, int ns, bool macro, cudaStream_t s)
{
    return launch_tma_f64(k, a, ns, macro, s);
}

template <typename T>
cudaError_t launch_step_p(lbm_ctx *c, const Planes &pl, bool macro, int peer_mode, const SlabSync *sync, cudaStream_t s,
                          const Consts<T> &k, const T (&stale)[2][Q])
{
    if (pl.count() <= 0) return cudaSuccess;
    StepArgs<T> a = make_step_args<T>(c, pl, peer_mode, k, stale);
    if (sync) a.sync = *sync;
    // a kernel that overwrites rho / u must not overtake an asynchronous read-back of them
    if (macro && c->copy_pending) {
        const cudaError_t e = cudaStreamWaitEvent(s, c->ev_copy_done, 0);
        if (e != cudaSuccess) return e;
    }
    c->launches += 1;
    if (c->aa) return launch_aa(c, a, macro, ((c->iteration + 1) % 2) == 0, s);
    const bool plain = peer_mode == PEER_NONE && pl.n_named == 0;
    if (c->spec && plain) return lbm_nvrtc_launch(c, &a, macro, pl.count(), s);
    if (c->tma && plain) {
        TmaCfg t{};
        t.map_src = &c->tmap[c->cur];
        t.map_dst = &c->tmap_st[c->cur ^ 1];
        t.osdiv = c->tma_osdiv;
        t.tx = c->tma_tx;
        t.grid = c->tma_grid;
        t.smem = c->tma_smem;
        t.error = c->tma_error;
        t.fast = c->p.fast_math != 0;
        return launch_tma(t, a, c->tma_ns, macro, s);
    }
    return launch_pull(c, a, macro, peer_mode, s);
}

// one launch of the step kernel over the given planes
cudaError_t launch_step(lbm_ctx *c, const Planes &pl, bool macro, int peer_mode, cudaStream_t s,
                        const SlabSync *sync = nullptr)
{
    if (c->p.precision == LBM_F32) return launch_step_p<float>(c, pl, macro, peer_mode, sync, s, c->cf, c->stale_f);
    return launch_step_p<double>(c, pl, macro, peer_mode, sync, s, c->cd, c->stale_d);
}

cudaError_t launch_init(lbm_ctx *c, cudaStream_t s)
{
    if (c->p.precision == LBM_F32) {
        InitArgs<float> a{};
        a.f0 = static_cast<float *>(c->f[0]);
        a.f1 = static_cast<float *>(c->f[1]);
        a.rho = static_cast<float *>(c->rho);
        a.u = static_cast<float *>(c->u);
        a.dim = c->dim;
        a.zs0 = c->zs0;
        a.nz_local = c->nz_local;
        a.n_local = c->n_local;
        a.lay = c->lay;
        a.c = c->cf;
        return launch_init_f32(a, c->aa, s);
    }
    InitArgs<double> a{};
    a.f0 = static_cast<double *>(c->f[0]);
    a.f1 = static_cast<double *>(c->f[1]);
    a.rho = static_cast<double *>(c->rho);
    a.u = static_cast<double *>(c->u);
    a.dim = c->dim;
    a.zs0 = c->zs0;
    a.nz_local = c->nz_local;
    a.n_local = c->n_local;
    a.lay = c->lay;
    a.c = c->cd;
    return launch_init_f64(a, c->aa, s);
}

template <typename T>
int compute_stale(lbm_ctx *c, const Consts<T> &k, T (&stale)[2][Q])
{
    T *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, sizeof(T) * 2 * Q));
    cudaError_t e;
    if constexpr (sizeof(T) == 4) e = launch_stale_f32(k, d, c->stream);
    else e = launch_stale_f64(k, d, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&stale[0][0], d, sizeof(T) * 2 * Q, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    LBM_CUDA(c, e);
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

bool has_neighbours(const lbm_ctx *c) { return c->peer_f[0][0] != nullptr || c->peer_f[1][0] != nullptr; }

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
            e = launch_step(c, Planes::range(c->z_begin, c->z_end), false, PEER_NONE, c->stream);
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

// The overlapped z-slab schedule of one rank with a communicator (include/lbm_b200.h, transports 2b / 2d).
//   bstream: wait interior(k-1) -> boundary planes(k), ONE launch -> [advance] pack -> send/recv -> unpack
//   stream : wait boundary kernels(k-1) -> interior planes(k)
// The events alternate with the iteration parity; everything is enqueued without host synchronisation.
int run_slab_with_comm(lbm_ctx *c, int n_iterations, int every)
{
    NcclApi &n = nccl();
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    const ncclDataType_t dt = c->p.precision == LBM_F32 ? ncclFloat32 : ncclFloat64;
    const size_t count = (size_t)5 * c->dim * c->dim;
    cudaStream_t S = c->stream, B = c->bstream;
    // token form: every interior face has an attached neighbour, the crossing populations are peer stores
    const bool fused = c->fused;
    const Planes bp = boundary_planes(c), ip = interior_planes(c);
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
        LBM_CUDA(c, cudaStreamWaitEvent(B, c->ev_in[prev], 0));
        LBM_CUDA(c, launch_step(c, bp, macro, fused ? PEER_STORE : PEER_NONE, B));
        LBM_CUDA(c, cudaEventRecord(c->ev_bk[par], B));

        LBM_CUDA(c, cudaStreamWaitEvent(S, c->ev_bk[prev], 0));
        LBM_CUDA(c, launch_step(c, ip, macro, PEER_NONE, S));
        LBM_CUDA(c, cudaEventRecord(c->ev_in[par], S));

        c->cur ^= 1;
        c->iteration = it;

        if (fused) {
            // the boundary kernel has already stored the crossing populations in the neighbours' halo
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

// The z-slab schedule with in-kernel epoch flags (transport 2c): ONE launch per iteration and rank on one
// stream.  The grid starts with the slab's boundary planes: their blocks wait (bounded) for the neighbours'
// previous phase, store the crossing populations straight into the neighbours' halo planes and the last
// block of each face publishes the new phase; the interior planes follow in the same grid and overlap the
// NVLink traffic.  Ranks can drift apart by at most one iteration; nothing else synchronises them.
int run_slab_flags(lbm_ctx *c, int n_iterations, int every)
{
    Planes all = boundary_planes(c);
    const Planes ip = interior_planes(c);
    all.z_begin = ip.z_begin;
    all.z_end = ip.z_end;
    for (int i = 0; i < n_iterations; ++i) {
        const int64_t it = c->iteration + 1;
        const bool macro = every != 0 && (it % every) == 0;
        c->phase += 1;
        const SlabSync y = make_sync(c, c->phase - 1, c->phase);
        LBM_CUDA(c, launch_step(c, all, macro, PEER_FLAGS, c->stream, &y));
        c->cur ^= 1;
        c->iteration = it;
    }
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
    // cells per row tile = consumer threads per CTA (measured: 256 beats 128 and 64 in both precisions)
    c->tma_tx = c->dim < 256 ? c->dim : 256;
    if (const char *e = std::getenv("LBM_TMA_TX")) {  // test / tuning hook: other tile widths (several segments per row)
        const int v = std::atoi(e);
        if ((v == 32 || v == 64 || v == 128 || v == 256) && v <= c->dim) c->tma_tx = v;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)S, (cuuint64_t)Q, (cuuint64_t)(c->n_alloc / S)};
    const cuuint64_t gstride[2] = {(cuuint64_t)(S * c->esize), (cuuint64_t)(Q * S * c->esize)};
    const cuuint32_t b0 = (cuuint32_t)(S < c->tma_tx ? S : c->tma_tx);
    const cuuint32_t box[3] = {b0, 1u, (cuuint32_t)(c->tma_tx / b0)};        // loads: one direction of one row tile
    const cuuint32_t s32 = (cuuint32_t)(S < 32 ? S : 32);
    const cuuint32_t box_st[3] = {s32, (cuuint32_t)Q, 32u / s32};             // stores: one warp's cells, all directions
    c->tma_osdiv = ilog2(s32);
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = c->p.precision == LBM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    for (int i = 0; i < 2; ++i) {
        CUresult r = encode(&c->tmap[i], dt, 3, c->f[i], gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = encode(&c->tmap_st[i], dt, 3, c->f[i], gdim, gstride, box_st, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, LBM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    // shared memory of one CTA: NS input tiles [Q][TX], (bulk-store build only) two output tiles [Q][32] per
    // consumer warp, 2 NS mbarriers + NS tile coordinates.  The producer warp keeps NS - 1 tiles in flight ahead of
    // the consumers.  Measured (profiles/r02_tma_experiment.md): 2, 3 or 4 stages and 2, 3 or 4 CTAs per SM all land
    // within 5 % of each other (the kernel is bound by instruction issue, not by latency), 2 stages marginally best.
    const size_t stage = (size_t)Q * c->tma_tx * c->esize;
    const size_t outs = tma_direct_store() ? 0 : (size_t)(c->tma_tx / 32) * 2 * Q * 32 * c->esize;
    const size_t fixed = outs + 256;
    int ns = 2;
    if (const char *e = std::getenv("LBM_TMA_NS")) ns = std::atoi(e);
    if (ns < 2) ns = 2;
    while (ns > 2 && (size_t)ns * stage + fixed > 200 * 1024) --ns;
    c->tma_ns = ns;
    c->tma_smem = (size_t)ns * stage + outs + (size_t)ns * (2 * sizeof(uint64_t) + 16) + 16;
    if (c->tma_smem > 200 * 1024) return fail(c, LBM_ERR_INVALID, "TMA variant: tile ring does not fit shared memory");
    LBM_CUDA(c, cudaMalloc(&c->tma_error, sizeof(int)));
    LBM_CUDA(c, cudaMemset(c->tma_error, 0, sizeof(int)));
    // opt in to the large dynamic shared memory once per device (never inside a stream capture)
    LBM_CUDA(c, tma_prepare(c->device));
    // one wave of persistent CTAs: as many per SM as registers and shared memory allow
    int ctas = tma_resident_ctas(c->p.precision == LBM_F64, c->tma_tx, c->tma_smem, c->p.fast_math != 0);
    if (const char *e = std::getenv("LBM_TMA_CTAS")) ctas = std::atoi(e) > 0 ? std::atoi(e) : ctas;
    c->tma_grid = ctas * n_sm;
    return LBM_OK;
}

static void lbm_nccl_destroy(ncclComm_t comm)
{
    if (comm && nccl().CommDestroy) nccl().CommDestroy(comm);
}

// stream / events of the two-stream slab schedules
static int ensure_boundary_stream(lbm_ctx *c)
{
    if (c->bstream) return LBM_OK;
    int lo = 0, hi = 0;
    LBM_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LBM_CUDA(c, cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; ++i) {
        LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_bk[i], cudaEventDisableTiming));
        LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    }
    LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    return LBM_OK;
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
    if ((rc = ensure_boundary_stream(c)) != LBM_OK) return rc;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    LBM_NCCL(c, n.CommInitRank(&c->comm, world, u, rank));
    c->comm_rank = rank;
    c->comm_world = world;
    LBM_CUDA(c, cudaMalloc(&c->token, 3 * sizeof(int)));
    LBM_CUDA(c, cudaMemset(c->token, 0, 3 * sizeof(int)));
    if (c->sync_mode == LBM_SYNC_NONE) c->sync_mode = LBM_SYNC_NCCL;  // dense halos until lbm_comm_fused says otherwise
    return LBM_OK;
}

int lbm_comm_fused(lbm_ctx *c, int mode)
{
    if (!c) return LBM_ERR_INVALID;
    const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
    const bool attached = (!has_lo || c->peer_f[0][0]) && (!has_hi || c->peer_f[1][0]);
    switch (mode) {
        case LBM_FUSED_OFF:
            c->fused = false;
            c->sync_mode = c->comm ? LBM_SYNC_NCCL : LBM_SYNC_NONE;
            return LBM_OK;
        case LBM_FUSED_FLAGS:
            if (!attached)
                return fail(c, LBM_ERR_STATE, "lbm_comm_fused: an interior face has no attached neighbour (lbm_ipc_attach)");
            c->fused = has_lo || has_hi;
            c->sync_mode = c->fused ? LBM_SYNC_FLAGS : LBM_SYNC_NONE;
            c->initialised = false;  // the phase protocol starts with lbm_init on every rank
            return LBM_OK;
        case LBM_FUSED_TOKEN:
            if (!c->comm) return fail(c, LBM_ERR_STATE, "lbm_comm_fused: the token form needs a communicator (lbm_comm_init)");
            if (!attached)
                return fail(c, LBM_ERR_STATE, "lbm_comm_fused: an interior face has no attached neighbour (lbm_ipc_attach)");
            c->fused = has_lo || has_hi;
            c->sync_mode = LBM_SYNC_NCCL;
            return LBM_OK;
        default: return fail(c, LBM_ERR_INVALID, "lbm_comm_fused: unknown mode %d", mode);
    }
}

// ---- CUDA IPC: let a neighbouring PROCESS's boundary kernel store straight into this lattice ----
namespace {
struct IpcBlob {
    cudaIpcMemHandle_t f[2];
    int32_t zs0, nz_local, dim, precision;
    int64_t stride, n_alloc;
    int32_t z_begin, z_end;   // owned planes: the attaching side checks that the slabs are adjacent
    int64_t flag_off;         // byte offset of the two incoming flag words behind lattice 0
};
static_assert(sizeof(IpcBlob) <= LBM_IPC_HANDLE_BYTES, "IpcBlob must fit LBM_IPC_HANDLE_BYTES");

// geometry checks shared by lbm_ipc_attach and lbm_peer_attach
int check_neighbour(lbm_ctx *c, int face, const char *who, int dim, int precision, int64_t stride, int zs0, int nz_local,
                    int z_begin, int z_end)
{
    if (c->peer_f[face][0]) return fail(c, LBM_ERR_STATE, "%s: face %d already has a neighbour", who, face);
    if ((face == 0 && c->z_begin == 0) || (face == 1 && c->z_end == c->dim))
        return fail(c, LBM_ERR_INVALID, "%s: face %d lies on the cube boundary", who, face);
    if (c->aa) return fail(c, LBM_ERR_INVALID, "%s: the AA variant is single-device", who);
    if (dim != c->dim || precision != c->p.precision || stride != c->p.stride)
        return fail(c, LBM_ERR_INVALID, "%s: the neighbour runs a different configuration", who);
    // the slabs must be adjacent: my boundary plane is the neighbour's HALO plane, not one it owns
    if ((face == 0 && z_end != c->z_begin) || (face == 1 && z_begin != c->z_end))
        return fail(c, LBM_ERR_INVALID, "%s: the neighbour owns planes [%d, %d), not adjacent to face %d of [%d, %d)", who,
                    z_begin, z_end, face, c->z_begin, c->z_end);
    const int my_plane = face == 0 ? c->z_begin : c->z_end - 1;
    if (my_plane < zs0 || my_plane >= zs0 + nz_local)
        return fail(c, LBM_ERR_INVALID, "%s: the neighbour does not store plane %d", who, my_plane);
    return LBM_OK;
}
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
    b.z_begin = c->z_begin;
    b.z_end = c->z_end;
    b.flag_off = (int64_t)c->flag_off;
    std::memset(out, 0, LBM_IPC_HANDLE_BYTES);
    std::memcpy(out, &b, sizeof b);
    return LBM_OK;
}

int lbm_ipc_attach(lbm_ctx *c, int face, const uint8_t in[LBM_IPC_HANDLE_BYTES])
{
    if (!c || !in || (face != 0 && face != 1)) return LBM_ERR_INVALID;
    IpcBlob b;
    std::memcpy(&b, in, sizeof b);
    int rc = check_neighbour(c, face, "lbm_ipc_attach", b.dim, b.precision, b.stride, b.zs0, b.nz_local, b.z_begin, b.z_end);
    if (rc != LBM_OK) return rc;
    if ((rc = use_device(c)) != LBM_OK) return rc;
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
    // I am the neighbour's high neighbour when it sits on my low face, and vice versa
    c->peer_flag[face] = reinterpret_cast<unsigned *>(static_cast<char *>(p[0]) + b.flag_off) + (face == 0 ? 1 : 0);
    c->peer_ipc[face] = true;
    return LBM_OK;
}

int lbm_peer_attach(lbm_ctx *c, int face, lbm_ctx *nb)
{
    if (!c || !nb || c == nb || (face != 0 && face != 1)) return LBM_ERR_INVALID;
    int rc = check_neighbour(c, face, "lbm_peer_attach", nb->dim, nb->p.precision, nb->p.stride, nb->zs0, nb->nz_local,
                             nb->z_begin, nb->z_end);
    if (rc != LBM_OK) return rc;
    if (nb->aa) return fail(c, LBM_ERR_INVALID, "lbm_peer_attach: the AA variant is single-device");
    if (nb->device != c->device) {
        int can = 0;
        if ((rc = use_device(c)) != LBM_OK) return rc;
        if (cudaDeviceCanAccessPeer(&can, c->device, nb->device) != cudaSuccess || !can)
            return fail(c, LBM_ERR_CUDA, "lbm_peer_attach: device %d cannot access peer %d", c->device, nb->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else LBM_CUDA(c, e);
    }
    c->peer[face] = nb;
    c->peer_f[face][0] = nb->f[0];
    c->peer_f[face][1] = nb->f[1];
    c->peer_zs0[face] = nb->zs0;
    c->peer_flag[face] = reinterpret_cast<unsigned *>(static_cast<char *>(nb->f[0]) + nb->flag_off) + (face == 0 ? 1 : 0);
    c->peer_ipc[face] = false;
    return LBM_OK;
}

int lbm_ipc_detach(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    // nothing of mine may still be storing into the neighbours
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->bstream) LBM_CUDA(c, cudaStreamSynchronize(c->bstream));
    for (int face = 0; face < 2; ++face) {
        if (c->peer_ipc[face])
            for (int i = 0; i < 2; ++i)
                if (c->peer_f[face][i]) cudaIpcCloseMemHandle(c->peer_f[face][i]);
        c->peer[face] = nullptr;
        c->peer_f[face][0] = c->peer_f[face][1] = nullptr;
        c->peer_flag[face] = nullptr;
        c->peer_ipc[face] = false;
    }
    c->fused = false;
    c->sync_mode = c->comm ? LBM_SYNC_NCCL : LBM_SYNC_NONE;
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
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (auto &e : c->compute_events) {
        if (e.start) cudaEventDestroy(e.start);
        if (e.stop) cudaEventDestroy(e.stop);
    }
    if (c->bstream) cudaStreamSynchronize(c->bstream);
    for (int face = 0; face < 2; ++face)
        if (c->peer_ipc[face])
            for (int i = 0; i < 2; ++i)
                if (c->peer_f[face][i]) cudaIpcCloseMemHandle(c->peer_f[face][i]);
    for (int i = 0; i < 2; ++i)
        if (c->graph_exec[i]) cudaGraphExecDestroy(c->graph_exec[i]);
    if (c->comm) lbm_nccl_destroy(c->comm);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_bk[i]) cudaEventDestroy(c->ev_bk[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    }
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->bstream) cudaStreamDestroy(c->bstream);
    if (c->ev_copy_ready) cudaEventDestroy(c->ev_copy_ready);
    if (c->ev_copy_done) cudaEventDestroy(c->ev_copy_done);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_init_start) cudaEventDestroy(c->ev_init_start);
    if (c->ev_last) cudaEventDestroy(c->ev_last);
    lbm_nvrtc_destroy(c);
    for (int i = 0; i < 2; ++i) {
        if (c->f[i]) cudaFree(c->f[i]);
        if (c->halo_send[i]) cudaFree(c->halo_send[i]);
        if (c->halo_recv[i]) cudaFree(c->halo_recv[i]);
    }
    if (c->rho) cudaFree(c->rho);
    if (c->u) cudaFree(c->u);
    if (c->tma_error) cudaFree(c->tma_error);
    if (c->sync_local) cudaFree(c->sync_local);
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
        case LBM_VARIANT_TMA: case LBM_VARIANT_NVRTC: break;
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
    // every stride is a power of two, so exactly one of the three uniform-offset modes applies
    if (p->stride <= p->dim) c->layout_natural = LM_ROWS;
    else if (p->stride >= c->n_alloc) c->layout_natural = LM_SOA;
    else c->layout_natural = LM_BLOCKROWS;
    c->layout_mode = c->layout_natural;
    if (p->reserved[0] == 1) c->layout_mode = LM_GENERIC;  // test hook: force the generic addressing

    // AUTO = one cell per thread.  Measured on B200 (profiles/): with 40-48 registers the scalar kernel
    // keeps 40+ warps per SM in flight and every warp request is already a full 128-byte line, so it
    // sustains 6.5-6.9 TB/s; the 2- and 4-cell variants (64- and 128-bit accesses, x shifts by warp
    // shuffle) need 64-150 registers, drop to 16-24 warps per SM and are 3-8 % slower.  They stay
    // selectable.  Vector width is limited to 16-byte accesses inside one CSoA run.
    const int vmax = p->precision == LBM_F32 ? 4 : 2;
    int vec = (p->variant == LBM_VARIANT_VEC2 || p->variant == LBM_VARIANT_VEC4) ? p->variant : 1;
    c->aa = p->variant == LBM_VARIANT_AA;
    if (p->variant == LBM_VARIANT_TMA) {
        // eligibility; otherwise the scalar kernel is used
        c->tma = p->stride <= p->dim && p->stride * (long long)(p->precision == LBM_F32 ? 4 : 8) >= 16 && p->dim >= 32;
    }
    if (vec > vmax) vec = vmax;
    while (vec > 1 && (p->stride % vec != 0 || p->dim % vec != 0)) vec /= 2;
    c->vec = vec;
    choose_block(c);
    // block order (lbm_kernels.cuh, block_yz).  Measured on B200 (profiles/r02_experiments.md section 6): the tile
    // order helps exactly one kernel -- the SHIFT step of the in-place variant on lattices whose planes span several
    // waves of blocks (1024^3: 28.08 -> 26.67 ms with tiles of 32 rows x 16 planes) -- and costs the kernels that
    // stream whole rows (pull: -2.7 % at 512^3, LOCAL step: +4 %).  So: SHIFT launches at DIM >= 512 only.
    // LBM_BLOCK_SWIZZLE="ly,lz" (log2 extents) forces a tile order on every step launch, "0" disables it.
    if (c->aa && c->dim >= 512) {
        c->swz_y = 5;
        c->swz_z = 4;
        c->swz_shift_only = true;
    }
    if (const char *e = std::getenv("LBM_BLOCK_SWIZZLE")) {
        int ly = -1, lz = -1;
        if (std::sscanf(e, "%d,%d", &ly, &lz) == 2 && ly >= 0 && lz >= 0 && ly + lz <= 20) {
            c->swz_y = ly;
            c->swz_z = lz;
            c->swz_shift_only = false;
        } else {
            c->swz_y = c->swz_z = -1;
        }
    }

    if (const char *t = std::getenv("LBM_SYNC_TIMEOUT_S")) {
        const double s = std::atof(t);
        if (s > 0) c->sync_timeout_ns = (unsigned long long)(s * 1e9);
    }

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

    c->f_bytes = (size_t)c->n_alloc * Q * c->esize;
    c->flag_off = (c->f_bytes + 255) / 256 * 256;  // lattice 0 carries the slab flag words behind it (one IPC handle)
    const size_t m_bytes = (size_t)c->n_local * c->esize;
    for (int i = 0; i < (c->aa ? 1 : 2); ++i) {
        const size_t bytes = i == 0 ? c->flag_off + 256 : c->f_bytes;
        cudaError_t err = cudaMalloc(&c->f[i], bytes);
        if (err == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            fail(c, LBM_ERR_OOM,
                 "lbm_create: cudaMalloc of lattice %d (%.1f GB) failed - cudaErrorMemoryAllocation; %s", i,
                 (double)c->f_bytes / 1e9,
                 c->aa ? "split the cube over more GPUs"
                       : "the in-place variant (-A, LBM_VARIANT_AA) needs one lattice instead of two, or split the "
                         "cube over more GPUs (-G N)");
            return bail(LBM_ERR_OOM);
        }
        if (cuda_or_bail(err, "cudaMalloc(f)")) return bail(LBM_ERR_CUDA);
        c->device_bytes += (int64_t)c->f_bytes;
    }
    // the flag words start at zero and are never reset: phases only grow (see run_slab_flags)
    if (cuda_or_bail(cudaMemset(static_cast<char *>(c->f[0]) + c->flag_off, 0, 256), "cudaMemset(flags)")) return bail(LBM_ERR_CUDA);
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
    if (has_face[0] || has_face[1]) {
        if (cuda_or_bail(cudaMalloc(&c->sync_local, 4 * sizeof(unsigned)), "cudaMalloc(sync)")) return bail(LBM_ERR_OOM);
        if (cuda_or_bail(cudaMemset(c->sync_local, 0, 4 * sizeof(unsigned)), "cudaMemset(sync)")) return bail(LBM_ERR_CUDA);
    }

    if (p->precision == LBM_F32) rc = compute_stale<float>(c, c->cf, c->stale_f);
    else rc = compute_stale<double>(c, c->cd, c->stale_d);
    if (rc != LBM_OK) return bail(rc);

    if (c->tma) {
        rc = setup_tma(c, prop.multiProcessorCount);
        if (rc != LBM_OK) return bail(rc);
    }
    if (p->variant == LBM_VARIANT_NVRTC) {
        rc = lbm_nvrtc_build(c);
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
    if (c->copy_pending) LBM_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));  // `initialize` rewrites rho / u
    if (c->sync_mode == LBM_SYNC_FLAGS) {
        // `initialize` is a phase of its own: it rewrites my halo planes, so the neighbours' last stores
        // into them (their previous phase) must have landed; afterwards they may store again
        const bool has_lo = c->z_begin > 0, has_hi = c->z_end < c->dim;
        c->phase += 1;
        const SlabSync y = make_sync(c, c->phase - 1, c->phase);
        LBM_CUDA(c, launch_slab_wait(y, has_lo, has_hi, c->stream));
        LBM_CUDA(c, launch_init(c, c->stream));
        LBM_CUDA(c, launch_slab_signal(y, has_lo, has_hi, c->stream));
    } else {
        LBM_CUDA(c, launch_init(c, c->stream));
    }
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
// hold an unbounded number of CUDA events.  Entries are removed as they are folded; a pair whose stop
// event was never recorded (an enqueue failed half way) is dropped without being measured.
static int fold_events(lbm_ctx *c, bool all)
{
    if (!all && c->compute_events.size() < 4096) return LBM_OK;
    int rc = LBM_OK;
    while (!c->compute_events.empty()) {
        EventPair e = c->compute_events.front();
        c->compute_events.erase(c->compute_events.begin());
        if (e.stop_recorded && rc == LBM_OK) {
            float ms = 0.f;
            cudaError_t err = cudaEventSynchronize(e.stop);
            if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e.start, e.stop);
            if (err == cudaSuccess) {
                c->kernels_ms_accum += ms;
                if (c->launch_ms.size() < (1u << 20)) c->launch_ms.push_back(ms);
            } else {
                rc = fail(c, LBM_ERR_CUDA, "fold_events: %s(%d)", cudaGetErrorName(err), (int)err);
            }
        }
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    return rc;
}

// Opens a timed batch on the main stream; close_batch() records its end.  A batch that is never closed
// (an error in between) is discarded by fold_events.
static int open_batch(lbm_ctx *c)
{
    int rc = fold_events(c, false);
    if (rc != LBM_OK) return rc;
    EventPair ep;
    if ((rc = push_pair(c, &ep)) != LBM_OK) return rc;
    const cudaError_t e = cudaEventRecord(ep.start, c->stream);
    if (e != cudaSuccess) {
        cudaEventDestroy(ep.start);
        cudaEventDestroy(ep.stop);
        return fail(c, LBM_ERR_CUDA, "cudaEventRecord(%d) - %s", (int)e, cudaGetErrorName(e));
    }
    c->compute_events.push_back(ep);
    return LBM_OK;
}
static int close_batch(lbm_ctx *c)
{
    LBM_CUDA(c, cudaEventRecord(c->compute_events.back().stop, c->stream));
    c->compute_events.back().stop_recorded = true;
    return record_last(c);
}

// n iterations of whatever schedule the context is in; iteration numbers continue from the counter
static int run_iterations(lbm_ctx *c, int n_iterations, int every, bool allow_graphs)
{
    if (c->sync_mode == LBM_SYNC_FLAGS) return run_slab_flags(c, n_iterations, every);
    if (c->sync_mode == LBM_SYNC_NCCL) return run_slab_with_comm(c, n_iterations, every);
    int left = n_iterations;
    while (left > 0) {
        const int64_t it = c->iteration + 1;  // 1-based like lbmcl.hpp:435
        // launch-bound lattices: replay a captured chunk of unflagged iterations as one graph
        // (capturing + instantiating a chunk costs a few hundred microseconds of host time: only
        // worth it when at least LBM_GRAPH_MIN_CHUNKS replays follow, or when the graph exists already)
        const int gpar = c->aa ? (int)(c->iteration & 1) : c->cur;
        const bool have_graph = c->graph_exec[gpar] != nullptr && c->graph_stream == c->stream;
        if (allow_graphs && c->dim <= LBM_GRAPH_MAX_DIM && left >= LBM_GRAPH_CHUNK &&
            (have_graph || left >= LBM_GRAPH_MIN_CHUNKS * LBM_GRAPH_CHUNK)) {
            const int64_t last = it + LBM_GRAPH_CHUNK - 1;
            const bool flagged = every != 0 && (last / every) != ((it - 1) / every);
            if (!flagged) {
                const int rc = launch_graph_chunk(c);
                if (rc != LBM_OK) return rc;
                left -= LBM_GRAPH_CHUNK;
                continue;
            }
        }
        const bool macro = every != 0 && (it % every) == 0;
        LBM_CUDA(c, launch_step(c, Planes::range(c->z_begin, c->z_end), macro, PEER_NONE, c->stream));
        c->cur ^= 1;
        c->iteration = it;
        --left;
    }
    return LBM_OK;
}

// a slab whose neighbours are attached but that has no schedule of its own is driven by lbm_group_run
static int check_drivable(lbm_ctx *c, const char *who)
{
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "%s before lbm_init", who);
    if (c->sync_mode == LBM_SYNC_NONE && has_neighbours(c))
        return fail(c, LBM_ERR_STATE, "%s: this slab has peer neighbours; drive it through lbm_group_run (same "
                                      "process), or enable a transport (lbm_comm_fused / lbm_comm_init)", who);
    return LBM_OK;
}

int lbm_step(lbm_ctx *c, int update_macro)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = check_drivable(c, "lbm_step");
    if (rc != LBM_OK) return rc;
    if ((rc = use_device(c)) != LBM_OK) return rc;
    if ((rc = open_batch(c)) != LBM_OK) return rc;
    // one reference `compute` launch: never a graph replay
    if ((rc = run_iterations(c, 1, update_macro ? 1 : 0, false)) != LBM_OK) return rc;
    return close_batch(c);
}

int lbm_run(lbm_ctx *c, int n_iterations, int every)
{
    if (!c) return LBM_ERR_INVALID;
    if (n_iterations < 0 || every < 0) return fail(c, LBM_ERR_INVALID, "lbm_run: negative argument");
    int rc = check_drivable(c, "lbm_run");
    if (rc != LBM_OK) return rc;
    if (n_iterations == 0) return LBM_OK;
    if ((rc = use_device(c)) != LBM_OK) return rc;
    if ((rc = open_batch(c)) != LBM_OK) return rc;
    if ((rc = run_iterations(c, n_iterations, every, true)) != LBM_OK) return rc;
    return close_batch(c);
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
    if (c->sync_mode == LBM_SYNC_FLAGS && c->sync_local) {
        unsigned flag = 0;
        LBM_CUDA(c, cudaMemcpy(&flag, c->sync_local + 2, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag)
            return fail(c, LBM_ERR_CUDA, "z-slab transport: a wait for a neighbour's phase flag timed out after %.0f s "
                                         "(a rank died or the ranks ran different schedules); results are invalid",
                        (double)c->sync_timeout_ns * 1e-9);
    }
    return LBM_OK;
}

// device -> host copies of rho / u on stream `s`; slab == false: global layouts, owned planes only
static int enqueue_macro_copies(lbm_ctx *c, void *rho_host, void *u_host, bool slab, cudaStream_t s)
{
    const long long plane = (long long)c->dim * c->dim;
    const long long n_cube = plane * c->dim;
    const size_t off_local = (size_t)(c->z_begin - c->zs0) * plane * c->esize;
    const size_t off_host = slab ? 0 : (size_t)c->z_begin * plane * c->esize;
    const size_t bytes = (size_t)(c->z_end - c->z_begin) * plane * c->esize;
    const size_t u_pitch = slab ? bytes : (size_t)n_cube * c->esize;
    if (rho_host)
        LBM_CUDA(c, cudaMemcpyAsync((char *)rho_host + off_host, (const char *)c->rho + off_local, bytes,
                                    cudaMemcpyDeviceToHost, s));
    if (u_host)
        for (int k = 0; k < 3; ++k)
            LBM_CUDA(c, cudaMemcpyAsync((char *)u_host + (size_t)k * u_pitch + off_host,
                                        (const char *)c->u + (size_t)k * c->n_local * c->esize + off_local, bytes,
                                        cudaMemcpyDeviceToHost, s));
    return LBM_OK;
}

static int read_macros_blocking(lbm_ctx *c, void *rho_host, void *u_host, bool slab, const char *who)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "%s before lbm_init", who);
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    if ((rc = enqueue_macro_copies(c, rho_host, u_host, slab, c->stream)) != LBM_OK) return rc;
    if ((rc = record_last(c)) != LBM_OK) return rc;
    LBM_CUDA(c, cudaStreamSynchronize(c->stream));
    return LBM_OK;
}

int lbm_read_macros(lbm_ctx *c, void *rho_host, void *u_host)
{
    return read_macros_blocking(c, rho_host, u_host, false, "lbm_read_macros");
}

int lbm_read_macros_slab(lbm_ctx *c, void *rho_slab, void *u_slab)
{
    return read_macros_blocking(c, rho_slab, u_slab, true, "lbm_read_macros_slab");
}

// ---- asynchronous read-back ----

int lbm_host_alloc(size_t bytes, void **out)
{
    if (!out) return LBM_ERR_INVALID;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, LBM_ERR_OOM, "lbm_host_alloc: cudaHostAlloc(%zu) - %s", bytes, cudaGetErrorName(e));
    }
    return LBM_OK;
}

void lbm_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int lbm_read_macros_async(lbm_ctx *c, void *rho_host, void *u_host)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_read_macros_async before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    if (!c->copy_stream) {
        LBM_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        LBM_CUDA(c, cudaEventCreateWithFlags(&c->ev_copy_ready, cudaEventDisableTiming));
        LBM_CUDA(c, cudaEventCreate(&c->ev_copy_done));
    }
    // the copy sees everything enqueued so far and nothing later; later kernels that overwrite rho / u
    // wait for ev_copy_done (launch_step_p), everything else runs alongside the copy
    LBM_CUDA(c, cudaEventRecord(c->ev_copy_ready, c->stream));
    LBM_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy_ready, 0));
    if ((rc = enqueue_macro_copies(c, rho_host, u_host, false, c->copy_stream)) != LBM_OK) return rc;
    LBM_CUDA(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
    c->copy_pending = true;
    return LBM_OK;
}

int lbm_read_wait(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->copy_pending) return LBM_OK;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, cudaEventSynchronize(c->ev_copy_done));
    c->copy_pending = false;
    return record_last(c);  // "Total time" covers the read-back like the reference's blocking reads do (lbmcl.hpp:548-556)
}

int lbm_read_map(lbm_ctx *c, int32_t *map_host)
{
    if (!c || !map_host) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long n_cube = (long long)c->dim * c->dim * c->dim;
    int *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, (size_t)n_cube * sizeof(int)));
    cudaError_t e = launch_map(d, c->dim, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(map_host, d, (size_t)n_cube * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    LBM_CUDA(c, e);
    return record_last(c);
}

// The reference's view of the lattice the next iteration reads, over the context's OWNED planes, written
// into the device buffer `d` of 19 * DIM^3 elements in the reference's global CSoA order.
static cudaError_t enqueue_reference_view(lbm_ctx *c, void *d)
{
    ViewCfg v{};
    v.dim = c->dim;
    v.zs0 = c->zs0;
    v.nz_local = c->nz_local;
    v.z_begin = c->z_begin;
    v.z_end = c->z_end;
    v.lay_local = c->lay;
    v.lay_global = c->lay;
    v.pristine = c->iteration == 0 ? 1 : 0;
    v.aa_swapped = (c->iteration % 2) == 1 ? 1 : 0;
    const void *src = c->aa ? c->f[0] : c->f[c->cur];
    if (c->p.precision == LBM_F32)
        return launch_view_f32((const float *)src, (float *)d, v, c->cf, c->aa, c->stream);
    return launch_view_f64((const double *)src, (double *)d, v, c->cd, c->aa, c->stream);
}

int lbm_read_f(lbm_ctx *c, void *f_host)
{
    if (!c || !f_host) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_read_f before lbm_init");
    if (c->z_begin != 0 || c->z_end != c->dim)
        return fail(c, LBM_ERR_INVALID, "lbm_read_f: only on a context that owns the whole cube (slabs: lbm_group_read_f)");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    const long long n_cube = (long long)c->dim * c->dim * c->dim;
    const size_t bytes = (size_t)n_cube * Q * c->esize;
    void *d = nullptr;
    LBM_CUDA(c, cudaMalloc(&d, bytes));
    cudaError_t e = enqueue_reference_view(c, d);
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

int lbm_mark_end(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    if (!c->initialised) return fail(c, LBM_ERR_STATE, "lbm_mark_end before lbm_init");
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    return record_last(c);
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
    if (c->sync_mode == LBM_SYNC_FLAGS)
        return fail(c, LBM_ERR_STATE, "lbm_step_planes: the flag transport drives whole iterations (lbm_run / lbm_step)");
    if (z_begin < c->z_begin || z_end > c->z_end || z_begin > z_end)
        return fail(c, LBM_ERR_INVALID, "lbm_step_planes: [%d, %d) outside the owned planes [%d, %d)", z_begin,
                    z_end, c->z_begin, c->z_end);
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    LBM_CUDA(c, launch_step(c, Planes::range(z_begin, z_end), update_macro != 0, PEER_NONE, c->stream));
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

static cudaError_t halo_launch(lbm_ctx *c, void *lattice, void *dense, long long plane_local, int dir_up, bool pack)
{
    c->launches += 1;
    if (c->p.precision == LBM_F32)
        return launch_halo_f32((float *)lattice, (float *)dense, c->dim, plane_local, c->lay, dir_up, pack, c->stream);
    return launch_halo_f64((double *)lattice, (double *)dense, c->dim, plane_local, c->lay, dir_up, pack, c->stream);
}

// Packs, from the lattice the NEXT iteration reads (i.e. the one just written), what the neighbours
// will gather: low face -> populations with e_z = -1 of plane z_begin; high face -> e_z = +1 of plane
// z_end - 1.
int lbm_halo_pack(lbm_ctx *c)
{
    if (!c) return LBM_ERR_INVALID;
    int rc = use_device(c);
    if (rc != LBM_OK) return rc;
    void *lat = c->f[c->cur];
    if (c->halo_send[0]) LBM_CUDA(c, halo_launch(c, lat, c->halo_send[0], c->z_begin - c->zs0, 0, true));
    if (c->halo_send[1]) LBM_CUDA(c, halo_launch(c, lat, c->halo_send[1], (c->z_end - 1) - c->zs0, 1, true));
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
    if (c->halo_recv[0]) LBM_CUDA(c, halo_launch(c, lat, c->halo_recv[0], (c->z_begin - 1) - c->zs0, 1, false));
    if (c->halo_recv[1]) LBM_CUDA(c, halo_launch(c, lat, c->halo_recv[1], c->z_end - c->zs0, 0, false));
    return LBM_OK;
}

}  // extern "C"

#include "lbm_group.inl"
