// Same-process z-slab group (included by lbm_capi.cu).  New functionality: the reference is
// single-device.  One host thread drives n devices; per iteration and per device
//
//   boundary stream (high priority):  the first and last owned plane -> step kernel with PEER stores:
//                                     the 5 crossing populations per face go straight into the
//                                     neighbour's halo plane (NVLink when the neighbour is a peer)
//   main stream:                      all interior planes, concurrently
//
// Ordering is by events only (no host synchronisation inside the loop):
//   boundary(k) of slab i   waits for  boundary(k-1) of slabs i-1, i+1  (their halo writes / reads)
//                           and for    interior(k-1) of slab i          (it read the planes rewritten now)
//   interior(k) of slab i   waits for  boundary(k-1) of slab i
// Two event sets alternate with the iteration parity so that a re-recorded event is never the one a
// neighbour still has to wait for.

struct lbm_group {
    std::vector<lbm_ctx *> ctx;
    std::vector<cudaStream_t> bstream;
    std::vector<cudaEvent_t> ev_b[2];
    std::vector<cudaEvent_t> ev_i[2];
    std::vector<cudaEvent_t> ev_join;
    std::string error;
};

namespace {

thread_local std::string g_group_create_error;

int gfail(lbm_group *g, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g) g->error = buf;
    else g_group_create_error = buf;
    return code;
}

#define LBM_GCUDA(g, ...)                                                                               \
    do {                                                                                                \
        cudaError_t e__ = (__VA_ARGS__);                                                                \
        if (e__ != cudaSuccess)                                                                         \
            return gfail((g), LBM_ERR_CUDA, "%s:%d %s(%d) - %s", __FILE__, __LINE__, #__VA_ARGS__, (int)e__, \
                         cudaGetErrorName(e__));                                                        \
    } while (0)

}  // namespace

extern "C" {

const char *lbm_group_last_error(const lbm_group *g) { return g ? g->error.c_str() : g_group_create_error.c_str(); }
int lbm_group_size(const lbm_group *g) { return g ? (int)g->ctx.size() : 0; }
lbm_ctx *lbm_group_ctx(lbm_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

void lbm_group_destroy(lbm_group *g)
{
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->bstream.size() && g->bstream[i]) {
            cudaStreamSynchronize(g->bstream[i]);
            cudaStreamDestroy(g->bstream[i]);
        }
        for (int p = 0; p < 2; ++p) {
            if (i < g->ev_b[p].size() && g->ev_b[p][i]) cudaEventDestroy(g->ev_b[p][i]);
            if (i < g->ev_i[p].size() && g->ev_i[p][i]) cudaEventDestroy(g->ev_i[p][i]);
        }
        if (i < g->ev_join.size() && g->ev_join[i]) cudaEventDestroy(g->ev_join[i]);
    }
    for (lbm_ctx *c : g->ctx) lbm_destroy(c);
    delete g;
}

int lbm_group_create(const lbm_params *p, const int32_t *devices, int n, lbm_group **out)
{
    if (!out) return gfail(nullptr, LBM_ERR_INVALID, "lbm_group_create: out is NULL");
    *out = nullptr;
    if (!p || !devices || n < 1) return gfail(nullptr, LBM_ERR_INVALID, "lbm_group_create: bad arguments");
    if (p->variant == LBM_VARIANT_AA && n > 1)
        return gfail(nullptr, LBM_ERR_INVALID, "lbm_group_create: the AA variant is single-device");
    if (p->dim < 4 || (p->dim % n) != 0 || p->dim / n < 1)
        return gfail(nullptr, LBM_ERR_INVALID, "lbm_group_create: dim %d is not divisible into %d z-slabs", p->dim, n);
    lbm_group *g = new (std::nothrow) lbm_group();
    if (!g) return gfail(nullptr, LBM_ERR_OOM, "lbm_group_create: host allocation failed");
    auto bail = [&](int code, const std::string &msg) {
        g_group_create_error = msg;
        lbm_group_destroy(g);
        return code;
    };
    const int nz = p->dim / n;
    for (int i = 0; i < n; ++i) {
        lbm_params q = *p;
        q.device = devices[i];
        q.z_begin = i * nz;
        q.z_end = (i + 1) * nz;
        lbm_ctx *c = nullptr;
        const int rc = lbm_create(&q, &c);
        if (rc != LBM_OK) return bail(rc, std::string("slab ") + std::to_string(i) + ": " + lbm_last_error(nullptr));
        g->ctx.push_back(c);
    }
    // neighbours: peer access between the devices, each other's lattices as store targets
    for (int i = 0; i < n; ++i) {
        for (int face = 0; face < 2; ++face) {
            lbm_ctx *nb = face == 0 ? (i > 0 ? g->ctx[i - 1] : nullptr) : (i + 1 < n ? g->ctx[i + 1] : nullptr);
            if (!nb) continue;
            const int rc = lbm_peer_attach(g->ctx[i], face, nb);  // peer access + the neighbour's lattices
            if (rc != LBM_OK) return bail(rc, std::string("slab ") + std::to_string(i) + ": " + lbm_last_error(g->ctx[i]));
        }
    }
    g->bstream.assign(n, nullptr);
    g->ev_join.assign(n, nullptr);
    for (int p2 = 0; p2 < 2; ++p2) {
        g->ev_b[p2].assign(n, nullptr);
        g->ev_i[p2].assign(n, nullptr);
    }
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(g->ctx[i]->device);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi is the numerically lowest = highest priority
        cudaError_t e = cudaStreamCreateWithPriority(&g->bstream[i], cudaStreamNonBlocking, hi);
        for (int p2 = 0; p2 < 2 && e == cudaSuccess; ++p2) {
            e = cudaEventCreateWithFlags(&g->ev_b[p2][i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_i[p2][i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_join[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return bail(LBM_ERR_CUDA, std::string("group stream/event setup - ") + cudaGetErrorName(e));
    }
    *out = g;
    return LBM_OK;
}

int lbm_group_init(lbm_group *g)
{
    if (!g) return LBM_ERR_INVALID;
    for (lbm_ctx *c : g->ctx) {
        const int rc = lbm_init(c);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
    }
    // initialisation is a one-off: a full synchronisation keeps the event protocol of the loop simple
    for (lbm_ctx *c : g->ctx) {
        const int rc = lbm_sync(c);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
    }
    return LBM_OK;
}

int lbm_group_run(lbm_group *g, int n_iterations, int every)
{
    if (!g) return LBM_ERR_INVALID;
    if (n_iterations < 0 || every < 0) return gfail(g, LBM_ERR_INVALID, "lbm_group_run: negative argument");
    if (n_iterations == 0) return LBM_OK;
    const int n = (int)g->ctx.size();
    for (int i = 0; i < n; ++i) {
        lbm_ctx *c = g->ctx[i];
        if (!c->initialised) return gfail(g, LBM_ERR_STATE, "lbm_group_run before lbm_group_init");
        LBM_GCUDA(g, cudaSetDevice(c->device));
        const int rc = open_batch(c);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
        // the boundary stream starts after whatever the main stream was asked to do before
        LBM_GCUDA(g, cudaEventRecord(g->ev_join[i], c->stream));
        LBM_GCUDA(g, cudaStreamWaitEvent(g->bstream[i], g->ev_join[i], 0));
    }
    for (int k = 0; k < n_iterations; ++k) {
        const int64_t it = g->ctx[0]->iteration + 1;
        const bool macro = every != 0 && (it % every) == 0;
        const int par = (int)(it & 1), prev = par ^ 1;
        for (int i = 0; i < n; ++i) {
            lbm_ctx *c = g->ctx[i];
            LBM_GCUDA(g, cudaSetDevice(c->device));
            const bool has_lo = c->peer[0] != nullptr, has_hi = c->peer[1] != nullptr;
            cudaStream_t bs = g->bstream[i];
            if (has_lo) LBM_GCUDA(g, cudaStreamWaitEvent(bs, g->ev_b[prev][i - 1], 0));
            if (has_hi) LBM_GCUDA(g, cudaStreamWaitEvent(bs, g->ev_b[prev][i + 1], 0));
            LBM_GCUDA(g, cudaStreamWaitEvent(bs, g->ev_i[prev][i], 0));
            LBM_GCUDA(g, launch_step(c, boundary_planes(c), macro, PEER_STORE, bs));  // both faces, one launch
            LBM_GCUDA(g, cudaEventRecord(g->ev_b[par][i], bs));

            LBM_GCUDA(g, cudaStreamWaitEvent(c->stream, g->ev_b[prev][i], 0));
            LBM_GCUDA(g, launch_step(c, interior_planes(c), macro, PEER_NONE, c->stream));
            LBM_GCUDA(g, cudaEventRecord(g->ev_i[par][i], c->stream));
        }
        for (lbm_ctx *c : g->ctx) {
            c->cur ^= 1;
            c->iteration = it;
        }
    }
    const int last = (int)(g->ctx[0]->iteration & 1);
    for (int i = 0; i < n; ++i) {
        lbm_ctx *c = g->ctx[i];
        LBM_GCUDA(g, cudaSetDevice(c->device));
        LBM_GCUDA(g, cudaStreamWaitEvent(c->stream, g->ev_b[last][i], 0));
        const int rc = close_batch(c);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
    }
    return LBM_OK;
}

int lbm_group_sync(lbm_group *g)
{
    if (!g) return LBM_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        LBM_GCUDA(g, cudaSetDevice(g->ctx[i]->device));
        LBM_GCUDA(g, cudaStreamSynchronize(g->bstream[i]));
        LBM_GCUDA(g, cudaStreamSynchronize(g->ctx[i]->stream));
    }
    return LBM_OK;
}

int lbm_group_read_macros(lbm_group *g, void *rho_host, void *u_host)
{
    if (!g) return LBM_ERR_INVALID;
    int rc = lbm_group_sync(g);
    if (rc != LBM_OK) return rc;
    for (lbm_ctx *c : g->ctx) {
        rc = lbm_read_macros(c, rho_host, u_host);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
    }
    return LBM_OK;
}

// The -f view over a group (reference storeF, lbmcl.hpp:206-258, has no single-device restriction because the
// reference IS single-device): every slab renders the reference's pre-collision view of its owned planes
// (its halo planes hold the neighbours' crossing populations), the host merges the owned cells.
int lbm_group_read_f(lbm_group *g, void *f_host)
{
    if (!g || !f_host) return LBM_ERR_INVALID;
    int rc = lbm_group_sync(g);
    if (rc != LBM_OK) return rc;
    lbm_ctx *c0 = g->ctx[0];
    const long long dim = c0->dim, plane = dim * dim, n_cube = plane * dim;
    const size_t es = c0->esize, bytes = (size_t)n_cube * Q * es;
    std::vector<unsigned char> tmp(bytes);
    unsigned char *out = static_cast<unsigned char *>(f_host);
    for (lbm_ctx *c : g->ctx) {
        if (!c->initialised) return gfail(g, LBM_ERR_STATE, "lbm_group_read_f before lbm_group_init");
        LBM_GCUDA(g, cudaSetDevice(c->device));
        void *d = nullptr;
        LBM_GCUDA(g, cudaMalloc(&d, bytes));
        cudaError_t e = enqueue_reference_view(c, d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp.data(), d, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(d);
        LBM_GCUDA(g, e);
        // owned cells of this slab, all 19 slots each, in the global CSoA order
        const long long S = c->lay.qpitch();
        for (long long id = (long long)c->z_begin * plane; id < (long long)c->z_end * plane; ++id) {
            const long long b = c->lay.base(id);
            for (int q = 0; q < Q; ++q) std::memcpy(out + (size_t)(b + q * S) * es, tmp.data() + (size_t)(b + q * S) * es, es);
        }
    }
    return LBM_OK;
}

// total = slowest slab's init-start -> last work; kernels = slowest slab's sum of batch durations
int lbm_group_time_ms(lbm_group *g, double *total_ms, double *kernels_ms)
{
    if (!g) return LBM_ERR_INVALID;
    int rc = lbm_group_sync(g);
    if (rc != LBM_OK) return rc;
    double tmax = 0.0, kmax = 0.0;
    for (lbm_ctx *c : g->ctx) {
        double t = 0.0, k = 0.0;
        rc = lbm_time_ms(c, &t, &k);
        if (rc != LBM_OK) return gfail(g, rc, "%s", lbm_last_error(c));
        if (t > tmax) tmax = t;
        if (k > kmax) kmax = k;
    }
    if (total_ms) *total_ms = tmax;
    if (kernels_ms) *kernels_ms = kmax;
    return LBM_OK;
}

}  // extern "C"
