/*
 * lbm_b200.h — C ABI of the B200-native D3Q19 BGK collide-and-stream path.
 *
 * This is the drop-in boundary for the one hot path of blackwut/LBMCL.  The reference has no
 * plugin/FFI interface: its host class LBMCL<T> (reference lbmcl.hpp) talks to the device through
 * the OpenCL C++ bindings wrapped by libs/CLUtil.hpp.  Every entry point below names the reference
 * call(s) it stands in for (paths relative to the reference tree), so that lbmcl.hpp could be
 * re-pointed at this library call by call.
 *
 * Conventions
 *   - plain C types only; no device pointers cross the ABI except where a function says so
 *     explicitly (the multi-process halo plumbing at the end);
 *   - every call returns LBM_OK (0) or a negative lbm_status; the message of the last failure on a
 *     context is available from lbm_last_error() (the reference prints "file:line what(code) - NAME"
 *     and exits, CLUtil.hpp:83-117; here the ABI never exits or throws — the host decides);
 *   - a context is driven by one host thread at a time (the reference has one in-order queue,
 *     CLUtil.hpp:190-199);
 *   - there is no CPU fallback: if no CUDA device is usable, lbm_create fails.
 *
 * Lattice conventions (identical to the reference): cube of DIM^3 cells, DIM a power of two,
 * linear cell id = x + y*DIM + z*DIM^2 (kernels.cl:67), outer one-cell WALL shell, D3Q19 direction
 * numbering of kernels.cl:131-205, CSoA(stride) population layout of kernels.cl:64.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 2

typedef enum lbm_status {
    LBM_OK = 0,
    LBM_ERR_INVALID = -1,     /* bad argument / unsupported configuration          */
    LBM_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the name) */
    LBM_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is no CPU fallback   */
    LBM_ERR_OOM = -4,         /* device or host allocation failed                  */
    LBM_ERR_STATE = -5        /* call made in the wrong order                      */
} lbm_status;

typedef enum lbm_precision { LBM_F32 = 0, LBM_F64 = 1 } lbm_precision;

/* Kernel organisation.  All variants compute the same function. */
typedef enum lbm_variant {
    LBM_VARIANT_AUTO = 0,   /* the fastest measured on B200: currently the scalar variant        */
    LBM_VARIANT_SCALAR = 1, /* two-lattice pull, one cell per thread                             */
    LBM_VARIANT_VEC2 = 2,   /* two-lattice pull, 2 cells per thread, x shifts by warp shuffle    */
    LBM_VARIANT_VEC4 = 4,   /* two-lattice pull, 4 cells per thread (128-bit fp32 access)        */
    LBM_VARIANT_AA = 8,     /* in-place AA pattern: ONE lattice (half the memory), one cell per
                               thread; whole cube on one device only                             */
    LBM_VARIANT_TMA = 16,   /* two-lattice pull fed by the TMA unit, warp-specialised: persistent CTAs of
                               consumer warps + one producer warp that keeps a ring of row tiles in flight
                               (bulk-tensor loads, full/empty mbarrier pairs); needs stride <= DIM,
                               stride*sizeof(T) >= 16, DIM >= 32 (else the scalar variant is used)  */
    LBM_VARIANT_NVRTC = 32  /* the scalar kernel compiled at run time (NVRTC) with DIM, stride, the
                               address offsets, INV_TAU and U as literals -- what the reference does
                               with its -D kernel options (lbmcl.hpp:131-156, CLUtil.hpp:201-229);
                               single-device launches only, slab launches use the scalar variant    */
} lbm_variant;

/*
 * Configuration of one context = one device = one z-slab of the cube.
 * Stands in for the LBMCL<T> constructor arguments (lbmcl.hpp:338-388), the -D kernel build options
 * (lbmcl.hpp:131-156) and the device choice (CLUtil.hpp:119-178).
 */
typedef struct lbm_params {
    int32_t abi_version;   /* must be LBM_B200_ABI_VERSION                                         */
    int32_t dim;           /* cube edge, power of two, >= 4   (-d; lbmcl.hpp:364-367)              */
    int32_t precision;     /* lbm_precision                   (-F; main.cpp:44-48)                 */
    int32_t fast_math;     /* 0 strict IEEE op order, 1 contracted / approximate division
                              (-o, "-cl-fast-relaxed-math", lbmcl.hpp:151-153)                      */
    double viscosity;      /* as given on the command line; the library applies the reference's
                              6-significant-digit text round trip (lbmcl.hpp:140-141)   */
    double velocity;       /* lid speed, same round trip                                            */
    int64_t stride;        /* CSoA stride, power of two, 1 .. DIM^3 (-s; lbmcl.hpp:384-387)         */
    int32_t block_x;       /* requested work-group shape (-w; lbmcl.hpp:371-382); a hint: by        */
    int32_t block_y;       /*   default the library uses x-major rows (see lbm_block_shape)         */
    int32_t block_z;
    int32_t device;        /* CUDA ordinal (-D); <0 = current device                                */
    int32_t variant;       /* lbm_variant                                                           */
    int32_t z_begin;       /* owned global planes [z_begin, z_end); 0, DIM for a single device      */
    int32_t z_end;
    int32_t reserved[8];   /* zero, except the tuning/test hooks: [0] = 1 forces the generic CSoA
                              addressing, [1] = 1 honours block_x/y/z exactly                        */
} lbm_params;

typedef struct lbm_ctx lbm_ctx;

/* Fill *p with the reference's defaults (lbm_options.hpp:32-50): dim 8, nu 0.0089, U 0.05,
 * stride 32, work group 8,8,8, fp32, strict math, whole cube on the current device. */
void lbm_default_params(lbm_params *p);

/* Device selection + context + queue + program build + the five buffers
 * (lbmcl.hpp:392-417 setupSimulation; CLUtil.hpp:119-229).  Allocates two f lattices, rho, u. */
int lbm_create(const lbm_params *p, lbm_ctx **out);

/* Releases everything the context owns (the cl::Buffer / cl::Kernel destructors, lbmcl.hpp:672). */
void lbm_destroy(lbm_ctx *ctx);

/* Message of the last failure on ctx (or of the last failed lbm_create when ctx is NULL). */
const char *lbm_last_error(const lbm_ctx *ctx);

/* enqueue of the `initialize` kernel (lbmcl.hpp:493-498; kernels.cl:277-318).  Asynchronous.
 * Resets the iteration counter to 0. */
int lbm_init(lbm_ctx *ctx);

/* enqueue of ONE `compute` launch = one iteration (lbmcl.hpp:505-511; kernels.cl:321-425) with the
 * reference's update_macro flag (lbmcl.hpp:436, 461).  Asynchronous; an event pair is recorded so
 * that the launch is counted by lbm_time_ms exactly like a reference "compute" event. */
int lbm_step(lbm_ctx *ctx, int update_macro);

/* The loop of lbmcl.hpp:505-511 without the readbacks: n_iterations launches, iteration numbers
 * continuing from the context's counter, update_macro = (every != 0 && it % every == 0).
 * Asynchronous.  One event pair brackets the whole batch. */
int lbm_run(lbm_ctx *ctx, int n_iterations, int every);

/* queue.finish() (lbmcl.hpp:525-532). */
int lbm_sync(lbm_ctx *ctx);

/* Blocking readback of rho and u (lbmcl.hpp:268-277 inside storeData).  rho -> host[N], u ->
 * host[3][N] with N = DIM^3 in the reference's GLOBAL layout (kernels.cl:67-70); a slab context
 * writes only its owned planes.  Either pointer may be NULL.  Element type = the context's precision. */
int lbm_read_macros(lbm_ctx *ctx, void *rho_host, void *u_host);

/* The same for a slab context without a cube-sized host array: rho -> host[P], u -> host[3][P] with
 * P = (z_end - z_begin) * DIM^2, i.e. only the owned planes, compact. */
int lbm_read_macros_slab(lbm_ctx *ctx, void *rho_slab, void *u_slab);

/* Asynchronous read-back for hosts that overlap output with computation (the reference's "Total MLUPS"
 * includes read-back and VTI writing, lbmcl.hpp:261-334, 548-556, 604-607).
 * lbm_host_alloc / lbm_host_free: page-locked host memory (the copies below are only asynchronous into it).
 * lbm_read_macros_async: same layouts as lbm_read_macros; returns at once.  The copy runs on the context's
 * copy stream after everything enqueued so far; iterations enqueued afterwards run concurrently with it,
 * except that the next iteration that overwrites rho / u (a flagged one) is ordered after the copy by the
 * library.  One read may be outstanding per context.  lbm_read_wait blocks until it has completed. */
int lbm_host_alloc(size_t bytes, void **out);
void lbm_host_free(void *p);
int lbm_read_macros_async(lbm_ctx *ctx, void *rho_host, void *u_host);
int lbm_read_wait(lbm_ctx *ctx);

/* Blocking readback of the cell-type map (lbmcl.hpp:164 inside storeMap), global layout int32[N]. */
int lbm_read_map(lbm_ctx *ctx, int32_t *map_host);

/* Blocking readback of the population lattice that the NEXT iteration will read, presented as the
 * reference stores it: pre-collision / post-streaming values in CSoA(stride) order over the whole
 * cube (lbmcl.hpp:211 inside storeF, called as in lbmcl.hpp:503, 517-519).  host[19*N]. */
int lbm_read_f(lbm_ctx *ctx, void *f_host);

/* Profiling numbers (CLUtil.hpp:231-243 over the event list, lbmcl.hpp:548-574):
 * total_ms   = start of `initialize` to end of the last enqueued work,
 * kernels_ms = sum of the durations of the compute launches only (lbmcl.hpp:568-571).
 * Synchronises the context. */
int lbm_time_ms(lbm_ctx *ctx, double *total_ms, double *kernels_ms);

/* Extends the profiled span to "now" on the context's stream: a host that overlaps its output work with the
 * device (asynchronous read-back, file writing on other threads) calls it when that work is finished, so that
 * total_ms covers it the way the reference's total covers its blocking reads and the VTI writing between
 * them (first event start -> last event end, lbmcl.hpp:548-556). */
int lbm_mark_end(lbm_ctx *ctx);

/* The individual durations behind kernels_ms, in enqueue order: one entry per lbm_step() launch or per
 * lbm_run() batch -- the per-event list of kernelsTimingsMS() (lbmcl.hpp:580-593).  *count receives the
 * number of entries; at most `capacity` of them are copied to `out` (may be NULL).  Synchronises. */
int lbm_launch_times_ms(lbm_ctx *ctx, double *out, int64_t capacity, int64_t *count);

/* CL_DEVICE_NAME (lbmcl.hpp:626, 651). */
int lbm_device_name(const lbm_ctx *ctx, char *buf, size_t buflen);

/* Effective parameters: what the reference would have compiled into the kernel.
 * out[0] = viscosity literal, out[1] = velocity literal, out[2] = INV_TAU (kernels.cl:61-62). */
int lbm_effective_params(const lbm_ctx *ctx, double out[3]);

/* The CUDA block actually used for the compute kernel, the cells per thread, and the
 * bytes of device memory the context allocated. */
int lbm_block_shape(const lbm_ctx *ctx, int32_t block[3], int32_t *cells_per_thread);
int64_t lbm_device_bytes(const lbm_ctx *ctx);

/* LBM_VARIANT_NVRTC without a device: compile the specialised step kernel for `p` (dim, stride, precision,
 * fast_math, viscosity, velocity) and hand back the sm_100a cubin -- for SASS inspection (cuobjdump) and to
 * check on a GPU-less box that the embedded kernel source builds.  *size = cubin bytes; at most `capacity`
 * bytes are copied to `out` (may be NULL). */
int lbm_spec_cubin(const lbm_params *p, void *out, size_t capacity, size_t *size);

/* Number of compute-kernel launches enqueued since lbm_init (for benchmark bookkeeping). */
int64_t lbm_launch_count(const lbm_ctx *ctx);

/* Iterations performed since lbm_init. */
int64_t lbm_iteration(const lbm_ctx *ctx);

/* Run all subsequent work of this context on an existing CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the context's own stream).  Lets a host that already owns a stream — e.g. the
 * Python benchmark's torch stream — time the kernels with its own events. */
int lbm_set_stream(lbm_ctx *ctx, void *cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * z-slab decomposition (new functionality; the reference is single-device).
 *
 * A context created with 0 < z_begin or z_end < DIM owns planes [z_begin, z_end) and stores one
 * extra halo plane per interior face.  Per iteration only the 5 populations that cross a face
 * travel: upward (e_z = +1) q = 6,15,16,17,18 and downward (e_z = -1) q = 5,11,12,13,14.
 *
 * Transports (all give the single-device bits):
 *  (1)  same process (lbm_group_*, CLI -G N): the boundary-plane kernel of a slab stores the crossing
 *       populations straight into the neighbour's halo plane (NVLink peer stores), the interior runs
 *       concurrently, devices are ordered with events only;
 *  (2a) one process per device, host-driven: lbm_step_planes / lbm_advance split the iteration,
 *       lbm_halo_pack() fills the two dense send buffers from the lattice the next iteration reads,
 *       lbm_halo_unpack() scatters the two dense receive buffers into the halo planes; the four
 *       buffers are DEVICE pointers owned by the context (5*DIM^2 elements each), exposed so that the
 *       host can hand them to its own communication library;
 *  (2b) one process per device, library-driven dense exchange over NCCL (lbm_comm_*);
 *  (2c) one process per device, fused: peer stores as in (1) through CUDA IPC mappings (lbm_ipc_*) and
 *       in-kernel epoch flags -- ONE kernel launch per iteration and rank does compute + exchange + ordering;
 *       no NCCL, no second stream, no events in the loop (lbm_comm_fused(ctx, LBM_FUSED_FLAGS));
 *  (2d) as (2c) but ordered by a one-word NCCL token per face and iteration (LBM_FUSED_TOKEN; the
 *       round-1 form, kept as the fallback where in-kernel polling is not wanted).
 * ------------------------------------------------------------------------------------------------ */

typedef enum lbm_face { LBM_FACE_LOW = 0, LBM_FACE_HIGH = 1 } lbm_face;

/* Split-phase iteration for hosts that overlap the halo exchange themselves: launch the compute
 * kernel over the owned global planes [z_begin, z_end) only, on the context's current stream
 * (lbm_set_stream), without advancing the iteration; lbm_advance() then swaps the two lattices and
 * increments the iteration counter (host-side bookkeeping only).  One iteration = every owned plane
 * covered exactly once by lbm_step_planes calls, then one lbm_advance. */
int lbm_step_planes(lbm_ctx *ctx, int z_begin, int z_end, int update_macro);
int lbm_advance(lbm_ctx *ctx);
/* owned plane range of the context */
int lbm_z_range(const lbm_ctx *ctx, int32_t *z_begin, int32_t *z_end);

/* number of elements (not bytes) in one packed halo: 5 * DIM * DIM */
int64_t lbm_halo_elems(const lbm_ctx *ctx);
/* device pointers of the dense halo buffers; NULL for a face on the cube boundary */
void *lbm_halo_send_buffer(lbm_ctx *ctx, int face);
void *lbm_halo_recv_buffer(lbm_ctx *ctx, int face);
/* pack the outgoing populations of the owned boundary planes (asynchronous, on the ctx stream) */
int lbm_halo_pack(lbm_ctx *ctx);
/* scatter the received populations into the halo planes (asynchronous, on the ctx stream) */
int lbm_halo_unpack(lbm_ctx *ctx);

/* (2b) one process per device, exchange driven by the library: the context owns an NCCL communicator
 * over the ranks of the job (libnccl.so.2 is resolved at run time with dlopen -- the one torch already
 * loaded under torchrun).  Rank 0 obtains an id with lbm_comm_unique_id(), the host broadcasts the 128
 * bytes (torch.distributed, MPI, ...), every rank calls lbm_comm_init() -- a collective call.  After
 * that lbm_run() on the slab context runs the overlapped schedule itself, per iteration: boundary planes
 * on a high-priority stream -> pack -> ncclSend/ncclRecv of the 5 crossing populations per face ->
 * unpack, with the interior planes concurrently on the main stream; no host synchronisation and no
 * per-step host round trip through the caller. */
#define LBM_COMM_ID_BYTES 128
int lbm_comm_unique_id(uint8_t id[LBM_COMM_ID_BYTES]);
int lbm_comm_init(lbm_ctx *ctx, const uint8_t id[LBM_COMM_ID_BYTES], int rank, int world);

/* (2c)/(2d) fused exchange.  lbm_ipc_export() describes this context's two lattices and its flag words
 * (CUDA IPC memory handles + geometry, LBM_IPC_HANDLE_BYTES opaque bytes); the host ships the blob to the
 * neighbouring ranks, which call lbm_ipc_attach(ctx, face, blob_of_the_neighbour_on_that_face).  The call
 * checks that the two slabs are adjacent (the neighbour's owned planes end where mine begin).
 * lbm_peer_attach() is the same for a neighbour context of the SAME process (any device with peer access).
 * lbm_ipc_detach() closes every mapping again and switches the fused transport off.
 *
 * lbm_comm_fused(ctx, mode) selects how a context whose interior faces are all attached runs lbm_run():
 *   LBM_FUSED_OFF    not fused: dense halos over NCCL if a communicator exists (2b);
 *   LBM_FUSED_FLAGS  (2c) one launch per iteration: the grid starts with the slab's boundary planes, whose
 *                    blocks wait -- bounded, LBM_SYNC_TIMEOUT_S, default 20 s; lbm_sync reports a timeout --
 *                    for the neighbours' previous phase, store the 5 crossing populations per face straight
 *                    into the neighbours' halo planes over NVLink and publish the new phase in the
 *                    neighbours' flag words; the interior planes follow in the same grid.  lbm_init is a
 *                    phase too, so re-initialising needs no host barrier.  Every rank must run the same
 *                    sequence of lbm_init / iterations; the neighbours must be on other devices (or the
 *                    grids small enough to be co-resident).  No communicator needed.  lbm_init must follow.
 *   LBM_FUSED_TOKEN  (2d) boundary planes on a high-priority stream with peer stores, interior concurrently,
 *                    one NCCL send/recv word per face and iteration; needs lbm_comm_init.
 * EVERY rank must make the same choice: the host enables a mode only after all ranks have reported
 * successful attachment (lbmcl_b200/slabs.py::connect_slabs). */
#define LBM_IPC_HANDLE_BYTES 256
typedef enum lbm_fused_mode { LBM_FUSED_OFF = 0, LBM_FUSED_FLAGS = 1, LBM_FUSED_TOKEN = 2 } lbm_fused_mode;
int lbm_ipc_export(lbm_ctx *ctx, uint8_t blob[LBM_IPC_HANDLE_BYTES]);
int lbm_ipc_attach(lbm_ctx *ctx, int face, const uint8_t blob[LBM_IPC_HANDLE_BYTES]);
int lbm_peer_attach(lbm_ctx *ctx, int face, lbm_ctx *neighbour);
int lbm_ipc_detach(lbm_ctx *ctx);
int lbm_comm_fused(lbm_ctx *ctx, int mode);

/* Same-process group of slab contexts, ordered by z.  lbm_group_create builds n contexts from one
 * parameter block (z range split evenly over devices[0..n-1]), enables peer access and links
 * neighbours.  The group calls mirror the single-context ones. */
typedef struct lbm_group lbm_group;
int lbm_group_create(const lbm_params *p, const int32_t *devices, int n, lbm_group **out);
void lbm_group_destroy(lbm_group *g);
const char *lbm_group_last_error(const lbm_group *g);
int lbm_group_size(const lbm_group *g);
lbm_ctx *lbm_group_ctx(lbm_group *g, int i);
int lbm_group_init(lbm_group *g);
int lbm_group_run(lbm_group *g, int n_iterations, int every);
int lbm_group_sync(lbm_group *g);
int lbm_group_read_macros(lbm_group *g, void *rho_host, void *u_host);
/* lbm_read_f over the slabs of a group (storeF, lbmcl.hpp:206-258): host[19*N], the reference's layout */
int lbm_group_read_f(lbm_group *g, void *f_host);
int lbm_group_time_ms(lbm_group *g, double *total_ms, double *kernels_ms);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
