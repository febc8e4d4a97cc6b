// sm_100a kernels of the D3Q19 BGK collide-and-stream path.
//
// What the reference computes per iteration (kernels.cl:321-425, a PUSH scheme): read the 19
// pre-collision populations of a cell, lid pre-fix, rho/u, optional macro store, lid equilibrium or
// in-node bounce-back swap, BGK, scatter to the 19 neighbours of the other lattice.
//
// What these kernels do instead (two-lattice PULL scheme, post-collision storage):
//
//   G_k(c, q)  = value that iteration k leaves in cell c for direction q and that iteration k+1 of
//                cell c + e_q will gather           (the reference's S_k(c + e_q, q))
//
//   iteration:  f_q = G_{k-1}(c - e_q, q)  ->  BC / collide exactly as the reference  ->  G_k(c, q)
//
// so that every load and every store of a thread is an aligned, x-contiguous vector (VEC cells per
// thread); the +-1 shifts in x are done with warp shuffles, and only the first/last lane of a row
// segment issues one extra scalar load.  Cell types come from the coordinates (0 B/cell, SURVEY
// F12) instead of the reference's `map` buffer.
//
// Ghost values.  In the reference a slot (c, q) whose source c - e_q is a WALL cell is never written
// and keeps its `initialize` value for ever.  Here
//   * wall ROWS (y or z on 0 / DIM-1) are never written by the step kernel, and `init_kernel` leaves
//     in them exactly those constants (G_0(c, q) = initial equilibrium of the destination c + e_q), so
//     that gathering from them needs no special case;
//   * wall cells INSIDE a row (x = 0, DIM-1) are overwritten by the vector stores with don't-care
//     values; the populations that cells x = 1 / x = DIM-2 would gather from them are replaced by the
//     same constants from StepArgs::stale.
//
// Algorithmic traffic: 19 loads + 19 stores per cell = 152 B (fp32) / 304 B (fp64); nothing else is
// read.  Macro stores (4 values per cell) happen only on iterations flagged by `every`.
#pragma once

#include "lbm_d3q19.cuh"

namespace lbm {

template <typename T, int N>
struct alignas(sizeof(T) * N) Pack {
    T v[N];
};

// In-kernel ordering between neighbouring z-slabs that run in different processes / on different devices
// (PEER == PEER_FLAGS): one 32-bit epoch word per face, written by the neighbour over NVLink when ALL blocks
// of its boundary plane have stored their crossing populations, polled by the blocks of my boundary plane
// before they gather from the halo plane.  No second stream, no events, no NCCL in the iteration loop.
struct SlabSync {
    unsigned *flag_in;         // [2] in MY memory: [0] written by the low neighbour, [1] by the high neighbour
    unsigned *flag_out[2];     // the neighbours' words I write: [0] low neighbour's flag_in[1], [1] high's flag_in[0]
    unsigned *count;           // [2] in my memory: boundary-plane blocks that have finished, per face
    int *error;                // set to 1 when a wait ran into the time limit (lbm_sync reports it)
    unsigned wait_epoch;       // the neighbours must have completed this phase before I touch the halo planes
    unsigned signal_epoch;     // the phase this launch completes
    unsigned long long timeout_ns;
};

template <typename T>
struct StepArgs {
    T *__restrict__ dst;        // lattice written by this iteration      (G_k)
    const T *__restrict__ src;  // lattice gathered from                  (G_{k-1})
    T *__restrict__ rho;        // [n_local]
    T *__restrict__ u;          // [3][n_local]
    // z-slab neighbours: where the crossing populations of the first / last owned plane are ALSO stored
    // (the neighbour's halo plane of the lattice it reads next), or nullptr
    T *__restrict__ peer_lo;    // receives q with e_z = -1 of plane z_own_begin
    T *__restrict__ peer_hi;    // receives q with e_z = +1 of plane z_own_end - 1
    long long peer_lo_plane;    // local plane index of that halo plane in the neighbour's storage
    long long peer_hi_plane;
    int z_own_begin, z_own_end; // owned global planes of this slab
    int dim;
    int zs0;                    // global z of local plane 0
    // planes computed by this launch, in grid order: first zmap_n (0..2) individually named planes -- the
    // slab's boundary planes, so that their crossing populations are on their way before the interior
    // starts -- then the contiguous range [z_begin, z_end)
    int zmap_n, zmap0, zmap1;
    int z_begin, z_end;
    // block order (block_yz): log2 of the tile's extent in block rows / block planes and of the number of tiles
    // along y; swz_z < 0: the grid's own order
    int swz_y, swz_z, swz_nty;
    long long n_local;          // cells in local storage (pitch of the u components)
    Layout lay;
    // LM_BLOCKROWS (DIM < stride): a CSoA block holds 2^row_shift whole x-rows
    int row_shift, row_mask;    // log2(stride / DIM), stride / DIM - 1
    long long blk18;            // 18 * stride * sizeof(T): what crossing into the next CSoA block adds to an address
    Consts<T> c;
    T stale[2][Q];              // [0]: w_q (rest equilibrium); [1]: f_eq_q(1, (U,0,0)); see header
    // Byte offsets precomputed by the host (uniform; they live in the constant bank):
    //   goff[q]  LM_ROWS / LM_SOA: from the address of (x, y, z, 0) to the address of (x, y - ey, z - ez, q);
    //            LM_BLOCKROWS: the same as LM_SOA, valid while the source row lies in the same CSoA block --
    //            every block boundary crossed on the way adds blk18
    //   soff[q]  from the address of (x, y, z, 0) to the address of (x, y, z, q)      ( = q * stride * sizeof(T) )
    long long goff[Q];
    long long soff[Q];
    // AA variant, SHIFT step: from the address of (x + ex, y, z, 0) to the address of
    // (x + ex, y + ey, z + ez, q)
    long long poff[Q];
    SlabSync sync;
};

// How neighbour addresses are formed (chosen by the host from stride and DIM; all strides are powers of two):
//   LM_ROWS       stride <= DIM: every x-row starts a CSoA block, so a shift in y or z is a constant
//                 address offset and only the x position inside the row needs the block arithmetic;
//   LM_SOA        stride >= number of stored cells: one block, every neighbour is a constant offset;
//   LM_BLOCKROWS  DIM < stride < stored cells: a CSoA block holds stride/DIM whole x-rows, so the address of a
//                 neighbour row is the SOA-style constant offset plus 18*stride elements per block boundary
//                 crossed: the thread forms one full base and 8 small block-crossing counts (one per
//                 (e_y, e_z)) instead of a CSoA index per access; the x shift is +-1 element;
//   LM_GENERIC    full CSoA index computation per access (any stride; kept as a cross-check, test hook).
enum : int { LM_GENERIC = 0, LM_ROWS = 1, LM_SOA = 2, LM_BLOCKROWS = 3 };

// How the crossing populations of a slab's boundary planes reach the neighbour:
//   PEER_NONE   not at all (single device, interior launches, dense-halo transports)
//   PEER_STORE  stored straight into the neighbour's halo plane; ordering by CUDA events (same process)
//   PEER_FLAGS  the same stores + the in-kernel epoch flags of SlabSync (one launch per iteration)
enum : int { PEER_NONE = 0, PEER_STORE = 1, PEER_FLAGS = 2 };

// Uniform quantities.  Ahead-of-time build: kernel arguments (constant bank).  Run-time specialised build
// (NVRTC, lbm_nvrtc.cu -- the counterpart of the reference's -D kernel options, lbmcl.hpp:131-156):
// literals, so that DIM, the stride, the 38 address offsets and the physical constants become instruction
// immediates.
#ifdef LBM_SPEC_DIM
#define LBM_U_DIM(a) (LBM_SPEC_DIM)
#define LBM_U_LAY(a) (Layout{LBM_SPEC_SDIV, (long long)(LBM_SPEC_STRIDE) - 1})
#define LBM_U_ROWSHIFT(a) (LBM_SPEC_ROWSHIFT)
#define LBM_U_ROWMASK(a) ((1 << (LBM_SPEC_ROWSHIFT)) - 1)
#define LBM_U_BLK18(T, a) (18ll * (LBM_SPEC_STRIDE) * (long long)sizeof(T))
#define LBM_U_ZS0(a) (0)
#define LBM_U_NLOCAL(a) ((long long)(LBM_SPEC_DIM) * (LBM_SPEC_DIM) * (LBM_SPEC_DIM))
#define LBM_U_CONSTS(T, a) (Consts<T>{T(LBM_SPEC_U_LID), T(LBM_SPEC_INV_TAU), {T(1.0) / T(3.0), T(1.0) / T(18.0), T(1.0) / T(36.0)}})
#define LBM_U_SOFF(T, a, q) ((long long)(q) * (LBM_SPEC_STRIDE) * (long long)sizeof(T))
#define LBM_U_GOFF(T, a, q, LM)                                                                              \
    (((long long)(q) * (LBM_SPEC_STRIDE) -                                                                    \
      ((LM) == LM_ROWS ? (long long)Q : 1ll) *                                                               \
          ((long long)ey(q) * (LBM_SPEC_DIM) + (long long)ez(q) * (LBM_SPEC_DIM) * (LBM_SPEC_DIM))) *        \
     (long long)sizeof(T))
#else
#define LBM_U_DIM(a) ((a).dim)
#define LBM_U_LAY(a) ((a).lay)
#define LBM_U_ROWSHIFT(a) ((a).row_shift)
#define LBM_U_ROWMASK(a) ((a).row_mask)
#define LBM_U_BLK18(T, a) ((a).blk18)
#define LBM_U_ZS0(a) ((a).zs0)
#define LBM_U_NLOCAL(a) ((a).n_local)
#define LBM_U_CONSTS(T, a) ((a).c)
#define LBM_U_SOFF(T, a, q) ((a).soff[q])
#define LBM_U_GOFF(T, a, q, LM) ((a).goff[q])
#endif

template <typename T>
struct InitArgs {
    T *__restrict__ f0;
    T *__restrict__ f1;
    T *__restrict__ rho;
    T *__restrict__ u;
    int dim;
    int zs0;
    int nz_local;               // stored planes
    long long n_local;
    Layout lay;
    Consts<T> c;
};

// Initial equilibrium f_eq_q(1, (ux, 0, 0)) with the reference's operation order (kernels.cl:303-309);
// always strict: initial values are not subject to -o in any observable way except through rounding,
// and a single definition keeps the ghost constants identical everywhere.
template <typename T, int q>
__device__ __forceinline__ T init_feq(const Consts<T> &c, T ux)
{
    using A = Arith<T, false>;
    const T u2 = A::add(A::add(A::mul(ux, ux), T(0)), T(0));
    const T eu = e_dot_u<A, q>(ux, T(0), T(0));
    return A::mul(A::mul(T(1), c.w[wclass(q)]), eq_poly<A>(eu, A::mul(T(1.5), u2)));
}

// `initialize` (kernels.cl:277-318) for the pull representation: one thread per stored cell.
//   rho/u : 1 and (U or 0, 0, 0) on FLUID / MOVING cells, NaN elsewhere (kernels.cl:297-300)
//   both lattices: G_0(c, q) = f_eq_q(1, u0(c + e_q)), i.e. the initial state already "pre-streamed",
//   so that the first iteration gathers exactly the reference's analytic S_0.
template <typename T>
__global__ void __launch_bounds__(256) init_kernel(const InitArgs<T> a)
{
    const int dim = a.dim;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= dim || y >= dim || zl >= a.nz_local) return;
    const int z = a.zs0 + zl;
    const long long id = x + (long long)y * dim + (long long)zl * dim * dim;

    const int t = cell_type(x, y, z, dim);
    const bool keep = is_collision(t);
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    const T ux0 = has_front_bit(x, y, z, dim) ? a.c.u_lid : T(0);
    a.rho[id] = keep ? T(1) : nan;
    a.u[id] = keep ? ux0 : nan;
    a.u[a.n_local + id] = keep ? T(0) : nan;
    a.u[2 * a.n_local + id] = keep ? T(0) : nan;

    const long long b = a.lay.base(id);
    const long long qp = a.lay.qpitch();
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const bool front = has_front_bit(x + ex(q), y + ey(q), z + ez(q), dim);
        const T v = init_feq<T, q>(a.c, front ? a.c.u_lid : T(0));
        a.f0[b + q * qp] = v;
        a.f1[b + q * qp] = v;
    });
}

// The two ghost-constant tables of StepArgs::stale, computed on the device with the same code as
// init_kernel so that they are bit-identical to what the lattices hold.
template <typename T>
__global__ void stale_kernel(Consts<T> c, T *out /* [2][Q] */)
{
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        out[q] = init_feq<T, q>(c, T(0));
        out[Q + q] = init_feq<T, q>(c, c.u_lid);
    });
}

// BGK relaxation of one cell, kernels.cl:355-380 and :412-418 (fluid cells).
template <typename T, bool FAST>
__device__ __forceinline__ void collide_fluid(T (&f)[Q], const Consts<T> &c, T &rho, T &ux, T &uy, T &uz)
{
    using A = Arith<T, FAST>;
    // kernels.cl:355 — left-to-right sum
    rho = f[0];
#pragma unroll
    for (int q = 1; q < Q; ++q) rho = A::add(rho, f[q]);
    // kernels.cl:376-378
    const T px = A::add(A::add(A::add(A::add(f[1], f[7]), f[10]), f[11]), f[15]);
    const T mx = A::add(A::add(A::add(A::add(f[3], f[8]), f[9]), f[13]), f[17]);
    const T py = A::add(A::add(A::add(A::add(f[2], f[7]), f[8]), f[12]), f[16]);
    const T my = A::add(A::add(A::add(A::add(f[4], f[9]), f[10]), f[14]), f[18]);
    const T pz = A::add(A::add(A::add(A::add(f[6], f[15]), f[16]), f[17]), f[18]);
    const T mz = A::add(A::add(A::add(A::add(f[5], f[11]), f[12]), f[13]), f[14]);
    ux = A::div(A::sub(px, mx), rho);
    uy = A::div(A::sub(py, my), rho);
    uz = A::div(A::sub(pz, mz), rho);
    // kernels.cl:390
    const T u2 = A::add(A::add(A::mul(ux, ux), A::mul(uy, uy)), A::mul(uz, uz));
    const T c15u2 = A::mul(T(1.5), u2);
    const T rw[3] = { A::mul(rho, c.w[0]), A::mul(rho, c.w[1]), A::mul(rho, c.w[2]) };
    // kernels.cl:412-418 with compute_bgk (kernels.cl:270-273): f + INV_TAU * (feq - f).
    // Opposite directions are evaluated together: e_o.u == -(e_q.u) exactly, hence 3*eu changes sign
    // exactly, (4.5*eu)*eu is identical, and 1 + (-(3eu)) == 1 - 3eu bit for bit -- the reference's
    // individually rounded operations, nine of them shared instead of repeated.
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        constexpr int o = opp(q);
        if constexpr (q == 0) {
            const T feq = A::mul(rw[0], eq_poly<A>(T(0), c15u2));
            f[0] = A::add(f[0], A::mul(c.inv_tau, A::sub(feq, f[0])));
        } else if constexpr (q < o) {
            const T eu = e_dot_u<A, q>(ux, uy, uz);
            const T a3 = A::mul(T(3), eu);
            const T sq = A::mul(A::mul(T(4.5), eu), eu);
            const T pq = A::sub(A::add(A::add(T(1), a3), sq), c15u2);
            const T po = A::sub(A::add(A::sub(T(1), a3), sq), c15u2);
            const T fq = A::mul(rw[wclass(q)], pq);
            const T fo = A::mul(rw[wclass(o)], po);
            f[q] = A::add(f[q], A::mul(c.inv_tau, A::sub(fq, f[q])));
            f[o] = A::add(f[o], A::mul(c.inv_tau, A::sub(fo, f[o])));
        }
    });
}

// Moving-lid cell, kernels.cl:343-349, :362-366, :393-398, :412-418.  rho comes from the gathered
// populations with the five wall-side ones replaced by their opposites; u is forced to (U,0,0); all
// 19 populations become f_eq(rho, u).  The BGK step that follows in the reference acts on an exact
// equilibrium: feq - f == +0, INV_TAU * 0 == 0, f + 0 == f, so it is the identity and is not issued.
template <typename T, bool FAST>
__device__ __forceinline__ void collide_lid(T (&f)[Q], const Consts<T> &c, T &rho)
{
    using A = Arith<T, FAST>;
    f[5] = f[opp(5)];
    f[11] = f[opp(11)];
    f[12] = f[opp(12)];
    f[13] = f[opp(13)];
    f[14] = f[opp(14)];
    rho = f[0];
#pragma unroll
    for (int q = 1; q < Q; ++q) rho = A::add(rho, f[q]);
    const T U = c.u_lid;
    const T c15u2 = A::mul(T(1.5), A::mul(U, U));  // (U*U + 0*0) + 0*0 == U*U exactly
    const T rw[3] = { A::mul(rho, c.w[0]), A::mul(rho, c.w[1]), A::mul(rho, c.w[2]) };
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const T eu = e_dot_u<A, q>(U, T(0), T(0));
        f[q] = A::mul(rw[wclass(q)], eq_poly<A>(eu, c15u2));
    });
}

// In-node full-way bounce-back, kernels.cl:400-408: swap the nine opposite pairs, no collision.
template <typename T>
__device__ __forceinline__ void bounce_back(T (&f)[Q])
{
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        if constexpr (q < opp(q)) {
            const T tmp = f[q];
            f[q] = f[opp(q)];
            f[opp(q)] = tmp;
        }
    });
}

// Gather load.  -DLBM_LD_NOALLOC (an A/B knob, profiles/r02_experiments.md): the e_x = 0 gathers, whose lines no
// other warp touches, bypass L1 allocation (ld.global.L1::no_allocate).
template <typename T, int EX>
__device__ __forceinline__ T ld_gather(const T *p)
{
#ifdef LBM_LD_NOALLOC
    if constexpr (EX == 0 && sizeof(T) == 4) {
        float v;
        asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
        return v;
    } else if constexpr (EX == 0 && sizeof(T) == 8) {
        double v;
        asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    } else {
        return *p;
    }
#else
    return *p;
#endif
}

// ---- in-kernel slab ordering (PEER_FLAGS) ----
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait until *flag has reached `epoch` (wrap-around safe).  Bounded: after timeout_ns the error word is set
// and the caller carries on (the results are then wrong, lbm_sync reports it -- but the GPU never hangs).
__device__ __forceinline__ void slab_wait(const unsigned *flag, unsigned epoch, unsigned long long timeout_ns, int *error)
{
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned spins = 1;; ++spins) {
        if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
        if ((spins & 0x3ffu) == 0 && global_timer_ns() - t0 > timeout_ns) {
            if (error) atomicExch(error, 1);
            return;
        }
    }
}
// Called by one thread per block after the block's stores (and a __syncthreads): the last of `n_blocks`
// blocks publishes `epoch` in the neighbour's flag word.  fence / atomic / fence is the threadFenceReduction
// pattern: every block's peer stores are ordered before the flag store of the last one.
__device__ __forceinline__ void slab_signal(unsigned *count, unsigned n_blocks, unsigned *peer_flag, unsigned epoch)
{
    __threadfence_system();
    const unsigned prev = atomicAdd(count, 1u);
    if (prev == n_blocks - 1) {
        atomicExch(count, 0u);  // ready for the next launch (stream order separates the launches)
        __threadfence_system();
        st_release_sys(peer_flag, epoch);
    }
}

// Block order.  CUDA hands out the blocks of a grid x-fastest, then y, then z: plane by plane.  On a large lattice
// one plane is several waves of blocks, so the rows of the planes z -+ 1 that a block gathers from (10 of the 19
// populations) are touched again only thousands of blocks later -- each DRAM page of those planes is opened for
// a quarter of its bytes now and for the rest much later.  With a tile order the (y, z) pairs are walked in tiles
// of 2^swz_y block rows x 2^swz_z block planes (y fastest inside a tile, tiles along y first), so that most
// y -+ 1 AND z -+ 1 neighbours of a block are in flight in the same wave.  A bijection on the grid's (y, z) block
// indices as long as gridDim.y is a multiple of 2^swz_y and gridDim.z of 2^swz_z (the host checks).
__device__ __forceinline__ void block_yz(const int swz_y, const int swz_z, const int swz_nty, int &by, int &bz)
{
    by = (int)blockIdx.y;
    bz = (int)blockIdx.z;
    if (swz_z >= 0) {
        const unsigned l = blockIdx.y + gridDim.y * blockIdx.z;
        const unsigned w = l & ((1u << (swz_y + swz_z)) - 1u), t = l >> (swz_y + swz_z);
        by = (int)(((t & ((1u << swz_nty) - 1u)) << swz_y) + (w & ((1u << swz_y) - 1u)));
        bz = (int)(((t >> swz_nty) << swz_z) + (w >> swz_y));
    }
}

// One cell-group of one iteration: the thread owns the VEC cells x0 .. x0+VEC-1 of row (y, z).
// `mask` = the lanes of the warp that execute this function (shuffles of the VEC > 1 variants).
template <typename T, int VEC, bool FAST, bool MACRO, int PEER, int LM>
__device__ __forceinline__ void step_pull_cells(const StepArgs<T> &a, const int x0, const int y, const int z,
                                                const int rowbits, const unsigned mask)
{
    using V = Pack<T, VEC>;
    const int dim = LBM_U_DIM(a);
    const Layout lay = LBM_U_LAY(a);
    const Consts<T> c = LBM_U_CONSTS(T, a);
    const int zl = z - LBM_U_ZS0(a);
    const int tx = threadIdx.x;

    const long long plane = (long long)dim * dim;
    const long long rowid = (long long)y * dim + (long long)zl * plane;  // id of the row's first cell
    const long long id0 = x0 + rowid;
    const long long qp = lay.qpitch();

    // ---- addresses: per-thread bases, per-direction offsets are uniform ----
    // LM_ROWS / LM_SOA: b0 = element index of (x0, y, z, q = 0); bm / bp = of (x0 - 1, ..) and (x0 + VEC, ..),
    // clamped into the row (the clamped values are only ever consumed by WALL cells).
    // LM_BLOCKROWS: one full base b0; nbc[(ey+1)*3 + ez+1] = byte correction for the CSoA block boundaries between
    // row (y, z) and row (y - ey, z - ez): (blocks crossed) * blk18.  The x +- 1 neighbours are the adjacent
    // elements (a row never leaves its CSoA block; reading one element past either end of a row stays inside
    // the allocation and is only consumed by WALL cells).
    long long b0, bm = 0, bp = 0;
    long long nbc[9];
    if constexpr (LM == LM_ROWS) {
        const long long rowbase = rowid * Q;
        const int xm = x0 > 0 ? x0 - 1 : 0;
        const int xp = x0 + VEC < dim ? x0 + VEC : dim - 1;
        const int sm = (int)lay.smod;
        b0 = rowbase + ((((x0 >> lay.sdiv) * Q) << lay.sdiv) + (x0 & sm));
        bm = rowbase + ((((xm >> lay.sdiv) * Q) << lay.sdiv) + (xm & sm));
        bp = rowbase + ((((xp >> lay.sdiv) * Q) << lay.sdiv) + (xp & sm));
    } else if constexpr (LM == LM_SOA) {
        b0 = id0;
        bm = id0 - 1;  // live rows have y >= 1, so id0 >= DIM
        bp = id0 + VEC;
    } else if constexpr (LM == LM_BLOCKROWS) {
        b0 = lay.base(rowid) + x0;
        const int rsh = LBM_U_ROWSHIFT(a);
        const int r = (y + zl * dim) & LBM_U_ROWMASK(a);  // row inside its CSoA block
        const long long blk18 = LBM_U_BLK18(T, a);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
                nbc[(dy + 1) * 3 + dz + 1] = (long long)((r - dy - dz * dim) >> rsh) * blk18;  // arithmetic shift: floor
    } else {
        b0 = lay.base(id0);
    }
    const char *const s0 = reinterpret_cast<const char *>(a.src + b0);
    const char *const sm1 = reinterpret_cast<const char *>(a.src + bm);
    const char *const sp1 = reinterpret_cast<const char *>(a.src + bp);
    (void)sm1;
    (void)sp1;
    (void)nbc;

    // ---- gather: f[q][j] = G(x0 + j - ex, y - ey, z - ez, q) ----
    T f[Q][VEC];
    if constexpr (VEC == 1) {
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            if constexpr (LM == LM_GENERIC) {
                const long long sid = id0 - ex(q) - (long long)ey(q) * dim - (long long)ez(q) * plane;
                f[q][0] = a.src[lay.base(sid) + q * qp];
            } else if constexpr (LM == LM_BLOCKROWS) {
                const char *p = s0 + nbc[(ey(q) + 1) * 3 + ez(q) + 1];
                f[q][0] = *(reinterpret_cast<const T *>(p + LBM_U_GOFF(T, a, q, LM)) - ex(q));
            } else {
                const char *p = ex(q) == 0 ? s0 : (ex(q) == 1 ? sm1 : sp1);
                f[q][0] = ld_gather<T, ex(q)>(reinterpret_cast<const T *>(p + LBM_U_GOFF(T, a, q, LM)));
            }
        });
    } else {
        // lanes of one row segment are consecutive lanes of the warp
        const int seg = blockDim.x < 32 ? blockDim.x : 32;
        const bool seg_first = (tx & (seg - 1)) == 0;
        const bool seg_last = (tx & (seg - 1)) == seg - 1;
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            V v;
            const T *pe_m, *pe_p;  // where the element left of / right of the vector lives
            if constexpr (LM == LM_GENERIC) {
                const long long sid = id0 - (long long)ey(q) * dim - (long long)ez(q) * plane;
                const T *p = a.src + q * qp;
                v = *reinterpret_cast<const V *>(p + lay.base(sid));
                pe_m = p + (ex(q) == 1 ? lay.base(sid - 1) : 0);
                pe_p = p + (ex(q) == -1 ? lay.base(sid + VEC) : 0);
            } else if constexpr (LM == LM_BLOCKROWS) {
                const char *p = s0 + nbc[(ey(q) + 1) * 3 + ez(q) + 1];
                const T *pv = reinterpret_cast<const T *>(p + LBM_U_GOFF(T, a, q, LM));
                v = *reinterpret_cast<const V *>(pv);
                pe_m = pv - 1;
                pe_p = pv + VEC;
            } else {
                v = *reinterpret_cast<const V *>(s0 + LBM_U_GOFF(T, a, q, LM));
                pe_m = reinterpret_cast<const T *>(sm1 + LBM_U_GOFF(T, a, q, LM));
                pe_p = reinterpret_cast<const T *>(sp1 + LBM_U_GOFF(T, a, q, LM));
            }
            if constexpr (ex(q) == 0) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) f[q][j] = v.v[j];
            } else if constexpr (ex(q) == 1) {
                T e = __shfl_up_sync(mask, v.v[VEC - 1], 1);
                if (seg_first) e = (x0 > 0) ? *pe_m : T(0);
                f[q][0] = e;
#pragma unroll
                for (int j = 1; j < VEC; ++j) f[q][j] = v.v[j - 1];
            } else {
                T e = __shfl_down_sync(mask, v.v[0], 1);
                if (seg_last) e = (x0 + VEC < dim) ? *pe_p : T(0);
                f[q][VEC - 1] = e;
#pragma unroll
                for (int j = 0; j < VEC - 1; ++j) f[q][j] = v.v[j + 1];
            }
        });
    }

    // Warps that lie entirely inside the fluid (interior row, no x-edge cell: 6 of 8 warps per row at
    // DIM = 256) skip the ghost-constant fix-up and the per-cell classification.
    const bool all_fluid =
        __all_sync(mask, rowbits == CT_NONE && x0 >= 2 && x0 + VEC - 1 <= dim - 3) != 0;

    // ---- ghost constants for what cells x = 1 / x = DIM-2 would gather from the x walls ----
    if (!all_fluid) {
        const int lid = (rowbits & CT_FRONT) ? 1 : 0;  // the gathering cell started with u = (U,0,0)
        constexpr int JL = VEC >= 2 ? 1 : 0;           // x == 1     is element JL of the thread at x0 == XL
        constexpr int XL = VEC >= 2 ? 0 : 1;
        constexpr int JR = VEC >= 2 ? VEC - 2 : 0;     // x == DIM-2 is element JR of the thread at x0 == DIM-XR
        constexpr int XR = VEC >= 2 ? VEC : 2;
        if (x0 == XL) {
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (ex(q) == 1) f[q][JL] = a.stale[lid][q];
            });
        }
        if (x0 == dim - XR) {
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (ex(q) == -1) f[q][JR] = a.stale[lid][q];
            });
        }
    }

    // ---- boundary conditions + collision, cell by cell ----
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    T m_rho[VEC], m_ux[VEC], m_uy[VEC], m_uz[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int t = all_fluid ? (int)CT_FLUID : cell_type_from_row(rowbits, x0 + j, dim);
        T fc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) fc[q] = f[q][j];
        T rho = nan, ux = nan, uy = nan, uz = nan;
        if (t == CT_FLUID) {
            collide_fluid<T, FAST>(fc, c, rho, ux, uy, uz);
        } else if (t & CT_MOVING) {
            collide_lid<T, FAST>(fc, c, rho);
            ux = c.u_lid;
            uy = T(0);
            uz = T(0);
        } else if (is_bounceback(t)) {
            bounce_back<T>(fc);
        }  // CORNER: pass-through; WALL (x = 0, DIM-1): don't care
#pragma unroll
        for (int q = 0; q < Q; ++q) f[q][j] = fc[q];
        m_rho[j] = rho;
        m_ux[j] = ux;
        m_uy[j] = uy;
        m_uz[j] = uz;
    }

    // ---- store G_k ----
    char *const d0 = reinterpret_cast<char *>(a.dst + b0);
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        V v;
#pragma unroll
        for (int j = 0; j < VEC; ++j) v.v[j] = f[q][j];
        *reinterpret_cast<V *>(d0 + LBM_U_SOFF(T, a, q)) = v;
    });
    if constexpr (PEER != PEER_NONE) {
        // fused halo exchange: the five populations that cross a slab face also go straight into the
        // neighbour's halo plane (NVLink stores when the neighbour is a peer device); plane-uniform branches
        if (a.peer_lo != nullptr && z == a.z_own_begin) {
            T *const p = a.peer_lo + lay.base(x0 + (long long)y * dim + a.peer_lo_plane * plane);
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (ez(q) == -1) {
                    V v;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v.v[j] = f[q][j];
                    *reinterpret_cast<V *>(p + q * qp) = v;
                }
            });
        }
        if (a.peer_hi != nullptr && z == a.z_own_end - 1) {
            T *const p = a.peer_hi + lay.base(x0 + (long long)y * dim + a.peer_hi_plane * plane);
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (ez(q) == 1) {
                    V v;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v.v[j] = f[q][j];
                    *reinterpret_cast<V *>(p + q * qp) = v;
                }
            });
        }
    }

    // ---- macro store (kernels.cl:383-388): only FLUID / MOVING cells; the others keep their NaN ----
    if constexpr (MACRO) {
        const bool row_stores = !(rowbits & (CT_TOP | CT_BOTTOM | CT_BACK));
        if (row_stores) {
            const long long n_local = LBM_U_NLOCAL(a);
            V r, vx, vy, vz;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                r.v[j] = m_rho[j];
                vx.v[j] = m_ux[j];
                vy.v[j] = m_uy[j];
                vz.v[j] = m_uz[j];
            }
            *reinterpret_cast<V *>(a.rho + id0) = r;
            *reinterpret_cast<V *>(a.u + id0) = vx;
            *reinterpret_cast<V *>(a.u + n_local + id0) = vy;
            *reinterpret_cast<V *>(a.u + 2 * n_local + id0) = vz;
        }
    }
}

// One iteration.  Thread (tx, ty, tz) of block (bx, by, bz) owns the VEC cells
//   x0 .. x0+VEC-1 = (blockIdx.x*bx + tx)*VEC .. ,  y = blockIdx.y*by + ty,  plane number blockIdx.z*bz + tz
// of the launch's plane list (StepArgs::zmap*, z_begin, z_end).
// Requirements (checked by the host): bx a power of two, bx*VEC divides DIM, by divides DIM,
// VEC divides stride (so every vector is aligned and inside one CSoA run); PEER_FLAGS: bz == 1.
template <typename T, int VEC, bool FAST, bool MACRO, int PEER, int LM>
__device__ __forceinline__ void step_pull_body(const StepArgs<T> &a)
{
    const int dim = LBM_U_DIM(a);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    int by, bz;
    block_yz(a.swz_y, a.swz_z, a.swz_nty, by, bz);
    const int y = by * blockDim.y + threadIdx.y;
    const int zi = bz * blockDim.z + threadIdx.z;
    int z;
    bool zok = true;
    if (zi < a.zmap_n) {
        z = zi == 0 ? a.zmap0 : a.zmap1;
    } else {
        z = a.z_begin + (zi - a.zmap_n);
        zok = z < a.z_end;
    }
    const int rowbits = zok ? row_bits(y, z, dim) : CT_WALL;
    const bool live = rowbits != CT_WALL;  // wall rows: never read, never written
    const unsigned mask = __ballot_sync(0xffffffffu, live);

    if constexpr (PEER == PEER_FLAGS) {
        // block-uniform (bz == 1): is this block part of a boundary plane that talks to a neighbour?
        const bool lo_face = a.peer_lo != nullptr && z == a.z_own_begin;
        const bool hi_face = a.peer_hi != nullptr && z == a.z_own_end - 1;
        if (lo_face || hi_face) {
            const bool first = threadIdx.x == 0 && threadIdx.y == 0;
            // the halo plane I gather from holds the neighbour's populations of the previous phase, and the
            // neighbour has finished reading the halo plane I am about to overwrite
            if (first) {
                if (lo_face) slab_wait(a.sync.flag_in + 0, a.sync.wait_epoch, a.sync.timeout_ns, a.sync.error);
                if (hi_face) slab_wait(a.sync.flag_in + 1, a.sync.wait_epoch, a.sync.timeout_ns, a.sync.error);
            }
            __syncthreads();
            if (live) step_pull_cells<T, VEC, FAST, MACRO, PEER_STORE, LM>(a, x0, y, z, rowbits, mask);
            __syncthreads();
            if (first) {
                const unsigned n_blocks = gridDim.x * gridDim.y;
                if (lo_face) slab_signal(a.sync.count + 0, n_blocks, a.sync.flag_out[0], a.sync.signal_epoch);
                if (hi_face) slab_signal(a.sync.count + 1, n_blocks, a.sync.flag_out[1], a.sync.signal_epoch);
            }
        } else {
            // interior planes (all but two of the launch): exactly the code of the single-device kernel, so that
            // the neighbour handling above costs them neither instructions nor registers
            if (live) step_pull_cells<T, VEC, FAST, MACRO, PEER_NONE, LM>(a, x0, y, z, rowbits, mask);
        }
    } else {
        if (!live) return;
        step_pull_cells<T, VEC, FAST, MACRO, PEER, LM>(a, x0, y, z, rowbits, mask);
    }
}

// Occupancy.  The step kernel is latency-bound as much as bandwidth-bound: its throughput follows the number
// of resident warps (measured, profiles/r02_experiments.md section 1: the same code at 72 instead of 40 registers per
// thread -- 3 instead of 6 blocks per SM -- runs 22 % slower), so the register budget is pinned per variant
// instead of being left to ptxas' heuristics:
//   fp32, 1 cell/thread     6 blocks of 256 threads per SM (40 registers; the kernel needs exactly that);
//                           the PEER_FLAGS kernel too: only its two boundary planes run the neighbour code
//                           and take the few spills it costs at 40 registers
//   fp64, 1 cell/thread     3 blocks (<= 80 registers; needs 73)
//   fp32, 2 cells/thread    3 blocks;  4 cells/thread and fp64 2 cells/thread: 1 block (117-134 registers)
// (boundary-only launches -- PEER_STORE -- and launches with macro stores are rare: one block less.)
#ifndef LBM_MINB_F32
#define LBM_MINB_F32 6
#endif
#ifndef LBM_MINB_F64
#define LBM_MINB_F64 3
#endif
template <typename T, int VEC, bool MACRO, int PEER>
__host__ __device__ constexpr int step_min_blocks()
{
    constexpr bool rare = MACRO || PEER == PEER_STORE;
    if (VEC == 1) return (sizeof(T) == 4 ? LBM_MINB_F32 : LBM_MINB_F64) - (rare ? 1 : 0);
    if (VEC == 2 && sizeof(T) == 4) return 3;
    return 1;
}

template <typename T, int VEC, bool FAST, bool MACRO, int PEER, int LM>
__global__ void __launch_bounds__(256, step_min_blocks<T, VEC, MACRO, PEER>()) step_pull_kernel(const StepArgs<T> a)
{
    step_pull_body<T, VEC, FAST, MACRO, PEER, LM>(a);
}

// ------------------------------------------------------------------------------------------------
// In-place AA-pattern variant (Bailey et al. 2009): ONE lattice, half the memory of the two-lattice
// scheme, same 19 loads + 19 stores per cell.  Iterations alternate between
//   LOCAL step (1st, 3rd, ..):  read the cell's own 19 slots (the lattice is in the reference's
//                               pre-collision form S), collide, store f'_q into the cell's slot opp(q);
//   SHIFT step (2nd, 4th, ..):  f_q = A(c - e_q, opp(q)), collide, store f'_q into A(c + e_q, q),
//                               which leaves the lattice in the S form again.
// In either step a thread reads and writes exactly the same 19 slots, so the update is race free.
// Populations that would arrive from a WALL cell are never-written slots in the reference:
// here they are replaced explicitly by their initial value (the `stale` constants) using the cell's
// side bits, for all six walls.
//
// Occupancy: left to ptxas by default (fp32: 40 registers LOCAL / 48 SHIFT, fp64: 73 / 80, no spills).  A bound of
// 6 blocks per SM makes ptxas spill 196 bytes in the LOCAL step at the very same 40 registers, so only the SHIFT
// step takes an optional bound (-DLBM_MINB_AA_SHIFT_F32=6: 40 registers, 16-28 bytes spilled; an A/B knob).
#ifndef LBM_MINB_AA_SHIFT_F32
#define LBM_MINB_AA_SHIFT_F32 0
#endif
#ifndef LBM_MINB_AA_SHIFT_F64
#define LBM_MINB_AA_SHIFT_F64 0
#endif
template <typename T, bool SHIFT>
__host__ __device__ constexpr int aa_min_blocks()
{
    return !SHIFT ? 0 : (sizeof(T) == 4 ? LBM_MINB_AA_SHIFT_F32 : LBM_MINB_AA_SHIFT_F64);
}

// In-place kernels, after the gather: ghost constants for populations that would arrive from a WALL cell,
// boundary condition / collision of cell (x, row).  Returns the cell type.
template <typename T, bool FAST>
__device__ __forceinline__ int aa_collide_cell(const StepArgs<T> &a, T (&f)[Q], const int rowbits, const int x,
                                               const int dim, T &rho, T &ux, T &uy, T &uz)
{
    // raw side bits (before the CORNER / MOVING rewriting of get_cell_type)
    int side = rowbits;
    if (x == 1) side |= CT_LEFT;
    if (x == dim - 2) side |= CT_RIGHT;
    if (side != CT_NONE) {
        const int lid = (side & CT_FRONT) ? 1 : 0;
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int from_wall = (ex(q) == 1 ? CT_LEFT : 0) | (ex(q) == -1 ? CT_RIGHT : 0) |
                                      (ey(q) == 1 ? CT_BOTTOM : 0) | (ey(q) == -1 ? CT_TOP : 0) |
                                      (ez(q) == 1 ? CT_BACK : 0) | (ez(q) == -1 ? CT_FRONT : 0);
            if constexpr (from_wall != 0) {
                if (side & from_wall) f[q] = a.stale[lid][q];
            }
        });
    }

    const int t = cell_type_from_row(rowbits, x, dim);
    if (t == CT_FLUID) {
        collide_fluid<T, FAST>(f, a.c, rho, ux, uy, uz);
    } else if (t & CT_MOVING) {
        collide_lid<T, FAST>(f, a.c, rho);
        ux = a.c.u_lid;
        uy = T(0);
        uz = T(0);
    } else if (is_bounceback(t)) {
        bounce_back<T>(f);
    }
    return t;
}

template <typename T, bool FAST, bool MACRO, bool SHIFT, int LM>
__global__ void __launch_bounds__(256, aa_min_blocks<T, SHIFT>()) step_aa_kernel(const StepArgs<T> a)
{
    const int dim = a.dim;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    int by, bz;
    block_yz(a.swz_y, a.swz_z, a.swz_nty, by, bz);
    const int y = by * blockDim.y + threadIdx.y;
    const int z = a.z_begin + bz * blockDim.z + threadIdx.z;
    if (z >= a.z_end) return;
    const int rowbits = row_bits(y, z, dim);
    if (rowbits == CT_WALL || x == 0 || x >= dim - 1) return;  // WALL cells neither read nor write

    const long long plane = (long long)dim * dim;
    const long long rowid = (long long)y * dim + (long long)(z - a.zs0) * plane;
    const long long id0 = x + rowid;
    const long long qp = a.lay.qpitch();
    T *const lat = a.dst;

    long long b0, bm = 0, bp = 0;
    long long nbc[9];  // LM_BLOCKROWS, SHIFT: block-boundary correction towards row (y - dy, z - dz), see step_pull_cells
    if constexpr (LM == LM_ROWS) {
        const long long rowbase = rowid * Q;
        const int sm = (int)a.lay.smod;
        b0 = rowbase + ((((x >> a.lay.sdiv) * Q) << a.lay.sdiv) + (x & sm));
        if constexpr (SHIFT) {
            bm = rowbase + (((((x - 1) >> a.lay.sdiv) * Q) << a.lay.sdiv) + ((x - 1) & sm));
            bp = rowbase + (((((x + 1) >> a.lay.sdiv) * Q) << a.lay.sdiv) + ((x + 1) & sm));
        }
    } else if constexpr (LM == LM_SOA) {
        b0 = id0;
        bm = id0 - 1;
        bp = id0 + 1;
    } else if constexpr (LM == LM_BLOCKROWS) {
        b0 = a.lay.base(rowid) + x;
        if constexpr (SHIFT) {
            const int r = (y + (z - a.zs0) * dim) & a.row_mask;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz)
                    nbc[(dy + 1) * 3 + dz + 1] = (long long)((r - dy - dz * dim) >> a.row_shift) * a.blk18;
        }
    } else {
        b0 = a.lay.base(id0);
    }
    (void)nbc;
    char *const c0 = reinterpret_cast<char *>(lat + b0);
    char *const cm = reinterpret_cast<char *>(lat + bm);
    char *const cp = reinterpret_cast<char *>(lat + bp);

    // where population q of this cell is read from / written to
    auto slot = [&](auto qc, bool for_store) -> T * {
        constexpr int q = decltype(qc)::value;
        if constexpr (!SHIFT) {
            // LOCAL: read slot q, write slot opp(q), both in this cell
            return reinterpret_cast<T *>(c0 + (for_store ? a.soff[opp(q)] : a.soff[q]));
        } else if constexpr (LM == LM_GENERIC) {
            const long long nid = for_store
                ? id0 + ex(q) + (long long)ey(q) * dim + (long long)ez(q) * plane
                : id0 - ex(q) - (long long)ey(q) * dim - (long long)ez(q) * plane;
            return lat + a.lay.base(nid) + (for_store ? q : opp(q)) * qp;
        } else if constexpr (LM == LM_BLOCKROWS) {
            // SHIFT: read (c - e_q, opp(q)), write (c + e_q, q)
            if (for_store) return reinterpret_cast<T *>(c0 + nbc[(-ey(q) + 1) * 3 - ez(q) + 1] + a.poff[q]) + ex(q);
            return reinterpret_cast<T *>(c0 + nbc[(ey(q) + 1) * 3 + ez(q) + 1] + a.goff[q]) - ex(q);
        } else {
            // SHIFT: read (c - e_q, opp(q)), write (c + e_q, q): the same address for q and opp(q) swapped
            if (for_store) return reinterpret_cast<T *>((ex(q) == 0 ? c0 : (ex(q) == 1 ? cp : cm)) + a.poff[q]);
            return reinterpret_cast<T *>((ex(q) == 0 ? c0 : (ex(q) == 1 ? cm : cp)) + a.goff[q]);
        }
    };

    T f[Q];
    static_for<Q>([&](auto qc) { f[decltype(qc)::value] = *slot(qc, false); });

    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    T rho = nan, ux = nan, uy = nan, uz = nan;
    const int t = aa_collide_cell<T, FAST>(a, f, rowbits, x, dim, rho, ux, uy, uz);

    static_for<Q>([&](auto qc) { *slot(qc, true) = f[decltype(qc)::value]; });

    if constexpr (MACRO) {
        if (is_collision(t)) {
            a.rho[id0] = rho;
            a.u[id0] = ux;
            a.u[a.n_local + id0] = uy;
            a.u[2 * a.n_local + id0] = uz;
        }
    }
}

// `initialize` for the AA variant: the lattice starts in the reference's own form,
// A(c, q) = f_eq_q(1, u0(c)) (kernels.cl:303-317); rho/u as in init_kernel.
template <typename T>
__global__ void __launch_bounds__(256) init_aa_kernel(const InitArgs<T> a)
{
    const int dim = a.dim;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= dim || y >= dim || zl >= a.nz_local) return;
    const int z = a.zs0 + zl;
    const long long id = x + (long long)y * dim + (long long)zl * dim * dim;
    const int t = cell_type(x, y, z, dim);
    const bool keep = is_collision(t);
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    const T ux0 = has_front_bit(x, y, z, dim) ? a.c.u_lid : T(0);
    a.rho[id] = keep ? T(1) : nan;
    a.u[id] = keep ? ux0 : nan;
    a.u[a.n_local + id] = keep ? T(0) : nan;
    a.u[2 * a.n_local + id] = keep ? T(0) : nan;
    const long long b = a.lay.base(id);
    const long long qp = a.lay.qpitch();
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        a.f0[b + q * qp] = init_feq<T, q>(a.c, ux0);
    });
}

// Reference view of the AA lattice (for -f): `swapped` != 0 after a LOCAL step (slot opp(q) of the
// source cell holds the value), 0 when the lattice is in S form.
template <typename T>
__global__ void __launch_bounds__(256) reference_view_aa_kernel(const T *__restrict__ lat, T *__restrict__ out,
                                                                int dim, Layout lay, Consts<T> c, int swapped,
                                                                int pristine)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= dim || y >= dim || z >= dim) return;
    const long long plane = (long long)dim * dim;
    const long long gid = x + (long long)y * dim + (long long)z * plane;
    const int t = cell_type(x, y, z, dim);
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    const T ux0 = has_front_bit(x, y, z, dim) ? c.u_lid : T(0);
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const int sx = x - ex(q), sy = y - ey(q), sz = z - ez(q);
        const bool inside = sx >= 0 && sx < dim && sy >= 0 && sy < dim && sz >= 0 && sz < dim;
        const bool src_live = inside && cell_type(sx, sy, sz, dim) != CT_WALL;
        T v;
        if (src_live && !(pristine && t == CT_WALL)) {
            if (swapped) {
                const long long sid = sx + (long long)sy * dim + (long long)sz * plane;
                v = lat[lay.base(sid) + opp(q) * lay.qpitch()];
            } else {
                v = lat[lay.base(gid) + q * lay.qpitch()];
            }
        } else if (t == CT_WALL) {
            v = nan;
        } else {
            v = init_feq<T, q>(c, ux0);
        }
        out[lay.base(gid) + q * lay.qpitch()] = v;
    });
}

// Reference view of the lattice the next iteration reads (for the -f dump, lbmcl.hpp:206-258):
//   S(c, q) = G(c - e_q, q)        if c - e_q is inside the cube and not WALL   (a pushed value)
//           = NaN                  else if c is WALL                            (kernels.cl:311-317)
//           = f_eq_q(1, u0(c))     otherwise                                    (never-written slot)
// `pristine` (no iteration executed yet): WALL cells are all NaN because nothing has been pushed.
// Output in the reference's CSoA(stride) order over the owned planes, global cell ids.
template <typename T>
__global__ void __launch_bounds__(256) reference_view_kernel(const T *__restrict__ g, T *__restrict__ out,
                                                             int dim, int zs0, int nz_local, int z_begin,
                                                             int z_end, Layout lay_local, Layout lay_global,
                                                             Consts<T> c, int pristine)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = z_begin + blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= dim || y >= dim || z >= z_end) return;
    const long long plane = (long long)dim * dim;
    const long long gid = x + (long long)y * dim + (long long)z * plane;
    const int t = cell_type(x, y, z, dim);
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    const T ux0 = has_front_bit(x, y, z, dim) ? c.u_lid : T(0);
    static_for<Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const int sx = x - ex(q), sy = y - ey(q), sz = z - ez(q);
        const bool inside = sx >= 0 && sx < dim && sy >= 0 && sy < dim && sz >= 0 && sz < dim;
        T v;
        if (inside && cell_type(sx, sy, sz, dim) != CT_WALL && !(pristine && t == CT_WALL)) {
            const int szl = sz - zs0;  // source plane must be stored locally (owned or halo)
            const long long sid = sx + (long long)sy * dim + (long long)szl * plane;
            v = (szl >= 0 && szl < nz_local) ? g[lay_local.base(sid) + q * lay_local.qpitch()] : nan;
        } else if (t == CT_WALL) {
            v = nan;
        } else {
            v = init_feq<T, q>(c, ux0);
        }
        out[lay_global.base(gid) + q * lay_global.qpitch()] = v;
    });
}

// Dense halo transport for the one-process-per-device mode.  A packed halo is [5][DIM][DIM]
// (population slot, y, x).  kDown / kUp list the crossing populations.
__device__ __constant__ const int kDown[5] = { 5, 11, 12, 13, 14 };   // e_z = -1
__device__ __constant__ const int kUp[5] = { 6, 15, 16, 17, 18 };     // e_z = +1

// dir_up != 0: populations with e_z = +1; plane_local: plane of the local storage to read / write
template <typename T, bool PACK>
__global__ void __launch_bounds__(256) halo_kernel(T *__restrict__ lattice, T *__restrict__ dense, int dim,
                                                   long long plane_local, Layout lay, int dir_up)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    if (x >= dim) return;
    const int q = dir_up ? kUp[s] : kDown[s];
    const long long plane = (long long)dim * dim;
    const long long id = x + (long long)y * dim + plane_local * plane;
    const long long li = lay.base(id) + q * lay.qpitch();
    const long long di = x + (long long)y * dim + (long long)s * plane;
    if (PACK) dense[di] = lattice[li];
    else lattice[li] = dense[di];
}

}  // namespace lbm
