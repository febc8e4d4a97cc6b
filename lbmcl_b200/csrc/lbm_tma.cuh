// TMA-fed variant of the D3Q19 step (sm_90+/sm_100a: cp.async.bulk.tensor + mbarrier), warp-specialised.
//
// Same function as step_pull_kernel (two-lattice pull, post-collision storage, see lbm_kernels.cuh);
// what changes is how the bytes move:
//
//   * the lattice is described to the TMA unit as a 3-D tensor  [block][q][stride]  (the CSoA layout of
//     kernels.cl:64 read literally), so ONE bulk-tensor copy fetches the TX populations of one direction
//     of one x-row segment -- whatever the stride -- into a contiguous shared-memory row;
//   * persistent CTAs (a few per SM) walk the live rows of the launch.  Each CTA is TX/32 CONSUMER warps (one
//     thread per cell of a row tile) plus one PRODUCER warp whose elected lane keeps a ring of NS input tiles in
//     flight: 19 bulk loads per tile, completion on the stage's `full` mbarrier (expect_tx); a stage is handed
//     back through its `empty` mbarrier, on which every consumer warp arrives as soon as its 19 values are in
//     registers -- the loads of tile i + NS are under way while tile i is still being collided;
//   * a consumer reads its 19 populations with conflict-free LDS at compile-time offsets (the x +- 1 shift is
//     just an index), applies the reference's BC / collision, and writes the results into its WARP's private
//     output buffer, laid out as the warp's piece of the CSoA lattice ([q][32] for stride >= 32): one
//     bulk-tensor store per warp and tile (19 x 128 bytes in fp32), double-buffered with
//     cp.async.bulk.wait_group.read.  Consumer warps never meet at a CTA-wide barrier.
//
// Round 1's version of this kernel (one elected thread issuing 19 loads + 19 stores between two __syncthreads per
// tile, tiles updated in place) was bound by that per-CTA critical path: 5.5 TB/s, profiles/r01_tma_experiment.md.
//
// Requirements (checked by the host): LM_ROWS layout (stride <= DIM), stride*sizeof(T) >= 16 bytes,
// DIM >= 32.  Rows wider than TX cells are cut into segments; the two segment-edge threads fetch their
// out-of-tile neighbour with a plain load.
#pragma once

#include <cuda.h>

#include "lbm_kernels.cuh"

namespace lbm {

template <typename T>
struct TmaArgs {
    const T *__restrict__ src;  // for the segment-edge loads only
    T *__restrict__ dst;        // direct-store build: the lattice written
    T *__restrict__ rho;
    T *__restrict__ u;
    int dim;
    int zs0;
    int z_first;        // first live global plane of this launch (already clipped to [1, DIM-2])
    int n_tiles;        // live rows x segments of this launch
    int n_xseg;         // DIM / TX
    int ns;             // input ring depth
    int osdiv;          // log2(S'), S' = min(stride, 32): a warp's output buffer is [32/S'][Q][S']
    long long n_local;
    Layout lay;
    Consts<T> c;
    T stale[2][Q];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// How the results leave the SM:
//   LBM_TMA_DIRECT_STORE = 1 (default)  plain coalesced stores from registers (19 full lines per warp, as in
//                                       step_pull_kernel): shared memory holds input tiles only, so three CTAs
//                                       (24 consumer warps in fp32) fit an SM;
//   LBM_TMA_DIRECT_STORE = 0            each consumer warp stages its 32 cells x 19 directions in a private,
//                                       double-buffered tile laid out as its piece of the CSoA lattice and sends it
//                                       with ONE bulk-tensor store (cp.async.bulk.wait_group.read before reuse):
//                                       twice the shared memory per warp, two CTAs per SM.
// Measured on B200 (profiles/r02_tma_experiment.md): the kernel is bound by the number of resident consumer
// warps (the strict collision is one long dependency chain per thread), i.e. by shared memory per warp.
#ifndef LBM_TMA_DIRECT_STORE
#define LBM_TMA_DIRECT_STORE 1
#endif
#ifndef LBM_TMA_MINB
#define LBM_TMA_MINB (LBM_TMA_DIRECT_STORE ? 3 : 2)
#endif
template <typename T>
__host__ __device__ constexpr int tma_min_blocks()
{
    return sizeof(T) == 4 ? LBM_TMA_MINB : 2;
}

// Bounded mbarrier wait: a TMA fault or a protocol error must not hang the GPU.  Returns false when the limit
// was hit or another warp of the CTA has already given up (`*abort_flag`).
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity, volatile int *abort_flag)
{
    if (mbar_try_wait(bar, parity)) return true;
    for (long long spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if ((spins & 0x3ff) == 0x3ff && (*abort_flag != 0 || spins > (1ll << 22))) return false;
    }
    return true;
}

// blockDim.x = TX + 32: threads [0, TX) are the consumers (TX = cells per row tile: 32, 64, 128 or 256 -- the
// launch's choice, not a template parameter), warp TX/32 is the producer.
template <typename T, bool FAST, bool MACRO>
__global__ void __launch_bounds__(288, tma_min_blocks<T>()) step_tma_kernel(const __grid_constant__ CUtensorMap map_src,
                                                                            const __grid_constant__ CUtensorMap map_dst,
                                                                            const TmaArgs<T> a, int *__restrict__ error_flag)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TX = (int)blockDim.x - 32;
    const int NW = TX >> 5;  // consumer warps
    const uint32_t STAGE_BYTES = (uint32_t)(Q * TX * sizeof(T));
    constexpr uint32_t OUT_BYTES = LBM_TMA_DIRECT_STORE ? 0u : (uint32_t)(Q * 32 * sizeof(T));  // one warp's output tile
    const int NS = a.ns;
    T *const ring = reinterpret_cast<T *>(smem_raw);                                         // [NS][Q][TX]
    T *const outb = reinterpret_cast<T *>(smem_raw + (size_t)NS * STAGE_BYTES);              // [NW][2][Q*32]
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NS * STAGE_BYTES + (size_t)NW * 2 * OUT_BYTES);
    uint64_t *const empty = full + NS;
    int4 *const coords = reinterpret_cast<int4 *>(empty + NS);                               // [NS] (x0, y, z, -)
    int *const abort_flag = reinterpret_cast<int *>(coords + NS);
    (void)outb;

    const int tx = threadIdx.x;
    const int warp = tx >> 5, lane = tx & 31;
    const int dim = a.dim;
    const long long plane = (long long)dim * dim;

    if (tx == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (uint32_t)NW);
        }
        *abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier of the kernel

    const int n_my = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == NW) {
        // ---------------- producer: one thread ----------------
        if (lane != 0) return;
        const int rows = dim - 2;  // live y per plane
        for (int i = 0; i < n_my; ++i) {
            const int s = i % NS;
            if (i >= NS && !mbar_wait_bounded(&empty[s], (uint32_t)(i / NS - 1) & 1u, abort_flag)) {
                *abort_flag = 1;
                if (error_flag) atomicExch(error_flag, 1);
                return;
            }
            // tile -> (row segment, y, z); the consumers read the coordinates from the stage
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int xs = t % a.n_xseg;
            const int r = t / a.n_xseg;
            const int x0 = xs * TX, y = 1 + r % rows, z = a.z_first + r / rows;
            coords[s] = make_int4(x0, y, z, 0);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);  // release: orders the coordinates before the completion
            T *const dst = ring + (size_t)s * Q * TX;
            // tensor coordinates of the TX values of direction q in row (yy, zz) starting at x0:
            //   c0 = offset inside the CSoA run, c1 = q, c2 = CSoA block index
            const int c0 = x0 & (int)a.lay.smod;
            const int xb = x0 >> a.lay.sdiv;
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                const long long row = (long long)(y - ey(q)) * dim + (long long)(z - ez(q) - a.zs0) * plane;
                tma_load_3d(dst + q * TX, &map_src, &full[s], c0, q, (int)(row >> a.lay.sdiv) + xb);
            });
        }
        return;
    }

    // ---------------- consumers: one thread per cell of the tile ----------------
    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
#if !LBM_TMA_DIRECT_STORE
    T *const my_out = outb + (size_t)warp * 2 * (Q * 32);
    // my slot inside the warp's output tile [32/S'][Q][S']
    const int so = ((lane >> a.osdiv) * Q << a.osdiv) + (lane & ((1 << a.osdiv) - 1));
    const int sq = 1 << a.osdiv;  // distance between consecutive q
#endif
    for (int i = 0; i < n_my; ++i) {
        const int s = i % NS;
        const T *const tile = ring + (size_t)s * Q * TX;

        if (!mbar_wait_bounded(&full[s], (uint32_t)(i / NS) & 1u, abort_flag)) {
            *abort_flag = 1;
            if (lane == 0 && error_flag) atomicExch(error_flag, 1);
            break;
        }
        const int4 co = coords[s];
        const int x0 = co.x, y = co.y, z = co.z;
        const int x = x0 + tx;

        T f[Q];
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            f[q] = tile[q * TX + tx - ex(q)];  // tx -+ 1 outside the tile: fixed below (or a WALL cell)
        });
        // this warp is done with the stage: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);

        const long long id0 = x + (long long)y * dim + (long long)(z - a.zs0) * plane;
        if (a.n_xseg > 1) {
            const long long qp = a.lay.qpitch();
            if (tx == 0 && x0 > 0) {
                static_for<Q>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    if constexpr (ex(q) == 1) {
                        const long long sid = id0 - 1 - (long long)ey(q) * dim - (long long)ez(q) * plane;
                        f[q] = a.src[a.lay.base(sid) + q * qp];
                    }
                });
            }
            if (tx == TX - 1 && x0 + TX < dim) {
                static_for<Q>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    if constexpr (ex(q) == -1) {
                        const long long sid = id0 + 1 - (long long)ey(q) * dim - (long long)ez(q) * plane;
                        f[q] = a.src[a.lay.base(sid) + q * qp];
                    }
                });
            }
        }
        const int rowbits = row_bits(y, z, dim);
        {
            const int lid = (rowbits & CT_FRONT) ? 1 : 0;
            if (x == 1) {
                static_for<Q>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    if constexpr (ex(q) == 1) f[q] = a.stale[lid][q];
                });
            }
            if (x == dim - 2) {
                static_for<Q>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    if constexpr (ex(q) == -1) f[q] = a.stale[lid][q];
                });
            }
        }

        const int t = cell_type_from_row(rowbits, x, dim);
        T rho = nan, ux = nan, uy = nan, uz = nan;
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

#if LBM_TMA_DIRECT_STORE
        {
            char *const d0 = reinterpret_cast<char *>(a.dst + a.lay.base(id0));
            const long long qs = a.lay.qpitch() * (long long)sizeof(T);
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                *reinterpret_cast<T *>(d0 + q * qs) = f[q];
            });
        }
#else
        // the bulk store that last read this output buffer (tile i - 2) must have finished reading it
        T *const ob = my_out + (i & 1) * (Q * 32);
        if (lane == 0) tma_wait_read<1>();
        __syncwarp();
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            ob[so + q * sq] = f[q];
        });
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk-copy (async) proxy
        __syncwarp();
        if (lane == 0) {
            const int xw = x0 + (warp << 5);  // the warp's first cell
            const long long row = (long long)y * dim + (long long)(z - a.zs0) * plane;
            tma_store_3d(&map_dst, ob, xw & (int)a.lay.smod, 0, (int)((row + xw) >> a.lay.sdiv));
            tma_commit();
        }
#endif

        if constexpr (MACRO) {
            if (!(rowbits & (CT_TOP | CT_BOTTOM | CT_BACK))) {
                a.rho[id0] = rho;
                a.u[id0] = ux;
                a.u[a.n_local + id0] = uy;
                a.u[2 * a.n_local + id0] = uz;
            }
        }
    }
#if !LBM_TMA_DIRECT_STORE
    if (lane == 0) tma_wait_all();
#endif
}

}  // namespace lbm
