// TMA-fed variant of the D3Q19 step (sm_90+/sm_100a: cp.async.bulk.tensor + mbarrier).
//
// Same function as step_pull_kernel (two-lattice pull, post-collision storage, see lbm_kernels.cuh);
// what changes is how the bytes move:
//
//   * the lattice is described to the TMA unit as a 3-D tensor  [block][q][stride]  (the CSoA layout of
//     kernels.cl:64 read literally), so ONE bulk-tensor copy fetches the TX populations of one direction
//     of one x-row segment -- whatever the stride -- into a contiguous shared-memory row;
//   * persistent CTAs (a few per SM) walk the live rows of the launch; one elected thread keeps a ring of
//     NS row-tiles in flight (19 bulk loads per tile, completion on an mbarrier with expect_tx);
//   * every thread owns one cell: 19 conflict-free LDS at compile-time offsets (the x +- 1 shift is
//     just an index), the reference's BC / collision, 19 STS into the same tile, and the elected
//     thread sends the tile back with 19 bulk-tensor stores.
//
// Per cell this replaces 19 LDG + 19 STG + ~76 64-bit address instructions by 19 LDS + 19 STS with
// immediate offsets, and decouples the HBM latency from the occupancy (the ring depth hides it).
// Requirements (checked by the host): LM_ROWS layout (stride <= DIM), stride*sizeof(T) >= 16 bytes,
// DIM >= 32.  Rows wider than TX = 256 cells are cut into segments; the two segment-edge threads fetch
// their out-of-tile neighbour with a plain load.
#pragma once

#include <cuda.h>

#include "lbm_kernels.cuh"

namespace lbm {

template <typename T>
struct TmaArgs {
    const T *__restrict__ src;  // for the segment-edge loads only
    T *__restrict__ rho;
    T *__restrict__ u;
    int dim;
    int zs0;
    int z_first;        // first live global plane of this launch (already clipped to [1, DIM-2])
    int n_tiles;        // live rows x segments of this launch
    int n_xseg;         // DIM / TX
    int ns;             // ring depth
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

// TX = cells per row tile = threads per CTA (32, 64, 128 or 256) is the launch's blockDim.x, not a template
// parameter: the kernel exists once per (T, FAST, MACRO) instead of four times.
template <typename T, bool FAST, bool MACRO>
__global__ void __launch_bounds__(256) step_tma_kernel(const __grid_constant__ CUtensorMap map_src,
                                                        const __grid_constant__ CUtensorMap map_dst,
                                                        const TmaArgs<T> a, int *__restrict__ error_flag)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TX = (int)blockDim.x;
    const uint32_t STAGE_BYTES = (uint32_t)(Q * TX * sizeof(T));
    const int NS = a.ns;
    T *const ring = reinterpret_cast<T *>(smem_raw);                                   // [NS][Q][TX]
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NS * STAGE_BYTES);  // [NS]

    const int tx = threadIdx.x;
    const int dim = a.dim;
    const int rows = dim - 2;  // live y per plane
    const long long plane = (long long)dim * dim;

    if (tx == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_my = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto tile_coords = [&](int i, int &x0, int &y, int &z) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        const int xs = t % a.n_xseg;
        const int r = t / a.n_xseg;
        x0 = xs * TX;
        y = 1 + r % rows;
        z = a.z_first + r / rows;
    };
    // tensor coordinates of the TX values of direction q in row (yy, zz) starting at x0:
    //   c0 = offset inside the CSoA run, c1 = q, c2 = CSoA block index
    auto issue_loads = [&](int i) {
        int x0, y, z;
        tile_coords(i, x0, y, z);
        const int s = i % NS;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        T *const dst = ring + (size_t)s * Q * TX;
        const int c0 = x0 & (int)a.lay.smod;
        const int xb = x0 >> a.lay.sdiv;
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            const long long row = (long long)(y - ey(q)) * dim + (long long)(z - ez(q) - a.zs0) * plane;
            tma_load_3d(dst + q * TX, &map_src, &full[s], c0, q, (int)(row >> a.lay.sdiv) + xb);
        });
    };

    if (tx == 0) {
        for (int i = 0; i < NS && i < n_my; ++i) issue_loads(i);
    }

    const T nan = static_cast<T>(__int_as_float(0x7fc00000));
    for (int i = 0; i < n_my; ++i) {
        int x0, y, z;
        tile_coords(i, x0, y, z);
        const int x = x0 + tx;
        const int s = i % NS;
        const uint32_t parity = (uint32_t)(i / NS) & 1u;
        T *const tile = ring + (size_t)s * Q * TX;

        // bounded wait: a TMA fault must not hang the GPU.  A thread that runs into the limit carries on
        // with whatever the tile holds; the CTA leaves together at the barrier below.
        bool timed_out = false;
        {
            long long spins = 0;
            while (!mbar_try_wait(&full[s], parity)) {
                if (++spins > (1ll << 22)) {
                    timed_out = true;
                    break;
                }
            }
        }

        T f[Q];
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            f[q] = tile[q * TX + tx - ex(q)];  // tx -+ 1 outside the tile: fixed below (or a WALL cell)
        });
        if (a.n_xseg > 1) {
            const long long id0 = x + (long long)y * dim + (long long)(z - a.zs0) * plane;
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
        // every thread has taken its inputs: the tile may be overwritten in place
        if (__syncthreads_or(timed_out ? 1 : 0)) {  // CTA-uniform exit
            if (tx == 0 && error_flag) atomicExch(error_flag, 1);
            return;
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
        static_for<Q>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            tile[q * TX + tx] = f[q];
        });
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk-copy (async) proxy
        __syncthreads();

        if (tx == 0) {
            const int c0 = x0 & (int)a.lay.smod;
            const long long row = (long long)y * dim + (long long)(z - a.zs0) * plane;
            const int c2 = (int)(row >> a.lay.sdiv) + (x0 >> a.lay.sdiv);
            static_for<Q>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                tma_store_3d(&map_dst, tile + q * TX, c0, q, c2);
            });
            tma_commit();
            // the stage of the previous tile is free once its stores have read shared memory
            if (i >= 1 && i - 1 + NS < n_my) {
                tma_wait_read<1>();
                issue_loads(i - 1 + NS);
            }
        }

        if constexpr (MACRO) {
            if (!(rowbits & (CT_TOP | CT_BOTTOM | CT_BACK))) {
                const long long id0 = x + (long long)y * dim + (long long)(z - a.zs0) * plane;
                a.rho[id0] = rho;
                a.u[id0] = ux;
                a.u[a.n_local + id0] = uy;
                a.u[2 * a.n_local + id0] = uz;
            }
        }
    }
    if (tx == 0) tma_wait_all();
}

}  // namespace lbm
