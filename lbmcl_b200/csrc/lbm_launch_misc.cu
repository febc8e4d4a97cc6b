// Kernels that run once or rarely: `initialize`, the ghost-constant table, the cell-type map, the
// reference views for the -f dump, the dense halo pack/unpack and the one-thread helpers of the in-kernel
// slab flags.
#include "lbm_launch.hpp"

namespace lbm {

namespace {

// Cell-type map (kernels.cl:290), only for the -m dump.
__global__ void map_kernel(int *__restrict__ map, int dim)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= dim || y >= dim) return;
    map[x + (long long)y * dim + (long long)z * dim * dim] = cell_type(x, y, z, dim);
}

// PEER_FLAGS transport around `initialize`: wait for both neighbours' previous phase (their last stores
// into my halo planes) / publish a phase that has no boundary kernel.
__global__ void slab_wait_kernel(SlabSync s, int has_lo, int has_hi)
{
    if (has_lo) slab_wait(s.flag_in + 0, s.wait_epoch, s.timeout_ns, s.error);
    if (has_hi) slab_wait(s.flag_in + 1, s.wait_epoch, s.timeout_ns, s.error);
}
__global__ void slab_signal_kernel(SlabSync s, int has_lo, int has_hi)
{
    __threadfence_system();
    if (has_lo) st_release_sys(s.flag_out[0], s.signal_epoch);
    if (has_hi) st_release_sys(s.flag_out[1], s.signal_epoch);
}

// x-major blocks of up to 256 threads covering one plane per grid z
void plane_launch_shape(int dim, int planes, dim3 &g, dim3 &b)
{
    const int bx = dim < 64 ? dim : 64;
    const int by = (256 / bx) < dim ? (256 / bx) : dim;
    b = dim3(bx, by, 1);
    g = dim3(dim / bx, dim / by, planes);
}

template <typename T>
cudaError_t launch_init(const InitArgs<T> &a, bool aa, cudaStream_t s)
{
    dim3 g, b;
    plane_launch_shape(a.dim, a.nz_local, g, b);
    if (aa) init_aa_kernel<T><<<g, b, 0, s>>>(a);
    else init_kernel<T><<<g, b, 0, s>>>(a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_view(const T *g_, T *out, const ViewCfg &v, const Consts<T> &c, bool aa, cudaStream_t s)
{
    dim3 g, b;
    if (aa) {
        plane_launch_shape(v.dim, v.dim, g, b);
        reference_view_aa_kernel<T><<<g, b, 0, s>>>(g_, out, v.dim, v.lay_local, c, v.aa_swapped, v.pristine);
    } else {
        plane_launch_shape(v.dim, v.z_end - v.z_begin, g, b);
        reference_view_kernel<T><<<g, b, 0, s>>>(g_, out, v.dim, v.zs0, v.nz_local, v.z_begin, v.z_end, v.lay_local,
                                                 v.lay_global, c, v.pristine);
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_halo(T *lattice, T *dense, int dim, long long plane_local, Layout lay, int dir_up, bool pack,
                        cudaStream_t s)
{
    const int bx = dim < 256 ? dim : 256;
    const dim3 b(bx, 1, 1), g(dim / bx, dim, 5);
    if (pack) halo_kernel<T, true><<<g, b, 0, s>>>(lattice, dense, dim, plane_local, lay, dir_up);
    else halo_kernel<T, false><<<g, b, 0, s>>>(lattice, dense, dim, plane_local, lay, dir_up);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_init_f32(const InitArgs<float> &a, bool aa, cudaStream_t s) { return launch_init<float>(a, aa, s); }
cudaError_t launch_init_f64(const InitArgs<double> &a, bool aa, cudaStream_t s) { return launch_init<double>(a, aa, s); }

cudaError_t launch_stale_f32(const Consts<float> &c, float *out, cudaStream_t s)
{
    stale_kernel<float><<<1, 1, 0, s>>>(c, out);
    return cudaGetLastError();
}
cudaError_t launch_stale_f64(const Consts<double> &c, double *out, cudaStream_t s)
{
    stale_kernel<double><<<1, 1, 0, s>>>(c, out);
    return cudaGetLastError();
}

cudaError_t launch_map(int *map, int dim, cudaStream_t s)
{
    dim3 g, b;
    plane_launch_shape(dim, dim, g, b);
    map_kernel<<<g, b, 0, s>>>(map, dim);
    return cudaGetLastError();
}

cudaError_t launch_view_f32(const float *g, float *out, const ViewCfg &v, const Consts<float> &c, bool aa, cudaStream_t s)
{
    return launch_view<float>(g, out, v, c, aa, s);
}
cudaError_t launch_view_f64(const double *g, double *out, const ViewCfg &v, const Consts<double> &c, bool aa, cudaStream_t s)
{
    return launch_view<double>(g, out, v, c, aa, s);
}

cudaError_t launch_halo_f32(float *lattice, float *dense, int dim, long long plane_local, Layout lay, int dir_up, bool pack,
                            cudaStream_t s)
{
    return launch_halo<float>(lattice, dense, dim, plane_local, lay, dir_up, pack, s);
}
cudaError_t launch_halo_f64(double *lattice, double *dense, int dim, long long plane_local, Layout lay, int dir_up, bool pack,
                            cudaStream_t s)
{
    return launch_halo<double>(lattice, dense, dim, plane_local, lay, dir_up, pack, s);
}

cudaError_t launch_slab_wait(const SlabSync &y, bool has_lo, bool has_hi, cudaStream_t s)
{
    slab_wait_kernel<<<1, 1, 0, s>>>(y, has_lo ? 1 : 0, has_hi ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_slab_signal(const SlabSync &y, bool has_lo, bool has_hi, cudaStream_t s)
{
    slab_signal_kernel<<<1, 1, 0, s>>>(y, has_lo ? 1 : 0, has_hi ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace lbm
