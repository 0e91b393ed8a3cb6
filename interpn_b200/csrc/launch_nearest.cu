// launch_nearest.cu — nearest-neighbour launchers (regular + rectilinear, f32/f64, N = 1..6).
#include "launch_common.cuh"

namespace ib200 {

template <class T>
cudaError_t launch_nearest(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                           unsigned long long index_base, cudaStream_t stream) {
    cudaError_t err = cudaErrorInvalidValue;
    // Points per thread: 4 on regular grids; 2 on rectilinear ones, whose twelve search chains per thread (3-D) cost
    // 78 registers and a quarter of the occupancy (measured: 3-D 78 -> 98 G points/s, 2-D 116 -> 120).
    // (f32 keeps 4: two floats per thread would halve the width of the streaming accesses, and its chains fit 64 registers)
    constexpr int P = IB200_P_NEAREST, PR = sizeof(T) == 4 ? IB200_P_NEAREST_RECT_F32 : IB200_P_NEAREST_RECT;
    LaunchOpts popts;
    popts.points_per_thread = g.rect ? PR : P;
    const int pp = popts.points_per_thread;
    const bool vec = pp > 1 && n >= static_cast<size_t>(pp) && vector_aligned<T>(obs, g.ndims, out, pp);
    if (index64(g)) {  // 64-bit index arithmetic: the basic kernel only
        if (g.rect) {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, true, 1, long long>, g, obs, n, out, first_bad, index_base, stream));)
        } else {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, false, 1, long long>, g, obs, n, out, first_bad, index_base, stream));)
        }
        return err;
    }
    if (g.rect) {
        if (vec) {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, true, PR, int>, g, obs, n, out, first_bad, index_base, stream, popts));)
        } else {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, true, 1, int>, g, obs, n, out, first_bad, index_base, stream));)
        }
    } else {
        if (vec) {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, false, P, int>, g, obs, n, out, first_bad, index_base, stream, popts));)
        } else {
            IB200_SWITCH_N(6, err = (launch_generic<T, N>(nearest_kernel<T, N, false, 1, int>, g, obs, n, out, first_bad, index_base, stream));)
        }
    }
    return err;
}

template cudaError_t launch_nearest<double>(const DeviceGrid&, const double* const*, size_t, double*,
                                            unsigned long long*, unsigned long long, cudaStream_t);
template cudaError_t launch_nearest<float>(const DeviceGrid&, const float* const*, size_t, float*, unsigned long long*,
                                           unsigned long long, cudaStream_t);

}  // namespace ib200
