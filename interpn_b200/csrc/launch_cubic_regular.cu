// launch_cubic_regular.cu — multicubic regular-grid launchers (f32/f64, N = 1..8).
#include "launch_common.cuh"

namespace ib200 {

template <class T>
cudaError_t launch_cubic_regular(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                 unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream) {
    cudaError_t err = cudaErrorInvalidValue;
    if (g.win != nullptr && g.win_width == 4 && g.ndims <= kMaxWindowDimsCubic) {
        IB200_SWITCH_N(kMaxWindowDimsCubic, err = (launch_generic<T, N>(cubic_kernel<T, N, false, true, cubic_min_blocks<N, false>()>, g, obs, n, out, first_bad, index_base, stream));)
    } else {
        IB200_SWITCH_N(8, err = (launch_generic<T, N>(cubic_kernel<T, N, false, false, cubic_min_blocks<N, false>()>, g, obs, n, out, first_bad, index_base, stream));)
    }
    return err;
}

template cudaError_t launch_cubic_regular<double>(const DeviceGrid&, const double* const*, size_t, double*,
                                                  unsigned long long*, unsigned long long, cudaStream_t);
template cudaError_t launch_cubic_regular<float>(const DeviceGrid&, const float* const*, size_t, float*,
                                                 unsigned long long*, unsigned long long, cudaStream_t);

}  // namespace ib200
