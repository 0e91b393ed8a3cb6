// launch_linear.cu — multilinear launchers (regular + rectilinear, f32/f64, N = 1..8).
#include "launch_common.cuh"

namespace ib200 {

template <class T, int N, bool RECT, bool WIN>
cudaError_t launch_linear_n(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                            unsigned long long index_base, cudaStream_t stream) {
    constexpr int P = linear_points_per_thread<N>();
    if (g.nvals >= (size_t(1) << 31))  // 64-bit index arithmetic: the basic kernel only
        return launch_generic<T, N>(linear_kernel<T, N, RECT, false, 1, long long>, g, obs, n, out, first_bad, index_base, stream);
    if (P > 1 && n >= static_cast<size_t>(P) && vector_aligned<T>(obs, N, out, P))
        return launch_generic<T, N>(linear_kernel<T, N, RECT, WIN, P, int>, g, obs, n, out, first_bad, index_base, stream, P);
    return launch_generic<T, N>(linear_kernel<T, N, RECT, WIN, 1, int>, g, obs, n, out, first_bad, index_base, stream);
}

template <class T>
cudaError_t launch_linear(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                          unsigned long long index_base, cudaStream_t stream) {
    cudaError_t err = cudaErrorInvalidValue;
    const bool win = g.win != nullptr && g.win_width == 2 && g.ndims <= kMaxWindowDimsLinear;
    if (g.rect) {
        if (win) {
            IB200_SWITCH_N(kMaxWindowDimsLinear, err = (launch_linear_n<T, N, true, true>(g, obs, n, out, first_bad, index_base, stream));)
        } else {
            IB200_SWITCH_N(8, err = (launch_linear_n<T, N, true, false>(g, obs, n, out, first_bad, index_base, stream));)
        }
    } else {
        if (win) {
            IB200_SWITCH_N(kMaxWindowDimsLinear, err = (launch_linear_n<T, N, false, true>(g, obs, n, out, first_bad, index_base, stream));)
        } else {
            IB200_SWITCH_N(8, err = (launch_linear_n<T, N, false, false>(g, obs, n, out, first_bad, index_base, stream));)
        }
    }
    return err;
}

template cudaError_t launch_linear<double>(const DeviceGrid&, const double* const*, size_t, double*,
                                           unsigned long long*, unsigned long long, cudaStream_t);
template cudaError_t launch_linear<float>(const DeviceGrid&, const float* const*, size_t, float*, unsigned long long*,
                                          unsigned long long, cudaStream_t);

}  // namespace ib200
