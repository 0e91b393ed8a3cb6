// launch_cubic_build.cu — builds the coefficient layout of a multicubic grid (cubic_quad4.cuh): one sector per
// (slot of dimension 0, node of dimensions 1..N-1), computed with the device code of the 1-D step itself.
#include "launch_common.cuh"
#include "cubic_quad4.cuh"

namespace ib200 {

cudaError_t launch_build_coef_window(const DeviceGrid& g, cudaStream_t stream) {
    if (!g.win || g.nvals == 0) return cudaSuccess;
    const unsigned long long s0 = static_cast<unsigned long long>(g.stride[0]);
    const size_t sectors = static_cast<size_t>(g.dim[0] + 1) * s0;
    const unsigned grid_dim = grid_for(sectors, g.sm_count, 8);
    if (g.elem == 8) {
        const double* vals = static_cast<const double*>(g.vals);
        double* win = static_cast<double*>(g.win);
        if (g.rect) build_coef_window_kernel<double, true><<<grid_dim, kBlock, 0, stream>>>(vals, win, g.dim[0], s0, static_cast<const double*>(g.axes) + g.ct_off[0]);
        else build_coef_window_kernel<double, false><<<grid_dim, kBlock, 0, stream>>>(vals, win, g.dim[0], s0, nullptr);
    } else {
        const float* vals = static_cast<const float*>(g.vals);
        float* win = static_cast<float*>(g.win);
        if (g.rect) build_coef_window_kernel<float, true><<<grid_dim, kBlock, 0, stream>>>(vals, win, g.dim[0], s0, static_cast<const float*>(g.axes) + g.ct_off[0]);
        else build_coef_window_kernel<float, false><<<grid_dim, kBlock, 0, stream>>>(vals, win, g.dim[0], s0, nullptr);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ib200
