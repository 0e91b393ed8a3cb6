// launch_fields.cu — fused evaluation of several fields that share one grid and one query batch:
// the cell location (divisions / axis searches, N coordinate loads) is done once per point, then every field's
// corner gather and lerp tree (multilinear) or single gather (nearest) runs on it. The callers' pattern is the
// reference's own benchmark, six interpolators over one grid (bench_cpu.py:501-510). Per-field arithmetic is the
// device code of the single-field kernels (kernels.cuh linear_tree / *_locate_any), so every field's result is
// bit-identical to its own `.interp` call. Multicubic fields, N > 6 and grids beyond L2 (bin-swept) are evaluated one
// field after the other by the caller (capi.cu eval_fields_device).
#include "launch_common.cuh"

namespace ib200 {

template <class T>
struct FieldArgs {
    const T* vals[kMaxFields];
    const T* win[kMaxFields];
    T* out[kMaxFields];
    int nf;
};

template <class T, int N, bool RECT, int WL>
__global__ void __launch_bounds__(kBlock) linear_fields_kernel(const __grid_constant__ EvalArgs<T, N> a,
                                                               const __grid_constant__ FieldArgs<T> f) {
    const int(&stride)[N] = a.istride;
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride) {
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
        T t[N];
        int base;
        if (linear_locate_any<T, N, RECT, int, (WL != 0)>(a, axes, xs, t, base)) {
#pragma unroll 2
            for (int k = 0; k < f.nf; ++k) store_result(f.out[k] + i, linear_tree<T, N, WL, int>(f.vals[k], f.win[k], base, stride, t));
        } else {
            report_bad(a, i);
        }
    }
}

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) nearest_fields_kernel(const __grid_constant__ EvalArgs<T, N> a,
                                                                const __grid_constant__ FieldArgs<T> f) {
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride) {
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
        int idx;
        if (nearest_locate_any<T, N, RECT, int>(a, axes, xs, idx)) {
#pragma unroll 4
            for (int k = 0; k < f.nf; ++k) store_result(f.out[k] + i, __ldg(f.vals[k] + idx));
        } else {
            report_bad(a, i);
        }
    }
}

template <class T, int N, class K>
static cudaError_t launch_fields_kernel(K kernel, const DeviceGrid& g, const T* const* obs, size_t n, const FieldArgs<T>& f,
                                        unsigned long long* first_bad, cudaStream_t stream, bool window) {
    EvalArgs<T, N> a = make_args<T, N>(g, obs, n, f.out[0], first_bad, 0ull);
    if (g.rect_cell && g.method == 0 && !window && g.axes_core > 0) {  // as launch_generic: no cell tables without a window
        a.rect_cell = 0;
        a.axes_total = g.axes_core;
    }
    const size_t smem = a.axes_in_smem ? static_cast<size_t>(a.axes_total) * sizeof(T) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid_for(n, g.sm_count, 8), kBlock, smem, stream>>>(a, f);
    count_launch();
    return cudaGetLastError();
}

template <class T, int N, bool RECT>
static cudaError_t linear_fields_n(const DeviceGrid& g, const T* const* obs, size_t n, const FieldArgs<T>& f,
                                   unsigned long long* first_bad, cudaStream_t stream, bool window) {
    constexpr int WLW = N >= 2 ? 4 : 2;
    if (window) return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, WLW>, g, obs, n, f, first_bad, stream, true);
    return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, 0>, g, obs, n, f, first_bad, stream, false);
}

// grids[k] all describe the same grid (checked by the caller); nf <= kMaxFields. Returns cudaErrorNotSupported when
// the combination has no fused kernel (the caller then evaluates field by field).
template <class T>
cudaError_t launch_eval_fields(const DeviceGrid* const* grids, int nf, const T* const* obs, size_t n, T* const* outs,
                               unsigned long long* first_bad, cudaStream_t stream) {
    const DeviceGrid& g = *grids[0];
    if (n == 0) return cudaSuccess;
    if (g.method == 1 || g.ndims > 6 || g.nvals >= (size_t(1) << 31) || nf > kMaxFields) return cudaErrorNotSupported;
    FieldArgs<T> f{};
    f.nf = nf;
    const int patch = g.ndims >= 2 ? 4 : 2;
    bool window = g.method == 0;
    for (int k = 0; k < nf; ++k) {
        f.vals[k] = static_cast<const T*>(grids[k]->vals);
        f.win[k] = static_cast<const T*>(grids[k]->win);
        f.out[k] = outs[k];
        // the fused multilinear kernel gathers from the patch (N >= 2) / row-pair (N = 1) copies, L2-resident ones only
        window = window && grids[k]->win != nullptr && grids[k]->win_width == patch &&
                 g.nvals * sizeof(T) * static_cast<size_t>(patch) <= kWindowL2Bytes;
    }
    if (g.method == 0 && !window && g.nvals * sizeof(T) > (size_t(96) << 20)) return cudaErrorNotSupported;  // bin-swept territory
    cudaError_t err = cudaErrorNotSupported;
    if (g.method == 0) {
        if (g.rect) {
            IB200_SWITCH_N(6, err = (linear_fields_n<T, N, true>(g, obs, n, f, first_bad, stream, window));)
        } else {
            IB200_SWITCH_N(6, err = (linear_fields_n<T, N, false>(g, obs, n, f, first_bad, stream, window));)
        }
    } else {
        if (g.rect) {
            IB200_SWITCH_N(6, err = (launch_fields_kernel<T, N>(nearest_fields_kernel<T, N, true>, g, obs, n, f, first_bad, stream, true));)
        } else {
            IB200_SWITCH_N(6, err = (launch_fields_kernel<T, N>(nearest_fields_kernel<T, N, false>, g, obs, n, f, first_bad, stream, true));)
        }
    }
    return err;
}

template cudaError_t launch_eval_fields<double>(const DeviceGrid* const*, int, const double* const*, size_t, double* const*,
                                                unsigned long long*, cudaStream_t);
template cudaError_t launch_eval_fields<float>(const DeviceGrid* const*, int, const float* const*, size_t, float* const*,
                                               unsigned long long*, cudaStream_t);

}  // namespace ib200
