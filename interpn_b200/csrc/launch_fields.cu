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

// Both kernels follow their single-field twins (kernels.cuh linear_kernel / nearest_kernel): a thread owns P consecutive
// points — one vector load per coordinate array, P independent locate chains — and then walks the fields, P gathers /
// lerp trees in flight per field and one vector store per field. The n % P tail is evaluated one point per thread.
template <class T, int N, bool RECT, int WL, int P>
__global__ void __launch_bounds__(kBlock) linear_fields_kernel(const __grid_constant__ EvalArgs<T, N> a,
                                                               const __grid_constant__ FieldArgs<T> f) {
    const int(&stride)[N] = a.istride;
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long gtid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long ngroups = a.n / P;
    for (unsigned long long g = gtid; g < ngroups; g += gstride) {
        const unsigned long long i0 = g * P;
        T xs[P][N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T v[P];
            load_query_vec<T, P>(a.obs[d] + i0, v);
#pragma unroll
            for (int p = 0; p < P; ++p) xs[p][d] = v[p];
        }
        T t[P][N];
        int base[P];
        bool ok[P];
        bool all_ok = true;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ok[p] = linear_locate_any<T, N, RECT, int, (WL != 0)>(a, axes, xs[p], t[p], base[p]);
            all_ok = all_ok && ok[p];
            if (!ok[p]) base[p] = 0;  // keep the gathers in range; the values are discarded
        }
#pragma unroll 1
        for (int k = 0; k < f.nf; ++k) {
            T res[P];
#pragma unroll
            for (int p = 0; p < P; ++p) res[p] = linear_tree<T, N, WL, int>(f.vals[k], f.win[k], base[p], stride, t[p]);
            if (all_ok) {
                store_result_vec<T, P>(f.out[k] + i0, res);
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (ok[p]) store_result(f.out[k] + i0 + p, res[p]);
            }
        }
        if (!all_ok) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (!ok[p]) report_bad(a, i0 + p);
        }
    }
    const unsigned long long i = ngroups * P + gtid;
    if (P > 1 && i < a.n) {
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
        T t[N];
        int base;
        if (linear_locate_any<T, N, RECT, int, (WL != 0)>(a, axes, xs, t, base)) {
            for (int k = 0; k < f.nf; ++k) store_result(f.out[k] + i, linear_tree<T, N, WL, int>(f.vals[k], f.win[k], base, stride, t));
        } else {
            report_bad(a, i);
        }
    }
}

template <class T, int N, bool RECT, int P>
__global__ void __launch_bounds__(kBlock) nearest_fields_kernel(const __grid_constant__ EvalArgs<T, N> a,
                                                                const __grid_constant__ FieldArgs<T> f) {
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long gtid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long ngroups = a.n / P;
    for (unsigned long long g = gtid; g < ngroups; g += gstride) {
        const unsigned long long i0 = g * P;
        T xs[P][N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T v[P];
            load_query_vec<T, P>(a.obs[d] + i0, v);
#pragma unroll
            for (int p = 0; p < P; ++p) xs[p][d] = v[p];
        }
        int idx[P];
        bool ok[P];
        bool all_ok = true;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ok[p] = nearest_locate_any<T, N, RECT, int>(a, axes, xs[p], idx[p]);
            all_ok = all_ok && ok[p];
            if (!ok[p]) idx[p] = 0;
        }
#pragma unroll 2
        for (int k = 0; k < f.nf; ++k) {
            T res[P];
#pragma unroll
            for (int p = 0; p < P; ++p) res[p] = __ldg(f.vals[k] + idx[p]);
            if (all_ok) {
                store_result_vec<T, P>(f.out[k] + i0, res);
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (ok[p]) store_result(f.out[k] + i0 + p, res[p]);
            }
        }
        if (!all_ok) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (!ok[p]) report_bad(a, i0 + p);
        }
    }
    const unsigned long long i = ngroups * P + gtid;
    if (P > 1 && i < a.n) {
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
        int idx;
        if (nearest_locate_any<T, N, RECT, int>(a, axes, xs, idx)) {
            for (int k = 0; k < f.nf; ++k) store_result(f.out[k] + i, __ldg(f.vals[k] + idx));
        } else {
            report_bad(a, i);
        }
    }
}

template <class T, int N, class K>
static cudaError_t launch_fields_kernel(K kernel, int P, const DeviceGrid& g, const T* const* obs, size_t n, const FieldArgs<T>& f,
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
    kernel<<<grid_for((n + P - 1) / P, g.sm_count, 8), kBlock, smem, stream>>>(a, f);
    count_launch();
    return cudaGetLastError();
}

// Points per thread as in the single-field launchers; 1 when a coordinate array or an output is not vector-aligned.
template <class T>
static bool fields_aligned(const T* const* obs, int ndims, const FieldArgs<T>& f, int P) {
    const uintptr_t mask = static_cast<uintptr_t>(P) * sizeof(T) - 1;
    uintptr_t bits = 0;
    for (int d = 0; d < ndims; ++d) bits |= reinterpret_cast<uintptr_t>(obs[d]);
    for (int k = 0; k < f.nf; ++k) bits |= reinterpret_cast<uintptr_t>(f.out[k]);
    return (bits & mask) == 0;
}

template <class T, int N, bool RECT>
static cudaError_t linear_fields_n(const DeviceGrid& g, const T* const* obs, size_t n, const FieldArgs<T>& f,
                                   unsigned long long* first_bad, cudaStream_t stream, bool window) {
    constexpr int WLW = N >= 2 ? 4 : 2;
    constexpr int P = linear_points_per_thread<N>();
    const bool vec = P > 1 && n >= static_cast<size_t>(P) && fields_aligned<T>(obs, N, f, P);
    if (window) {
        if (vec) return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, WLW, P>, P, g, obs, n, f, first_bad, stream, true);
        return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, WLW, 1>, 1, g, obs, n, f, first_bad, stream, true);
    }
    if (vec) return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, 0, P>, P, g, obs, n, f, first_bad, stream, false);
    return launch_fields_kernel<T, N>(linear_fields_kernel<T, N, RECT, 0, 1>, 1, g, obs, n, f, first_bad, stream, false);
}

template <class T, int N, bool RECT>
static cudaError_t nearest_fields_n(const DeviceGrid& g, const T* const* obs, size_t n, const FieldArgs<T>& f,
                                    unsigned long long* first_bad, cudaStream_t stream) {
    constexpr int P = (RECT && sizeof(T) == 8) ? IB200_P_NEAREST_RECT : IB200_P_NEAREST;
    const bool vec = n >= static_cast<size_t>(P) && fields_aligned<T>(obs, N, f, P);
    if (vec) return launch_fields_kernel<T, N>(nearest_fields_kernel<T, N, RECT, P>, P, g, obs, n, f, first_bad, stream, true);
    return launch_fields_kernel<T, N>(nearest_fields_kernel<T, N, RECT, 1>, 1, g, obs, n, f, first_bad, stream, true);
}

// grids[k] all describe the same grid (checked by the caller); nf <= kMaxFields. Returns cudaErrorNotSupported when
// the combination has no fused kernel (the caller then evaluates field by field).
template <class T>
cudaError_t launch_eval_fields(const DeviceGrid* const* grids, int nf, const T* const* obs, size_t n, T* const* outs,
                               unsigned long long* first_bad, cudaStream_t stream) {
    const DeviceGrid& g = *grids[0];
    if (n == 0) return cudaSuccess;
    if (g.method == 1 || g.ndims > 6 || index64(g) || nf > kMaxFields) return cudaErrorNotSupported;
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
    // Fusing multiplies the gathered working set by the number of fields: it pays while all of them sit in L2 together
    // (measured, six fields: 3-D nearest 128^3 — 6 x 16 MB — 126 fused vs 160 G field-points/s field by field, the
    // patch copies of 100^3 — 6 x 32 MB — 69 vs 101; a rectilinear search shared six ways still wins, 127 vs 98).
    const size_t gathered = g.nvals * sizeof(T) * static_cast<size_t>(window ? patch : 1);
    const size_t budget = (g.method == 2 && g.rect) ? (size_t(128) << 20) : (size_t(40) << 20);
    if (gathered * static_cast<size_t>(nf) > budget) return cudaErrorNotSupported;
    cudaError_t err = cudaErrorNotSupported;
    if (g.method == 0) {
        if (g.rect) {
            IB200_SWITCH_N(6, err = (linear_fields_n<T, N, true>(g, obs, n, f, first_bad, stream, window));)
        } else {
            IB200_SWITCH_N(6, err = (linear_fields_n<T, N, false>(g, obs, n, f, first_bad, stream, window));)
        }
    } else {
        if (g.rect) {
            IB200_SWITCH_N(6, err = (nearest_fields_n<T, N, true>(g, obs, n, f, first_bad, stream));)
        } else {
            IB200_SWITCH_N(6, err = (nearest_fields_n<T, N, false>(g, obs, n, f, first_bad, stream));)
        }
    }
    return err;
}

template cudaError_t launch_eval_fields<double>(const DeviceGrid* const*, int, const double* const*, size_t, double* const*,
                                                unsigned long long*, cudaStream_t);
template cudaError_t launch_eval_fields<float>(const DeviceGrid* const*, int, const float* const*, size_t, float* const*,
                                               unsigned long long*, cudaStream_t);

}  // namespace ib200
