// launch_common.cuh — shared launcher plumbing for the per-method translation units.
#pragma once
#include "interp_internal.h"
#include "kernels.cuh"

namespace ib200 {

// Rectilinear axes are staged in shared memory when they fit this budget (two CTAs per SM stay
// resident); larger axes are searched in global memory through L1/L2.
constexpr int kAxesSmemBudget = 96 * 1024;

template <class T, int N>
inline EvalArgs<T, N> make_args(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                unsigned long long* first_bad, unsigned long long index_base) {
    EvalArgs<T, N> a{};
    for (int d = 0; d < N; ++d) {
        a.obs[d] = obs[d];
        a.stride[d] = g.stride[d];
        a.dim[d] = g.dim[d];
        a.start[d] = static_cast<T>(g.start[d]);
        a.step[d] = static_cast<T>(g.step[d]);
        a.axis_off[d] = g.axis_off[d];
    }
    a.out = out;
    a.n = n;
    a.vals = static_cast<const T*>(g.vals);
    a.axes = static_cast<const T*>(g.axes);
    a.axes_total = g.axes_total;
    a.axes_in_smem = g.rect && static_cast<size_t>(g.axes_total) * sizeof(T) <= kAxesSmemBudget;
    a.linearize = g.linearize;
    a.first_bad = first_bad;
    a.index_base = index_base;
    return a;
}

// Grid-stride launch: enough CTAs to fill every SM several times over, never more than needed.
inline unsigned grid_for(size_t n, int sm_count, int ctas_per_sm) {
    size_t want = (n + kBlock - 1) / kBlock;
    size_t cap = static_cast<size_t>(sm_count) * ctas_per_sm;
    return static_cast<unsigned>(want < cap ? (want ? want : 1) : cap);
}

template <class T, int N, class K>
inline cudaError_t launch_generic(K kernel, const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                  unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream,
                                  int ctas_per_sm = 8) {
    if (n == 0) return cudaSuccess;
    EvalArgs<T, N> a = make_args<T, N>(g, obs, n, out, first_bad, index_base);
    size_t smem = a.axes_in_smem ? static_cast<size_t>(g.axes_total) * sizeof(T) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid_for(n, g.sm_count, ctas_per_sm), kBlock, smem, stream>>>(a);
    count_launch();
    return cudaGetLastError();
}

#define IB200_SWITCH_N(NMAX, BODY)                     \
    switch (g.ndims) {                                 \
        case 1: { constexpr int N = 1; BODY } break;   \
        case 2: { constexpr int N = 2; BODY } break;   \
        case 3: { constexpr int N = 3; BODY } break;   \
        case 4: { constexpr int N = 4; BODY } break;   \
        case 5: { constexpr int N = 5; BODY } break;   \
        case 6: { constexpr int N = 6; BODY } break;   \
        case 7: if constexpr (NMAX >= 7) { constexpr int N = 7; BODY } break; \
        case 8: if constexpr (NMAX >= 8) { constexpr int N = 8; BODY } break; \
        default: break;                                \
    }

// Per-method launchers, one translation unit each (parallel compilation).
template <class T>
cudaError_t launch_linear(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                          unsigned long long index_base, cudaStream_t stream);
template <class T>
cudaError_t launch_cubic_regular(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                 unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream);
template <class T>
cudaError_t launch_cubic_rect(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                              unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream);
template <class T>
cudaError_t launch_nearest(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                           unsigned long long index_base, cudaStream_t stream);

}  // namespace ib200
