// launch_common.cuh — shared launcher plumbing for the per-method translation units.
#pragma once
#include "interp_internal.h"
#include "kernels.cuh"

namespace ib200 {

// Rectilinear axes are staged in shared memory when they fit this budget (two CTAs per SM stay
// resident); larger axes are searched in global memory through L1/L2.
constexpr int kAxesSmemBudget = 96 * 1024;

size_t sweep_env_common(const char* name, size_t fallback);

// True when the kernels stage the rectilinear axes blob in shared memory (INTERPN_B200_AXES_SMEM_KB overrides the budget).
template <class T>
inline bool axes_fit_smem(const DeviceGrid& g) {
    const size_t axes_budget = sweep_env_common("INTERPN_B200_AXES_SMEM_KB", kAxesSmemBudget >> 10) << 10;  // read per call (tests)
    return g.rect && static_cast<size_t>(g.axes_total) * sizeof(T) <= axes_budget;
}

template <class T, int N>
inline EvalArgs<T, N> make_args(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                unsigned long long* first_bad, unsigned long long index_base,
                                const unsigned* remap = nullptr, unsigned long long* work = nullptr) {
    EvalArgs<T, N> a{};
    a.remap = remap;
    a.work = work;
    a.slab_lo = 0;
    a.slab_hi = -1;
    for (int d = 0; d < N; ++d) {
        a.obs[d] = obs[d];
        a.stride[d] = g.stride[d];
        a.istride[d] = g.nvals < (size_t(1) << 31) ? static_cast<int>(g.stride[d]) : 0;  // unused by the 64-bit kernels
        a.dim[d] = g.dim[d];
        a.start[d] = static_cast<T>(g.start[d]);
        a.step[d] = static_cast<T>(g.step[d]);
        a.rstep[d] = T(1) / a.step[d];  // correctly rounded in T (host IEEE division)
        a.axis_off[d] = g.axis_off[d];
        a.rc_off[d] = g.rc_off[d];
        a.lut_off[d] = g.lut_off[d];
        a.lut_nb[d] = g.lut_nb[d];
        a.ct_off[d] = g.ct_off[d];
        a.clut_off[d] = g.clut_off[d];
        a.clut_nb[d] = g.clut_nb[d];
        a.clut_scale[d] = static_cast<T>(g.clut_scale[d]);
        a.lut_scale[d] = static_cast<T>(g.lut_scale[d]);
    }
    a.out = out;
    a.n = n;
    a.vals = static_cast<const T*>(g.vals);
    a.win = static_cast<const T*>(g.win);
    a.fast_div = 0;
    if (!g.rect) {
        // guards of the division-free sequences (device_math.cuh): f64 2^-300 <= step < 2^301 and dim < 2^30;
        // f32 2^-60 <= step < 2^60 and dim <= 4096 (the remainder proof needs |cell| * 2^-24 < 2^-12)
        a.fast_div = 1;
        const double lo = sizeof(T) == 8 ? 0x1p-300 : 0x1p-60, hi = sizeof(T) == 8 ? 0x1p301 : 0x1p60;
        const int max_dim = sizeof(T) == 8 ? (1 << 30) - 1 : 4096;
        for (int d = 0; d < N; ++d)
            if (!(g.step[d] >= lo && g.step[d] < hi) || g.dim[d] > max_dim) a.fast_div = 0;
    }
    for (int d = 0; d < N; ++d) {  // exact: power-of-two scalings of a step in the guarded range (lim: rounded, with slack)
        a.hstep[d] = g.step[d] * 0.5;
        a.tau[d] = g.step[d] * (sizeof(T) == 8 ? 0x1p-54 : 0x1p-25);
        a.lim[d] = g.step[d] * (sizeof(T) == 8 ? 1.0 - 0x1p-20 : 1.0 - 0x1p-11);
    }
    a.axes = static_cast<const T*>(g.axes);
    a.axes_total = g.axes_total;
    a.rect_fast = g.rect_fast;
    a.rect_fast_div = g.rect_fast_div;
    a.rect_cubic_table = g.rect_cubic_table;
    a.rect_cell = g.rect_cell;
    a.axes_in_smem = axes_fit_smem<T>(g);
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

// True when every coordinate array and `out` can be accessed as vectors of P elements.
template <class T>
inline bool vector_aligned(const T* const* obs, int ndims, const T* out, int P) {
    const uintptr_t mask = static_cast<uintptr_t>(P) * sizeof(T) - 1;
    uintptr_t bits = reinterpret_cast<uintptr_t>(out);
    for (int d = 0; d < ndims; ++d) bits |= reinterpret_cast<uintptr_t>(obs[d]);
    return (bits & mask) == 0;
}

// Points per thread of the streaming kernels (kernels.cuh linear_kernel / nearest_kernel).
#ifndef IB200_P_NEAREST
#define IB200_P_NEAREST 4
#endif
#ifndef IB200_P_NEAREST_RECT
#define IB200_P_NEAREST_RECT 2
#endif
#ifndef IB200_P_NEAREST_RECT_F32
#define IB200_P_NEAREST_RECT_F32 4
#endif
#ifndef IB200_P_LINEAR_LO
#define IB200_P_LINEAR_LO 4  // N <= 3
#endif
#ifndef IB200_P_LINEAR_HI
#define IB200_P_LINEAR_HI 2  // N = 4, 5
#endif
template <int N>
constexpr int linear_points_per_thread() {
    return N <= 3 ? IB200_P_LINEAR_LO : (N <= 5 ? IB200_P_LINEAR_HI : 1);
}

struct LaunchOpts {
    int slab_lo = 0, slab_hi = -1;       // slab passes (kernels.cuh linear_slab_kernel)
    int points_per_thread = 1;
    const unsigned* remap = nullptr;     // bin-swept path: original indices of the sorted points
    unsigned long long* work = nullptr;  // bin-swept path: dynamic block scheduling counter
    int threads_per_point = 1;           // 4 for the quad-cooperative kernels
    bool window = false;                 // the kernel gathers from the window copy, not from vals
    bool cell_tables = false;            // the kernel locates rectilinear cells through the cell tables although it gathers from vals
    int ctas_per_sm = 8;
    size_t extra_smem = 0;               // dynamic shared memory beyond the staged axes (cubic_quad4.cuh)
};

template <class T, int N, class K>
inline cudaError_t launch_generic(K kernel, const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                                  unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream,
                                  const LaunchOpts& o = LaunchOpts()) {
    if (n == 0) return cudaSuccess;
    const int points_per_thread = o.points_per_thread, threads_per_point = o.threads_per_point, ctas_per_sm = o.ctas_per_sm;
    EvalArgs<T, N> a = make_args<T, N>(g, obs, n, out, first_bad, index_base, o.remap, o.work);
    a.slab_lo = o.slab_lo;
    a.slab_hi = o.slab_hi;
    if (g.rect_cell && g.method == 0 && !o.window && !o.cell_tables && g.axes_core > 0) {
        // multilinear straight from `vals` (grid beyond L2, DRAM-bound): bucket search, cell tables neither read nor
        // staged — the shared memory they would take comes out of L1 (C3-linear 14.2 vs 13.5 G points/s)
        a.rect_cell = 0;
        a.axes_total = g.axes_core;
    }
    size_t smem = a.axes_in_smem ? static_cast<size_t>(a.axes_total) * sizeof(T) : 0;
    if (o.extra_smem) smem = (smem + 15) / 16 * 16 + o.extra_smem;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const size_t items = (n + points_per_thread - 1) / points_per_thread * threads_per_point;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid_for(items, g.sm_count, ctas_per_sm));
    cfg.blockDim = dim3(kBlock);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, a);
    count_launch();
    return le != cudaSuccess ? le : cudaGetLastError();
}

// Window-layout kernels are instantiated up to these dimensionalities (interp_internal.h mirrors them
// in window_policy): the cubic footprint is unrolled row by row only in the flattened range N <= 4.
// A window copy no larger than this is treated as L2-resident by the direct kernels.
constexpr size_t kWindowL2Bytes = size_t(64) << 20;
constexpr int kMaxWindowDimsLinear = 6;
constexpr int kMaxWindowDimsCubic = 4;

// Register budget of the unrolled cubic kernels, as resident CTAs (of kBlock threads) per SM. The
// footprint of a 3-D / 4-D cubic is 64 / 256 gathers deep, so these kernels live on latency hiding:
// the values below are the measured optimum between spilling and occupancy.
#ifndef IB200_MINB_CUBIC3
#define IB200_MINB_CUBIC3 2
#endif
#ifndef IB200_MINB_CUBIC4
#define IB200_MINB_CUBIC4 3
#endif
#ifndef IB200_MINB_CUBIC3_RECT
#define IB200_MINB_CUBIC3_RECT 2
#endif
#ifndef IB200_MINB_CUBIC4_RECT
#define IB200_MINB_CUBIC4_RECT 3
#endif
template <int N, bool RECT>
constexpr int cubic_min_blocks() {
    if (N == 3) return RECT ? IB200_MINB_CUBIC3_RECT : IB200_MINB_CUBIC3;
    if (N == 4) return RECT ? IB200_MINB_CUBIC4_RECT : IB200_MINB_CUBIC4;
    return 1;
}

#define IB200_SWITCH_N(NMAX, BODY)                     \
    switch (g.ndims) {                                 \
        case 1: { constexpr int N = 1; BODY } break;   \
        case 2: { constexpr int N = 2; BODY } break;   \
        case 3: { constexpr int N = 3; BODY } break;   \
        case 4: { constexpr int N = 4; BODY } break;   \
        case 5: if constexpr (NMAX >= 5) { constexpr int N = 5; BODY } break; \
        case 6: if constexpr (NMAX >= 6) { constexpr int N = 6; BODY } break; \
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
