// launch_linear.cu — multilinear launchers (regular + rectilinear, f32/f64, N = 1..8).
#include "sweep.cuh"

namespace ib200 {

// The ordinary (direct) kernels. `remap` is set by the bin-swept path, which runs them on sorted coordinates.
template <class T, int N, bool RECT, int WL>
cudaError_t launch_linear_direct(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                                 unsigned long long index_base, cudaStream_t stream, const unsigned* remap,
                                 unsigned long long* work) {
    constexpr int P = linear_points_per_thread<N>();
    auto opts = [&](int ppt) {
        LaunchOpts o;
        o.points_per_thread = ppt;
        o.remap = remap;
        o.work = work;
        o.window = WL != 0;
        // bin-swept evaluation: fewer resident CTAs = a narrower window of keys in flight (tuning hook)
        if (work) o.ctas_per_sm = static_cast<int>(sweep_env("INTERPN_B200_SWEEP_EVAL_CTAS", 8));
        return o;
    };
    if (index64(g)) {  // 64-bit index arithmetic: the basic kernel only, straight from `vals`
        LaunchOpts o = opts(1);
        o.window = false;
        return launch_generic<T, N>(linear_kernel<T, N, RECT, 0, 1, long long>, g, obs, n, out, first_bad, index_base, stream, o);
    }
    if (P > 1 && n >= static_cast<size_t>(P) && vector_aligned<T>(obs, N, out, P))
        return launch_generic<T, N>(linear_kernel<T, N, RECT, WL, P, int>, g, obs, n, out, first_bad, index_base, stream, opts(P));
    return launch_generic<T, N>(linear_kernel<T, N, RECT, WL, 1, int>, g, obs, n, out, first_bad, index_base, stream, opts(1));
}

template <class T, int N, bool RECT>
cudaError_t launch_linear_n(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                            unsigned long long index_base, cudaStream_t stream) {
    constexpr bool kCanWin = N <= kMaxWindowDimsLinear;
    // Window copy (capi.cu window_width): row pairs (N = 1, and small grids whose 2-fold copy still fits L1), else 2x2
    // patches of the last two dimensions.
    constexpr int kPatch = N >= 2 ? 4 : 2;
    if constexpr (N >= 4 && N <= 6) {  // grids beyond L2: 2^(N-4) aligned blocks per point from the hypercube layout (capi.cu window_width)
        if (g.win != nullptr && g.win_width == 16 && !index64(g)) {
            LaunchOpts o;
            o.window = true;
            o.extra_smem = linear_hyper_smem_bytes<T>();
            o.ctas_per_sm = static_cast<int>(sweep_env("INTERPN_B200_HYPER_CTAS", 8));
            return launch_generic<T, N>(linear_hyper_kernel<T, N, RECT>, g, obs, n, out, first_bad, index_base, stream, o);
        }
    }
    if constexpr (N == 3) {  // 3-D grids beyond L2: one aligned 8-value block per point, read by a lane pair
        if (g.win != nullptr && g.win_width == 8 && !index64(g)) {
            LaunchOpts o;
            o.window = true;
            o.extra_smem = linear_hyper_smem_bytes<T>();
            o.ctas_per_sm = static_cast<int>(sweep_env("INTERPN_B200_HYPER_CTAS", 8));
            return launch_generic<T, 3>(linear_hyper3_kernel<T, RECT>, g, obs, n, out, first_bad, index_base, stream, o);
        }
    }
    const bool has_patch = kCanWin && g.win != nullptr && g.win_width == kPatch;
    const bool has_rows = N >= 2 && N <= 4 && g.win != nullptr && g.win_width == 2;
    if constexpr (N >= 2) {  // grids beyond L2: bin-swept evaluation (sweep.cuh), from the patch layout when there is one
        bool swept = false;
        cudaError_t e = launch_sweep<T, N, RECT>(
            g, 2, has_patch ? kPatch : 1, has_patch ? 1 << (N - 2) : 1 << (N - 1), 2, obs, n, out, first_bad, index_base, stream,
            [&](const T* const* sobs, size_t cnt, T* res, const unsigned* orig, unsigned long long base, unsigned long long* work) {
                if constexpr (kCanWin) {
                    if (has_patch) return launch_linear_direct<T, N, RECT, kPatch>(g, sobs, cnt, res, first_bad, base, stream, orig, work);
                }
                return launch_linear_direct<T, N, RECT, 0>(g, sobs, cnt, res, first_bad, base, stream, orig, work);
            },
            swept);
        if (e != cudaSuccess || swept) return e;
    }
    // Direct kernels gather from a window copy only while it is L2-resident.
    if constexpr (kCanWin) {
        if (has_patch && g.nvals * sizeof(T) * kPatch <= kWindowL2Bytes)
            return launch_linear_direct<T, N, RECT, kPatch>(g, obs, n, out, first_bad, index_base, stream, nullptr, nullptr);
    }
    if constexpr (N >= 2 && N <= 4) {
        if (has_rows) return launch_linear_direct<T, N, RECT, 2>(g, obs, n, out, first_bad, index_base, stream, nullptr, nullptr);
    }
    // Grids a little beyond L2, footprints too small for the bin-swept path to pay (C3-linear): slab passes
    // (kernels.cuh linear_slab_kernel), each over the cells of dimension 0 whose rows of `vals` make one L2-resident slab.
    if constexpr (N >= 3 && N <= 5) {
        const size_t pass_kb = sweep_env("INTERPN_B200_SLAB_PASS_KB", 45 << 10);  // 0 = off
        const size_t min_kb = sweep_env("INTERPN_B200_SLAB_MIN_KB", 80 << 10);
        const size_t min_points = sweep_env("INTERPN_B200_SLAB_MIN_POINTS", 1 << 20);
        const size_t ctas = sweep_env("INTERPN_B200_SLAB_CTAS", 8);
        const size_t bytes = g.nvals * sizeof(T);
        if (pass_kb && bytes > (min_kb << 10) && !index64(g) && n >= min_points) {
            const int passes = static_cast<int>((bytes + (pass_kb << 10) - 1) / (pass_kb << 10));
            const int cells = g.dim[0] - 1;
            if (passes >= 2 && passes <= 8 && cells >= passes) {
                for (int p = 0; p < passes; ++p) {
                    LaunchOpts o;
                    o.points_per_thread = kSlabTile / 32;
                    o.ctas_per_sm = static_cast<int>(ctas);
                    o.cell_tables = IB200_SLAB_CELL != 0;
                    o.slab_lo = static_cast<int>(static_cast<long long>(cells) * p / passes);
                    o.slab_hi = p + 1 == passes ? -1 : static_cast<int>(static_cast<long long>(cells) * (p + 1) / passes);
                    cudaError_t e = launch_generic<T, N>(linear_slab_kernel<T, N, RECT, int>, g, obs, n, out, first_bad, index_base, stream, o);
                    if (e != cudaSuccess) return e;
                }
                return cudaSuccess;
            }
        }
    }
    return launch_linear_direct<T, N, RECT, 0>(g, obs, n, out, first_bad, index_base, stream, nullptr, nullptr);
}

template <class T>
cudaError_t launch_linear(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                          unsigned long long index_base, cudaStream_t stream) {
    cudaError_t err = cudaErrorInvalidValue;
    if (g.rect) {
        IB200_SWITCH_N(8, err = (launch_linear_n<T, N, true>(g, obs, n, out, first_bad, index_base, stream));)
    } else {
        IB200_SWITCH_N(8, err = (launch_linear_n<T, N, false>(g, obs, n, out, first_bad, index_base, stream));)
    }
    return err;
}

template cudaError_t launch_linear<double>(const DeviceGrid&, const double* const*, size_t, double*,
                                           unsigned long long*, unsigned long long, cudaStream_t);
template cudaError_t launch_linear<float>(const DeviceGrid&, const float* const*, size_t, float*, unsigned long long*,
                                          unsigned long long, cudaStream_t);

}  // namespace ib200
