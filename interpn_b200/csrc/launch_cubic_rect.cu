// launch_cubic_rect.cu — multicubic rectilinear-grid launchers (f32/f64, N = 1..8).
#include <type_traits>

#include "sweep.cuh"
#include "cubic_quad4.cuh"

namespace ib200 {

template <class T>
cudaError_t launch_cubic_rect(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                              unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream) {
    cudaError_t err = cudaErrorInvalidValue;
    // Window copies (capi.cu window_width): rows of four for N = 1, the coefficient layout for N = 2..4 (cubic_quad4.cuh;
    // rectilinear: only when the axes have a cell table, i.e. are strictly increasing and finite).
    const bool has_win = g.win != nullptr && g.win_width == 4 && g.ndims <= kMaxWindowDimsCubic;
    // `remap` / `work` are set by the bin-swept path, which runs the same kernels on sorted coordinates.
    auto direct = [&](bool win, const T* const* o, size_t cnt, T* dst, unsigned long long base, const unsigned* remap,
                      unsigned long long* work) {
        cudaError_t e = cudaErrorInvalidValue;
        auto lo = [&](int threads_per_point, bool window) {
            LaunchOpts r;
            r.remap = remap;
            r.work = work;
            r.threads_per_point = threads_per_point;
            r.window = window;
            return r;
        };
        if (win && g.ndims == 1) {
            e = launch_generic<T, 1>(cubic_kernel<T, 1, true, true, 1>, g, o, cnt, dst, first_bad, base, stream, lo(1, true));
        } else if (win) {
            // Four points per quad over the coefficient layout (cubic_quad4.cuh); INTERPN_B200_QUAD4_MINB selects the
            // register budget (CTAs per SM) for tuning runs.
            static const int minb = static_cast<int>(sweep_env("INTERPN_B200_QUAD4_MINB", 0));
            auto q4 = [&](auto kernel, auto ntag) {
                constexpr int N = decltype(ntag)::value;
                LaunchOpts r = lo(1, true);
                r.extra_smem = quad4_smem_bytes<T, N, true>();
                return launch_generic<T, N>(kernel, g, o, cnt, dst, first_bad, base, stream, r);
            };
            using std::integral_constant;
            const bool in_smem = axes_fit_smem<T>(g);  // the blob staged in shared memory (LDS) or read through L1
#define IB200_Q4R(NN, MB)                                                                                        \
    e = in_smem ? q4(cubic_quad4_kernel<T, NN, true, MB, true>, integral_constant<int, NN>())                    \
                : q4(cubic_quad4_kernel<T, NN, true, MB, false>, integral_constant<int, NN>())
            switch (g.ndims) {
                case 2: IB200_Q4R(2, 3); break;
                case 3:
                    if (minb == 2) IB200_Q4R(3, 2);
                    else if (minb == 4) IB200_Q4R(3, 4);
                    else IB200_Q4R(3, 3);
                    break;
                case 4:
                    // measured (gpurun_out/r2_exp6, G points/s at 2 / 3 / 4 CTAs per SM): L2-resident 32^4 5.65 / 5.43 / 6.06,
                    // C3-cubic through the bin-swept path 4.66 / 4.70 / 4.54
                    if (minb == 2) IB200_Q4R(4, 2);
                    else if (minb == 3 || (minb == 0 && work != nullptr)) IB200_Q4R(4, 3);
                    else IB200_Q4R(4, 4);
                    break;
                default: break;
            }
#undef IB200_Q4R
        } else {
            IB200_SWITCH_N(8, e = (launch_generic<T, N>(cubic_kernel<T, N, true, false, cubic_min_blocks<N, true>()>, g, o, cnt, dst, first_bad, base, stream, lo(1, false)));)
        }
        return e;
    };
    // Grids beyond L2: bin-swept evaluation (sweep.cuh), gathering from the window layout when there is one.
    if (g.ndims >= 2 && g.ndims <= kMaxWindowDimsCubic) {
        bool swept = false;
        auto eval = [&](const T* const* sobs, size_t cnt, T* res, const unsigned* orig, unsigned long long base,
                        unsigned long long* work) {
            return direct(has_win, sobs, cnt, res, base, orig, work);
        };
        IB200_SWITCH_N(kMaxWindowDimsCubic, if constexpr (N >= 2) err = (launch_sweep<T, N, true>(g, 4, has_win ? 4 : 1, 1 << (2 * (N - 1)), has_win ? 1 : 4, obs, n, out, first_bad, index_base, stream, eval, swept));)
        if (err != cudaSuccess || swept) return err;
    }
    // Direct kernels gather from the window layout only while it is L2-resident.
    err = direct(has_win && g.win_bytes <= kWindowL2Bytes, obs, n, out, index_base, nullptr, nullptr);
    return err;
}

template cudaError_t launch_cubic_rect<double>(const DeviceGrid&, const double* const*, size_t, double*,
                                               unsigned long long*, unsigned long long, cudaStream_t);
template cudaError_t launch_cubic_rect<float>(const DeviceGrid&, const float* const*, size_t, float*,
                                              unsigned long long*, unsigned long long, cudaStream_t);

}  // namespace ib200
