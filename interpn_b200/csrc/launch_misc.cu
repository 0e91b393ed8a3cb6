// launch_misc.cu — one_dim evaluators, check_bounds, and the method dispatcher.
#include <cstdlib>
#include <mutex>

#include "launch_common.cuh"

namespace ib200 {

size_t sweep_env(const char* name, size_t fallback) {
    const char* e = getenv(name);
    return e && *e ? static_cast<size_t>(strtoull(e, nullptr, 10)) : fallback;
}

size_t sweep_env_common(const char* name, size_t fallback) { return sweep_env(name, fallback); }

cudaMemPool_t sweep_scratch_pool() {
    static cudaMemPool_t pools[64] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceGetDefaultMemPool(&pools[dev], dev);  // fall back to the default pool, untouched
        } else {
            uint64_t keep = UINT64_MAX;  // keep the scratch cached between calls (several GB for C3-cubic)
            cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return pools[dev];
}

bool force_index64() {
    const char* e = getenv("INTERPN_B200_INDEX64");
    return e && e[0] == '1';
}

// ---------------------------------------------------------------------------------------------
// Window layout builder: win[f*W + j] = vals[min(f + j, nvals - 1)]  (kernels.cuh load_row)
// ---------------------------------------------------------------------------------------------

template <class T, int W>
__global__ void __launch_bounds__(kBlock) build_window_kernel(const T* __restrict__ vals, T* __restrict__ win,
                                                              unsigned long long nvals) {
    const unsigned long long total = nvals * W;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
         k += gstride) {
        unsigned long long src = k / W + k % W;
        win[k] = vals[src < nvals ? src : nvals - 1];
    }
}

struct HyperDims {
    int dim[kMaxNd];
    long long stride[kMaxNd];
};

// Patch layout (kernels.cuh linear_patches): pwin[f*4 + 2*i + j] = vals[f + i*Db + j], with i and j dropped to 0 where
// the patch would leave the grid (never read: a footprint origin is at most dim-2).
template <class T>
__global__ void __launch_bounds__(kBlock) build_pwindow_kernel(const T* __restrict__ vals, T* __restrict__ pwin,
                                                               unsigned long long nvals, unsigned long long da,
                                                               unsigned long long db) {
    const unsigned long long total = nvals * 4;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
         k += gstride) {
        const unsigned long long f = k >> 2, i = (k >> 1) & 1, j = k & 1;
        const unsigned long long ib = f % db, ia = (f / db) % da;
        const unsigned long long ii = ia + i < da ? i : 0, jj = ib + j < db ? j : 0;
        pwin[k] = vals[f + ii * db + jj];
    }
}

// Hypercube layout (kernels.cuh linear_hyper_kernel, linear_hyper3_kernel): hwin[f*2^nbits + v] = vals[f + sum_b bit_b(v)*stride_{N-nbits+b}]
// (nbits = 4 for N = 4..6, 3 for N = 3); offsets
// that would leave the grid along a dimension are dropped (never read: a footprint origin is at most dim-2).
template <class T>
__global__ void __launch_bounds__(kBlock) build_hwindow_kernel(const T* __restrict__ vals, T* __restrict__ hwin,
                                                               unsigned long long nvals, int nbits, const __grid_constant__ HyperDims hd) {
    const unsigned long long total = nvals << nbits;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < total; k += gstride) {
        const unsigned long long f = k >> nbits;
        const unsigned v = static_cast<unsigned>(k) & ((1u << nbits) - 1u);
        unsigned long long src = f;
        for (int b = 0; b < nbits; ++b) {  // hd holds the last `nbits` dimensions
            const unsigned long long id = (f / static_cast<unsigned long long>(hd.stride[b])) % static_cast<unsigned long long>(hd.dim[b]);
            if (((v >> b) & 1u) && id + 1 < static_cast<unsigned long long>(hd.dim[b])) src += static_cast<unsigned long long>(hd.stride[b]);
        }
        hwin[k] = vals[src];
    }
}

cudaError_t launch_build_window(const DeviceGrid& g, cudaStream_t stream) {
    if (!g.win || g.nvals == 0) return cudaSuccess;
    if (g.method == 0 && ((g.ndims >= 4 && g.win_width == 16) || (g.ndims == 3 && g.win_width == 8))) {  // INTERPN_B200_LINEAR, hypercube layout
        const int nbits = g.ndims == 3 ? 3 : 4;
        HyperDims hd{};
        for (int b = 0; b < nbits; ++b) {
            hd.dim[b] = g.dim[g.ndims - nbits + b];
            hd.stride[b] = g.stride[g.ndims - nbits + b];
        }
        const unsigned grid_dim = grid_for(g.nvals << nbits, g.sm_count, 8);
        if (g.elem == 8) build_hwindow_kernel<double><<<grid_dim, kBlock, 0, stream>>>(static_cast<const double*>(g.vals), static_cast<double*>(g.win), g.nvals, nbits, hd);
        else build_hwindow_kernel<float><<<grid_dim, kBlock, 0, stream>>>(static_cast<const float*>(g.vals), static_cast<float*>(g.win), g.nvals, nbits, hd);
        count_launch();
        return cudaGetLastError();
    }
    const unsigned grid_dim = grid_for(g.nvals * g.win_width, g.sm_count, 8);
    if (g.win_cross && g.method == 0) {  // INTERPN_B200_LINEAR
        const unsigned long long da = g.dim[g.ndims - 2], db = g.dim[g.ndims - 1];
        if (g.elem == 8) build_pwindow_kernel<double><<<grid_dim, kBlock, 0, stream>>>(static_cast<const double*>(g.vals), static_cast<double*>(g.win), g.nvals, da, db);
        else build_pwindow_kernel<float><<<grid_dim, kBlock, 0, stream>>>(static_cast<const float*>(g.vals), static_cast<float*>(g.win), g.nvals, da, db);
        count_launch();
        return cudaGetLastError();
    }
    if (g.win_cross) return launch_build_coef_window(g, stream);  // INTERPN_B200_CUBIC, N = 2..4
    if (g.elem == 8) {
        if (g.win_width == 4) build_window_kernel<double, 4><<<grid_dim, kBlock, 0, stream>>>(static_cast<const double*>(g.vals), static_cast<double*>(g.win), g.nvals);
        else build_window_kernel<double, 2><<<grid_dim, kBlock, 0, stream>>>(static_cast<const double*>(g.vals), static_cast<double*>(g.win), g.nvals);
    } else {
        if (g.win_width == 4) build_window_kernel<float, 4><<<grid_dim, kBlock, 0, stream>>>(static_cast<const float*>(g.vals), static_cast<float*>(g.win), g.nvals);
        else build_window_kernel<float, 2><<<grid_dim, kBlock, 0, stream>>>(static_cast<const float*>(g.vals), static_cast<float*>(g.win), g.nvals);
    }
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// one_dim (ref: one_dim/mod.rs:85-187, one_dim/linear.rs:24-85, one_dim/hold.rs:23-107)
// ---------------------------------------------------------------------------------------------

template <class T>
struct OneDimArgs {
    int kind;
    T start, stop, step;  // regular; `stop` computed once like RegularGrid1D::new
    const T* grid;        // rectilinear
    const T* vals;
    int nvals;
    const T* locs;
    T* out;
    unsigned long long n;
    unsigned long long* first_bad;
    unsigned long long index_base;
};

enum : int { kInside = 0, kOutsideLow = 1, kOutsideHigh = 2 };

template <class T, bool RECT>
__global__ void __launch_bounds__(kBlock) one_dim_kernel(const __grid_constant__ OneDimArgs<T> a) {
    using O = Ops<T>;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += gstride) {
        const T loc = a.locs[i];
        int cell, extrap = kInside;
        T x0, x1;
        if constexpr (RECT) {
            cell = clamp_cell(lower_bound(a.grid, a.nvals, loc) - 1, a.nvals - 2);
            if (loc < a.grid[0]) extrap = kOutsideLow;
            else if (loc > a.grid[a.nvals - 1]) extrap = kOutsideHigh;
            x0 = a.grid[cell];
            x1 = a.grid[cell + 1];
        } else {
            if (loc > a.stop) extrap = kOutsideHigh;
            else if (loc < a.start) extrap = kOutsideLow;
            int iloc = 0;
            if (!floor_cell(loc, a.start, a.step, a.step, false, iloc)) {
                atomicMin(a.first_bad, a.index_base + i);
                continue;
            }
            cell = clamp_cell(iloc, a.nvals - 2);
            x0 = O::add(a.start, O::mul(a.step, O::from_int(cell)));
            x1 = O::add(x0, a.step);
        }
        const T y0 = __ldg(a.vals + cell);
        const T y1 = __ldg(a.vals + cell + 1);
        T v;
        switch (a.kind) {
            case 0:    // Linear1D
            case 1: {  // LinearHoldLast1D
                if (a.kind == 1 && extrap != kInside) {
                    v = extrap == kOutsideLow ? y0 : y1;
                } else {
                    T slope = O::div(O::sub(y1, y0), O::sub(x1, x0));
                    T dx = O::sub(loc, x0);
                    v = muladd(slope, dx, y0);  // fused under the fma feature (one_dim/linear.rs:30-35)
                }
            } break;
            case 2: v = extrap == kOutsideHigh ? y1 : y0; break;  // Left1D
            case 3: v = extrap == kOutsideLow ? y0 : y1; break;   // Right1D
            default: {                                             // Nearest1D: tie -> left
                T dx0 = O::abs(O::sub(loc, x0));
                T dx1 = O::abs(O::sub(loc, x1));
                v = (dx1 >= dx0) ? y0 : y1;
            } break;
        }
        a.out[i] = v;
    }
}

template <class T>
cudaError_t launch_one_dim(int kind, bool rect, T start, T step, const T* grid, const T* vals, size_t nvals,
                           const T* locs, size_t n, T* out, unsigned long long* first_bad,
                           unsigned long long index_base, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    OneDimArgs<T> a{};
    a.kind = kind;
    a.start = start;
    a.step = step;
    // RegularGrid1D::new: stop = start + step * (T)(len - 1)  (ref: one_dim/mod.rs:86-88), in T arithmetic.
    volatile T prod = step * static_cast<T>(nvals - 1);
    a.stop = start + prod;
    a.grid = grid;
    a.vals = vals;
    a.nvals = static_cast<int>(nvals);
    a.locs = locs;
    a.out = out;
    a.n = n;
    a.first_bad = first_bad;
    a.index_base = index_base;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned grid_dim = grid_for(n, sms, 8);
    if (rect) one_dim_kernel<T, true><<<grid_dim, kBlock, 0, stream>>>(a);
    else one_dim_kernel<T, false><<<grid_dim, kBlock, 0, stream>>>(a);
    count_launch();
    return cudaGetLastError();
}

template cudaError_t launch_one_dim<double>(int, bool, double, double, const double*, const double*, size_t,
                                            const double*, size_t, double*, unsigned long long*, unsigned long long,
                                            cudaStream_t);
template cudaError_t launch_one_dim<float>(int, bool, float, float, const float*, const float*, size_t, const float*,
                                           size_t, float*, unsigned long long*, unsigned long long, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// check_bounds (ref: multilinear/regular.rs:168-171, multilinear/rectilinear.rs:124-128):
// bad = any((x - lo) <= -atol || (x - hi) >= atol). Streaming OR-reduction over one axis.
// ---------------------------------------------------------------------------------------------

template <class T>
__global__ void __launch_bounds__(kBlock) check_bounds_kernel(const T* __restrict__ x, unsigned long long n, T lo, T hi,
                                                              T atol, int* flag) {
    using O = Ops<T>;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    bool bad = false;
    const T natol = -atol;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += gstride) {
        T v = x[i];
        bad = bad || (O::sub(v, lo) <= natol) || (O::sub(v, hi) >= atol);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

template <class T>
cudaError_t launch_check_bounds(const T* x, size_t n, T lo, T hi, T atol, int* flag, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    check_bounds_kernel<T><<<grid_for(n, sms, 8), kBlock, 0, stream>>>(x, n, lo, hi, atol, flag);
    count_launch();
    return cudaGetLastError();
}

template cudaError_t launch_check_bounds<double>(const double*, size_t, double, double, double, int*, cudaStream_t);
template cudaError_t launch_check_bounds<float>(const float*, size_t, float, float, float, int*, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// Method dispatcher
// ---------------------------------------------------------------------------------------------

template <class T>
cudaError_t launch_eval(const DeviceGrid& g, const T* const* obs, size_t n, T* out, unsigned long long* first_bad,
                        unsigned long long index_base, cudaStream_t stream) {
    switch (g.method) {
        case 0: return launch_linear<T>(g, obs, n, out, first_bad, index_base, stream);
        case 1:
            return g.rect ? launch_cubic_rect<T>(g, obs, n, out, first_bad, index_base, stream)
                          : launch_cubic_regular<T>(g, obs, n, out, first_bad, index_base, stream);
        case 2: return launch_nearest<T>(g, obs, n, out, first_bad, index_base, stream);
        default: return cudaErrorInvalidValue;
    }
}

template cudaError_t launch_eval<double>(const DeviceGrid&, const double* const*, size_t, double*, unsigned long long*,
                                         unsigned long long, cudaStream_t);
template cudaError_t launch_eval<float>(const DeviceGrid&, const float* const*, size_t, float*, unsigned long long*,
                                        unsigned long long, cudaStream_t);

}  // namespace ib200
