// kernels.cuh — general-purpose evaluation kernels: one query point per thread, any N.
//
// These cover every (method, grid kind, dtype, N) the reference supports. Shape-specialised
// kernels for the headline configurations live in their own translation units and are chosen by
// the dispatcher in launch_*.cu; whatever they do not take falls through to the kernels here.
//
// Data layout: obs = N contiguous coordinate arrays (SoA) read coalesced, one
// element per thread per array; vals = flat C-order in HBM/L2; rectilinear axes packed back to
// back and staged once per CTA into shared memory when they fit.
#pragma once
#include "device_math.cuh"

namespace ib200 {

template <class T, int N>
struct EvalArgs {
    const T* obs[N];
    T* out;
    unsigned long long n;
    const T* vals;
    long long stride[N];
    int dim[N];
    T start[N];  // regular
    T step[N];   // regular
    const T* axes;    // rectilinear: packed axes (global)
    int axis_off[N];  // rectilinear
    int axes_total;   // rectilinear
    int axes_in_smem;
    int linearize;
    unsigned long long* first_bad;
    unsigned long long index_base;
};

constexpr int kBlock = 256;

// Stage the packed rectilinear axes into dynamic shared memory (returns the pointer to search).
template <class T, int N>
__device__ __forceinline__ const T* stage_axes(const EvalArgs<T, N>& a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (!a.axes_in_smem) return a.axes;
    T* s = reinterpret_cast<T*>(smem_raw);
    for (int i = threadIdx.x; i < a.axes_total; i += blockDim.x) s[i] = a.axes[i];
    __syncthreads();
    return s;
}

__device__ __forceinline__ void report_bad(unsigned long long* first_bad, unsigned long long idx) {
    atomicMin(first_bad, idx);
}

// ---------------------------------------------------------------------------------------------
// Multilinear (ref: multilinear/regular.rs:296-404, multilinear/rectilinear.rs:244-346)
// ---------------------------------------------------------------------------------------------

template <int D, class T, int N>
__device__ __forceinline__ T linear_tree(const T* __restrict__ p, const long long (&stride)[N], const T (&t)[N]) {
    using O = Ops<T>;
    if constexpr (D == 0) {
        return __ldg(p);
    } else {
        // Reduce dimension D-1 over the two sub-trees; dimension 0 is innermost, N-1 outermost.
        T y0 = linear_tree<D - 1, T, N>(p, stride, t);
        T y1 = linear_tree<D - 1, T, N>(p + stride[D - 1], stride, t);
        T dy = O::sub(y1, y0);
        return O::add(y0, O::mul(t[D - 1], dy));
    }
}

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) linear_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    using O = Ops<T>;
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += gstride) {
        T t[N];
        long long base = 0;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T x = a.obs[d][i];
            int origin;
            if constexpr (RECT) {
                const T* g = axes + a.axis_off[d];
                origin = clamp_cell(static_cast<long long>(lower_bound(g, a.dim[d], x)) - 1, a.dim[d] - 2);
                T x0 = g[origin];
                T x1 = g[origin + 1];
                t[d] = O::div(O::sub(x, x0), O::sub(x1, x0));
            } else {
                long long iloc = 0;
                ok = floor_cell(x, a.start[d], a.step[d], iloc) && ok;
                origin = clamp_cell(iloc, a.dim[d] - 2);
                T x0 = O::add(a.start[d], O::mul(a.step[d], O::from_int(origin)));
                t[d] = O::div(O::sub(x, x0), a.step[d]);
            }
            base += static_cast<long long>(origin) * a.stride[d];
        }
        if (!ok) {
            report_bad(a.first_bad, a.index_base + i);
            continue;
        }
        a.out[i] = linear_tree<N, T, N>(a.vals + base, a.stride, t);
    }
}

// ---------------------------------------------------------------------------------------------
// Nearest (ref: nearest/regular.rs:234-295, nearest/rectilinear.rs:193-241)
// ---------------------------------------------------------------------------------------------

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) nearest_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    using O = Ops<T>;
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const T half = T(0.5);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += gstride) {
        long long idx = 0;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T x = a.obs[d][i];
            int origin;
            T dt;
            if constexpr (RECT) {
                const T* g = axes + a.axis_off[d];
                origin = clamp_cell(static_cast<long long>(lower_bound(g, a.dim[d], x)) - 1, a.dim[d] - 2);
                T x0 = g[origin];
                T x1 = g[origin + 1];
                dt = O::div(O::sub(x, x0), O::sub(x1, x0));
            } else {
                long long iloc = 0;
                ok = floor_cell(x, a.start[d], a.step[d], iloc) && ok;
                origin = clamp_cell(iloc, a.dim[d] - 2);
                T x0 = O::add(a.start[d], O::mul(a.step[d], O::from_int(origin)));
                dt = O::div(O::sub(x, x0), a.step[d]);
            }
            int off = (dt <= half) ? 0 : 1;  // tie -> lower index; NaN (rectilinear only) -> upper
            idx += static_cast<long long>(origin + off) * a.stride[d];
        }
        if (!ok) {
            report_bad(a.first_bad, a.index_base + i);
            continue;
        }
        a.out[i] = __ldg(a.vals + idx);
    }
}

// ---------------------------------------------------------------------------------------------
// Multicubic, regular grid (ref: multicubic/regular.rs:325-469 and interp_inner :474-623)
// ---------------------------------------------------------------------------------------------

// Per-dimension quantities that depend only on (dimension, query point): hoisted out of the
// 4^(N-1)-fold repeats of the 1-D step (same inputs -> same bits).
template <class T>
struct CubicRegDim {
    T tt;      // None: t; Low: -t; High: t - 1
    T ttm1;    // tt - 1 (linearized extrapolation)
    int mode;  // CubicMode
    bool lin;  // outside the grid on this dimension AND linearize_extrapolation
};

template <class T>
__device__ __forceinline__ T cubic_regular_step(const T (&v)[4], const CubicRegDim<T>& c) {
    using O = Ops<T>;
    const T half = T(0.5);  // `/ two` is exact and equals `* 0.5` bit-for-bit
    const T two = T(2);
    if (c.mode == kModeNone) {
        T dy = O::sub(v[2], v[1]);
        T k0 = O::mul(O::sub(v[2], v[0]), half);
        T k1 = O::mul(O::sub(v[3], v[1]), half);
        return hermite(c.tt, v[1], dy, k0, k1);
    }
    T y0, y1, k0;
    if (c.mode == kModeLow) {
        y0 = v[1];
        y1 = v[0];
        k0 = O::mul(-O::sub(v[2], v[0]), half);
    } else {
        y0 = v[2];
        y1 = v[3];
        k0 = O::mul(O::sub(v[3], v[1]), half);
    }
    T dy = O::sub(y1, y0);
    T k1 = O::sub(O::mul(two, dy), k0);  // natural-spline end condition
    if (c.lin) return O::add(y1, O::mul(k1, c.ttm1));
    return hermite(c.tt, y0, dy, k0, k1);
}

template <class T>
__device__ __forceinline__ bool cubic_regular_locate(T x, T start, T step, int dim, int linearize, int& origin,
                                                     CubicRegDim<T>& c) {
    using O = Ops<T>;
    long long iloc = 0;
    bool ok = floor_cell(x, start, step, iloc);
    iloc -= 1;
    origin = clamp_cell(iloc, dim - 4);
    const long long n = dim;
    bool outside;
    if (iloc < -1) { c.mode = kModeLow; outside = true; }
    else if (iloc == -1) { c.mode = kModeLow; outside = false; }
    else if (iloc > n - 3) { c.mode = kModeHigh; outside = true; }
    else if (iloc == n - 3) { c.mode = kModeHigh; outside = false; }
    else { c.mode = kModeNone; outside = false; }
    // t is relative to footprint index 1 and its origin coordinate is never fused
    // (ref: multicubic/regular.rs:356-360).
    T x1 = O::add(start, O::mul(step, O::from_int(origin + 1)));
    T t = O::div(O::sub(x, x1), step);
    const T one = T(1);
    c.tt = c.mode == kModeNone ? t : (c.mode == kModeLow ? -t : O::sub(t, one));
    c.ttm1 = O::sub(c.tt, one);
    c.lin = outside && linearize;
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Multicubic, rectilinear grid (ref: multicubic/rectilinear.rs:265-408 and interp_inner :413-545)
// ---------------------------------------------------------------------------------------------

template <class T>
struct CubicRectDim {
    T tt;      // normalized coordinate of the active cell formula
    T ttm1;
    T wa, wc;  // a and c weights of centered_difference_nonuniform for k0 (ref: multicubic/mod.rs:104,106)
    T div0;    // the non-unit spacing ratio k0's data-dependent quotient divides by
    T wa1, wc1, div1;  // same for k1 (interior cells only)
    int mode;
    bool lin;
};

template <class T>
__device__ __forceinline__ void cubic_rect_locate(T x, const T* __restrict__ g, int dim, int linearize, int& origin,
                                                  CubicRectDim<T>& c) {
    using O = Ops<T>;
    const T one = T(1);
    const long long n = dim;
    long long iloc = static_cast<long long>(lower_bound(g, dim, x)) - 2;
    origin = clamp_cell(iloc, dim - 4);
    bool outside;
    if (iloc == -2) { c.mode = kModeLow; outside = true; }
    else if (iloc == -1) { c.mode = kModeLow; outside = false; }
    else if (iloc == n - 2) { c.mode = kModeHigh; outside = true; }
    else if (iloc == n - 3) { c.mode = kModeHigh; outside = false; }
    else { c.mode = kModeNone; outside = false; }
    const T g0 = g[origin], g1 = g[origin + 1], g2 = g[origin + 2], g3 = g[origin + 3];
    const T h01 = O::sub(g1, g0), h12 = O::sub(g2, g1), h23 = O::sub(g3, g2);
    c.wa1 = c.wc1 = c.div1 = one;
    if (c.mode == kModeNone) {
        // k0 = cdn(v0,v1,v2, h01/h12, 1);  k1 = cdn(v1,v2,v3, 1, h23/h12);  t = (x-g1)/h12
        T r = O::div(h01, h12);
        c.wa = O::div(r, O::add(r, one));
        c.wc = O::div(one, O::add(one, r));
        c.div0 = r;
        T s = O::div(h23, h12);
        c.wa1 = O::div(one, O::add(one, s));
        c.wc1 = O::div(s, O::add(s, one));
        c.div1 = s;
        c.tt = O::div(O::sub(x, g1), h12);
    } else if (c.mode == kModeLow) {
        // k0 = -cdn(v0,v1,v2, 1, h12/h01);  t = -(x-g1)/h01
        T q = O::div(h12, h01);
        c.wa = O::div(one, O::add(one, q));
        c.wc = O::div(q, O::add(q, one));
        c.div0 = q;
        c.tt = O::div(-O::sub(x, g1), h01);
    } else {
        // k0 = cdn(v1,v2,v3, h12/h23, 1);  t = (x-g2)/h23
        T p = O::div(h12, h23);
        c.wa = O::div(p, O::add(p, one));
        c.wc = O::div(one, O::add(one, p));
        c.div0 = p;
        c.tt = O::div(O::sub(x, g2), h23);
    }
    c.ttm1 = O::sub(c.tt, one);
    c.lin = outside && linearize;
}

// centered_difference_nonuniform(y0,y1,y2,h01,h12) = a*b + c*d with the unit-spacing divisions
// (x / 1.0, exact) dropped (ref: multicubic/mod.rs:103-117, rectilinear.rs:449-450).
template <class T>
__device__ __forceinline__ T cubic_rect_step(const T (&v)[4], const CubicRectDim<T>& c) {
    using O = Ops<T>;
    const T two = T(2);
    if (c.mode == kModeNone) {
        T dy = O::sub(v[2], v[1]);
        // k0: h01 = r, h12 = 1 -> b = (v2-v1)/1, d = (v1-v0)/r
        T k0 = O::add(O::mul(c.wa, dy), O::mul(c.wc, O::div(O::sub(v[1], v[0]), c.div0)));
        // k1: h01 = 1, h12 = s -> b = (v3-v2)/s, d = (v2-v1)/1
        T k1 = O::add(O::mul(c.wa1, O::div(O::sub(v[3], v[2]), c.div1)), O::mul(c.wc1, dy));
        return hermite(c.tt, v[1], dy, k0, k1);
    }
    T y0, y1, k0;
    if (c.mode == kModeLow) {
        y0 = v[1];
        y1 = v[0];
        // cdn(v0,v1,v2, 1, q): b = (v2-v1)/q, d = (v1-v0)/1
        k0 = -O::add(O::mul(c.wa, O::div(O::sub(v[2], v[1]), c.div0)), O::mul(c.wc, O::sub(v[1], v[0])));
    } else {
        y0 = v[2];
        y1 = v[3];
        // cdn(v1,v2,v3, p, 1): b = (v3-v2)/1, d = (v2-v1)/p
        k0 = O::add(O::mul(c.wa, O::sub(v[3], v[2])), O::mul(c.wc, O::div(O::sub(v[2], v[1]), c.div0)));
    }
    T dy = O::sub(y1, y0);
    T k1 = O::sub(O::mul(two, dy), k0);
    if (c.lin) return O::add(y1, O::mul(k1, c.ttm1));
    return hermite(c.tt, y0, dy, k0, k1);
}

template <class T, bool RECT>
struct CubicDimOf {
    using type = CubicRegDim<T>;
};
template <class T>
struct CubicDimOf<T, true> {
    using type = CubicRectDim<T>;
};

template <class T, bool RECT>
__device__ __forceinline__ T cubic_step(const T (&v)[4], const typename CubicDimOf<T, RECT>::type& c) {
    if constexpr (RECT) return cubic_rect_step(v, c);
    else return cubic_regular_step(v, c);
}

// Fully unrolled 4^N tree (N <= 4, the reference's "flattened" range).
template <int D, class T, int N, bool RECT>
__device__ __forceinline__ T cubic_tree(const T* __restrict__ p, const long long (&stride)[N],
                                        const typename CubicDimOf<T, RECT>::type (&c)[N]) {
    if constexpr (D == 0) {
        return __ldg(p);
    } else {
        T v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = cubic_tree<D - 1, T, N, RECT>(p + k * stride[D - 1], stride, c);
        return cubic_step<T, RECT>(v, c[D - 1]);
    }
}

// Looping 4^N tree for N = 5..8 (the reference's "recursive" range): same visiting order as the
// flattened algorithm, O(N) live values (ref: multicubic/regular.rs:368-412).
template <class T, int N, bool RECT>
__device__ __noinline__ T cubic_tree_loop(const T* __restrict__ p, const long long (&stride)[N],
                                          const typename CubicDimOf<T, RECT>::type (&c)[N]) {
    T store[N][4];
    const unsigned nverts = 1u << (2 * N);
    for (unsigned i = 0; i < nverts; ++i) {
        long long idx = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) idx += static_cast<long long>((i >> (2 * k)) & 3u) * stride[k];
        store[0][i & 3u] = __ldg(p + idx);
#pragma unroll
        for (int j = 1; j < N; ++j) {
            const unsigned q = 1u << (2 * j);
            if (((i + 1) & (q - 1)) == 0) {
                const unsigned slot = (((i + 1) >> (2 * j)) - 1) & 3u;
                store[j][slot] = cubic_step<T, RECT>(store[j - 1], c[j - 1]);
            }
        }
    }
    return cubic_step<T, RECT>(store[N - 1], c[N - 1]);
}

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) cubic_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
         i += gstride) {
        typename CubicDimOf<T, RECT>::type c[N];
        long long base = 0;
        bool ok = true;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T x = a.obs[d][i];
            int origin;
            if constexpr (RECT) {
                cubic_rect_locate(x, axes + a.axis_off[d], a.dim[d], a.linearize, origin, c[d]);
            } else {
                ok = cubic_regular_locate(x, a.start[d], a.step[d], a.dim[d], a.linearize, origin, c[d]) && ok;
            }
            base += static_cast<long long>(origin) * a.stride[d];
        }
        if (!ok) {
            report_bad(a.first_bad, a.index_base + i);
            continue;
        }
        if constexpr (N <= 4) a.out[i] = cubic_tree<N, T, N, RECT>(a.vals + base, a.stride, c);
        else a.out[i] = cubic_tree_loop<T, N, RECT>(a.vals + base, a.stride, c);
    }
}

}  // namespace ib200
