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
#include <type_traits>

#include "device_math.cuh"

namespace ib200 {

template <class T, int N>
struct EvalArgs {
    const T* obs[N];
    T* out;
    unsigned long long n;
    const T* vals;
    const T* win;  // window layout of vals (see load_row), or nullptr
    long long stride[N];
    int istride[N];  // the same strides when the grid has fewer than 2^31 values (32-bit index arithmetic), else 0
    int dim[N];
    T start[N];  // regular
    T step[N];   // regular
    T rstep[N];  // regular: RN(1/step), for exact_div
    int fast_div;  // regular: every step is within exact_div's exponent range
    double hstep[N], tau[N], lim[N];  // regular f64: step/2, step*2^-54, step*(1-2^-20) (device_math.cuh fast_cell)
    const T* axes;    // rectilinear: packed axes (global)
    int axis_off[N];  // rectilinear
    int axes_total;   // rectilinear: elements of the blob (axes, reciprocal cell widths, bucket tables)
    int rect_fast, rect_fast_div;  // rectilinear: search / division accelerators are valid (capi.cu rect_new)
    int rc_off[N], lut_off[N], lut_nb[N];
    int rect_cell, clut_off[N], clut_nb[N];  // rectilinear linear / nearest: cell tables (rect_cell_locate)
    T clut_scale[N];
    int ct_off[N];  // rectilinear cubic: per-cell constant tables (capi.cu cubic_cell_table), valid when rect_cubic_table
    int rect_cubic_table;
    T lut_scale[N];
    int axes_in_smem;
    int linearize;
    unsigned long long* first_bad;
    unsigned long long index_base;
    int slab_lo, slab_hi;  // slab passes (linear_slab_kernel): cells [slab_lo, slab_hi) of dimension 0; hi < 0 = to the end
    const unsigned* remap;  // bin-swept evaluation: original (chunk-local) index of point i, else nullptr
    unsigned long long* work;  // bin-swept evaluation: zeroed counter for dynamic block scheduling, else nullptr
};

constexpr int kBlock = 256;

// Index arithmetic of the streaming kernels is 32-bit whenever the grid has fewer than 2^31 values
// (I = int); the 64-bit variants (I = long long) cover everything else.
template <class I, class T, int N>
__device__ __forceinline__ const I (&strides_of(const EvalArgs<T, N>& a))[N] {
    if constexpr (sizeof(I) == 4) return a.istride;
    else return a.stride;
}

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

// The same with the address space known at compile time: AXSM = the blob is staged and the returned pointer is a
// shared-memory pointer the compiler can see through (every table access becomes LDS with 32-bit address arithmetic instead
// of a generic load: the rectilinear streaming kernels spent a third of their stall samples behind generic loads,
// profiles/r2_c5_n2rect_f32_ncu.json), else the global pointer. The kernels branch once on a.axes_in_smem.
template <bool AXSM, class T, int N>
__device__ __forceinline__ const T* stage_axes_as(const EvalArgs<T, N>& a) {
    if constexpr (AXSM) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        T* s = reinterpret_cast<T*>(smem_raw);
        for (int i = threadIdx.x; i < a.axes_total; i += blockDim.x) s[i] = a.axes[i];
        __syncthreads();
        return s;
    } else {
        return a.axes;
    }
}

// partition_point(|g| g < x) on axis d. On strictly increasing axes a bucket table narrows the bisection to
// the nodes of three adjacent buckets (two buckets per node on average): the computed bucket is within one of
// the true one, lut[k] = partition_point(g < edge_k) is monotone in k, so the answer lies in
// [lut[b-1], lut[b+2]] and any correct search inside that range returns the reference's index. NaN compares
// false everywhere -> 0, +-inf saturate to the end buckets, like slice::partition_point.
template <class T, int N>
__device__ __forceinline__ int rect_lower_bound(const EvalArgs<T, N>& a, const T* __restrict__ axes, int d, T x) {
    const T* g = axes + a.axis_off[d];
    if (!a.rect_fast) return lower_bound(g, a.dim[d], x);
    const int* lut = reinterpret_cast<const int*>(axes + a.lut_off[d]);
    const int nb = a.lut_nb[d];
    const int b = min(max(Ops<T>::floor_sat((x - g[0]) * a.lut_scale[d]), 0), nb - 1);
    int lo = lut[max(b - 1, 0)];
    int len = lut[min(b + 2, nb)] - lo;
    while (len > 0) {
        const int half = len >> 1;
        const bool below = g[lo + half] < x;
        lo = below ? lo + half + 1 : lo;
        len = below ? len - half - 1 : half;
    }
    return lo;
}

// Cell of x on a strictly increasing axis, for the multilinear / nearest kernels: origin = clamp(partition_point(g < x)
// - 1, 0, n-2) and its two nodes, from THREE shared-memory loads (the bucket search above costs 5-6, and these kernels
// sit on the shared-memory wavefront rate: ncu l1tex 93 %, profiles/r1_p5_c5n3_rect_ncu.json). A table with four
// buckets per node gives the cell c0 that contains the bucket's left edge; with at most one node inside a bucket the
// answer is c0 or c0+1, decided by g[c0+1] < x, and the third load (g[c0] or g[c0+2]) both completes the cell and
// PROVES it: the chosen cell must satisfy g[c] < x <= g[c+1] (or be the clamped end cell). Anything else — a bucket
// index off by one at a bucket edge, two nodes in one bucket — fails the proof and takes the plain bisection.
// NaN lands in cell 0 like slice::partition_point (every compare false).
// The rare unproven point: plain bisection of the whole axis (inline: an out-of-line call cost the callers a stack frame).
template <class T>
static __device__ __forceinline__ int rect_cell_bisect(const T* __restrict__ g, int n, T x) {
    return clamp_cell(lower_bound(g, n, x) - 1, n - 2);
}

template <class T, int N>
__device__ __forceinline__ int rect_cell_locate(const EvalArgs<T, N>& a, const T* __restrict__ axes, int d, T x, T& g0, T& g1) {
    const T* g = axes + a.axis_off[d];
    const int n = a.dim[d];
    if (a.rect_cell) {
        const int* lut = reinterpret_cast<const int*>(axes + a.clut_off[d]);
        const int b = min(max(Ops<T>::floor_sat((x - g[0]) * a.clut_scale[d]), 0), a.clut_nb[d] - 1);
        const int c0 = lut[b];
        const T gb = g[c0 + 1];
        const bool up = gb < x && c0 < n - 2;
        const T other = g[up ? c0 + 2 : c0];
        const bool proven = up ? (!(other < x) || c0 + 1 == n - 2) : (other < x || c0 == 0);
        if (proven) {
            g0 = up ? gb : other;
            g1 = up ? other : gb;
            return c0 + (up ? 1 : 0);
        }
    }
    const int origin = a.rect_cell ? rect_cell_bisect<T>(g, n, x) : clamp_cell(rect_lower_bound<T, N>(a, axes, d, x) - 1, n - 2);
    g0 = g[origin];
    g1 = g[origin + 1];
    return origin;
}

// Records an unrepresentable query point: the smallest failing index of the caller's batch wins.
template <class T, int N>
__device__ __forceinline__ void report_bad(const EvalArgs<T, N>& a, unsigned long long i) {
    atomicMin(a.first_bad, a.index_base + (a.remap ? static_cast<unsigned long long>(a.remap[i]) : i));
}

// Which block of blockDim.x work items this CTA takes next: statically strided over the grid, or — when
// the launch carries a work counter (bin-swept evaluation) — handed out on demand in ascending order, so
// that all CTAs stay within ~gridDim blocks of each other along the sorted batch no matter how their
// speeds differ. Returns false when the block starts at or beyond `nitems`. Uniform across the CTA.
struct BlockSchedule {
    unsigned long long blk;
    bool started;
    __device__ __forceinline__ BlockSchedule() : blk(blockIdx.x), started(false) {}
    __device__ __forceinline__ bool next(unsigned long long* work, unsigned long long nitems) {
        if (work) {
            __shared__ unsigned long long s_blk;
            __syncthreads();
            if (threadIdx.x == 0) s_blk = atomicAdd(work, 1ull);
            __syncthreads();
            blk = s_blk;
        } else if (started) {
            blk += gridDim.x;
        }
        started = true;
        return blk * blockDim.x < nitems;
    }
};

// ---------------------------------------------------------------------------------------------
// Row gathers. The last grid dimension is contiguous, so every footprint (2^N / 4^N corners) is a
// set of rows of W = 2 / 4 consecutive values. Mid-size grids additionally keep a *window layout*
// in HBM/L2: win[f*W + j] = vals[f + j], i.e. the row starting at ANY flat index f is one naturally
// aligned W*sizeof(T) vector — one 32-byte sector, one LDG.256/LDG.128 — instead of W scalar loads
// that straddle two sectors 75 % (W=4) of the time.
// ---------------------------------------------------------------------------------------------

template <class T, int W, bool WIN, class I = long long>
__device__ __forceinline__ void load_row(const T* __restrict__ vals, const T* __restrict__ win, I idx, T (&r)[W]) {
    if constexpr (!WIN) {
#pragma unroll
        for (int j = 0; j < W; ++j) r[j] = __ldg(vals + idx + j);
    } else if constexpr (sizeof(T) == 8 && W == 4) {
#ifdef IB200_LDG_NOALLOC
        asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
#else
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
#endif
            : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3])
            : "l"(win + static_cast<long long>(idx) * 4));
    } else if constexpr (sizeof(T) == 8 && W == 2) {
        double2 q = __ldg(reinterpret_cast<const double2*>(win) + idx);
        r[0] = q.x; r[1] = q.y;
    } else if constexpr (sizeof(T) == 4 && W == 4) {
        float4 q = __ldg(reinterpret_cast<const float4*>(win) + idx);
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
    } else {
        float2 q = __ldg(reinterpret_cast<const float2*>(win) + idx);
        r[0] = q.x; r[1] = q.y;
    }
}

// A row of two values straight from `vals` as ONE aligned pair load plus, for odd flat indices only, a predicated
// scalar load: on average 1.5 L1 wavefronts per lane instead of the 2 of two scalar loads (the slab-pass kernel sits
// on the L1 wavefront rate: l1tex 85 %, profiles/r1_p6_c3l_slab_ncu.json). `vals` is the library's own allocation
// (256-byte aligned), and the pair that contains idx never reaches past idx + 1.
template <class T, class I>
__device__ __forceinline__ void load_row_aligned(const T* __restrict__ vals, I idx, T (&r)[2]) {
    using V = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
    const V q = __ldg(reinterpret_cast<const V*>(vals) + (idx >> 1));
    if (idx & 1) {
        r[0] = q.y;
        r[1] = __ldg(vals + idx + 1);
    } else {
        r[0] = q.x;
        r[1] = q.y;
    }
}

// Query coordinates and results are touched exactly once: streaming loads/stores (evict-first) keep
// them from displacing the grid in L1/L2.
template <class T>
__device__ __forceinline__ T load_query(const T* p) { return __ldcs(p); }
template <class T>
__device__ __forceinline__ void store_result(T* p, T v) { __stcs(p, v); }

// P consecutive coordinates / results as one vector access (p must be P*sizeof(T)-aligned).
template <class T, int P>
__device__ __forceinline__ void load_query_vec(const T* p, T (&v)[P]) {
    if constexpr (P == 1) {
        v[0] = __ldcs(p);
    } else if constexpr (sizeof(T) == 8 && P == 2) {
        double2 q = __ldcs(reinterpret_cast<const double2*>(p));
        v[0] = q.x; v[1] = q.y;
    } else if constexpr (sizeof(T) == 8 && P == 4) {
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    } else if constexpr (sizeof(T) == 4 && P == 2) {
        float2 q = __ldcs(reinterpret_cast<const float2*>(p));
        v[0] = q.x; v[1] = q.y;
    } else {
        static_assert(sizeof(T) == 4 && P == 4, "unsupported points-per-thread");
        float4 q = __ldcs(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
}
template <class T, int P>
__device__ __forceinline__ void store_result_vec(T* p, const T (&v)[P]) {
    if constexpr (P == 1) {
        __stcs(p, v[0]);
    } else if constexpr (sizeof(T) == 8 && P == 2) {
        __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
    } else if constexpr (sizeof(T) == 8 && P == 4) {
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    } else if constexpr (sizeof(T) == 4 && P == 2) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
}

// ---------------------------------------------------------------------------------------------
// Multilinear (ref: multilinear/regular.rs:296-404, multilinear/rectilinear.rs:244-346)
// ---------------------------------------------------------------------------------------------

// Reduces dimensions 0..D-1 of the sub-block at flat index `idx` for both positions of the last
// (contiguous) dimension at once. Each lerp is the reference's `y0 + t*(y1 - y0)` with dimension 0
// innermost, so every output is bit-identical to the reference's tree; only the load schedule differs.
template <int D, class T, int N, bool WIN, class I, bool AL = false>
__device__ __forceinline__ void linear_rows(const T* __restrict__ vals, const T* __restrict__ win, I idx,
                                            const I (&stride)[N], const T (&t)[N], T (&out)[2]) {
    using O = Ops<T>;
    if constexpr (D == 0) {
        if constexpr (AL) load_row_aligned<T, I>(vals, idx, out);
        else load_row<T, 2, WIN, I>(vals, win, idx, out);
    } else {
        T lo[2], hi[2];
        linear_rows<D - 1, T, N, WIN, I, AL>(vals, win, idx, stride, t, lo);
        linear_rows<D - 1, T, N, WIN, I, AL>(vals, win, idx + stride[D - 1], stride, t, hi);
#pragma unroll
        for (int j = 0; j < 2; ++j) out[j] = muladd(t[D - 1], O::sub(hi[j], lo[j]), lo[j]);
    }
}

// Patch layout (multilinear, N >= 2): pwin[f*4 + 2*i + j] = vals[f + i*D_{N-1} + j] — the 32-byte sector (f64) at flat
// index f holds the 2x2 patch of the last two dimensions that starts at f, so a footprint is 2^(N-2) sectors = L1
// wavefronts instead of 2^(N-1) rows (the multilinear kernels on L2-resident grids sit on the L1 wavefront rate,
// DESIGN.md §4.1). Reduces dimensions 0..D-1 for the four patch positions at once; the caller finishes with
// dimension N-2 (between the patch's two rows) and N-1, so every lerp and their order are the reference's.
template <int D, class T, int N, class I>
__device__ __forceinline__ void linear_patches(const T* __restrict__ win, I idx, const I (&stride)[N], const T (&t)[N],
                                               T (&out)[4]) {
    using O = Ops<T>;
    if constexpr (D == 0) {
        load_row<T, 4, true, I>(nullptr, win, idx, out);
    } else {
        T lo[4], hi[4];
        linear_patches<D - 1, T, N, I>(win, idx, stride, t, lo);
        linear_patches<D - 1, T, N, I>(win, idx + stride[D - 1], stride, t, hi);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = muladd(t[D - 1], O::sub(hi[j], lo[j]), lo[j]);
    }
}

// The whole lerp tree of one located point. WL = layout gathered from: 0 the grid itself (1: the same through
// load_row_aligned), 2 the row-pair window copy, 4 the 2x2 patch copy.
template <class T, int N, int WL, class I>
__device__ __forceinline__ T linear_tree(const T* __restrict__ vals, const T* __restrict__ win, I base,
                                         const I (&stride)[N], const T (&t)[N]) {
    using O = Ops<T>;
    constexpr bool WIN = WL >= 2;
    if constexpr (WL == 4) {
        static_assert(N >= 2, "the patch layout needs two dimensions");
        T v[4];
        linear_patches<N - 2, T, N, I>(win, base, stride, t, v);
        const T r0 = muladd(t[N - 2], O::sub(v[2], v[0]), v[0]);
        const T r1 = muladd(t[N - 2], O::sub(v[3], v[1]), v[1]);
        return muladd(t[N - 1], O::sub(r1, r0), r0);
    } else {
        T r[2];
        linear_rows<N - 1, T, N, WIN, I, WL == 1>(vals, win, base, stride, t, r);
        return muladd(t[N - 1], O::sub(r[1], r[0]), r[0]);
    }
}

// Locate one query point: per-dimension cell origin -> flat index of the footprint's first corner,
// and the normalized coordinates t. Returns false for an unrepresentable coordinate (regular grids).
// CELL: rectilinear axes are located through the cell table (rect_cell_locate) — used by the kernels that gather from
// an L2-resident window copy; the kernels for grids beyond L2 are DRAM-bound, gain nothing from a cheaper search and
// lost 5 % to its extra live registers (C3-linear), so they keep the bucket search.
template <class T, int N, bool RECT, class I, bool CELL = false>
__device__ __forceinline__ bool linear_locate(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&xs)[N],
                                              T (&t)[N], I& base) {
    const I(&stride)[N] = strides_of<I>(a);
    using O = Ops<T>;
    bool ok = true;
    base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const T x = xs[d];
        int origin;
        if constexpr (RECT) {
            T x0, x1;
            if constexpr (CELL) {
                origin = rect_cell_locate<T, N>(a, axes, d, x, x0, x1);
            } else {
                const T* g = axes + a.axis_off[d];
                origin = clamp_cell(rect_lower_bound<T, N>(a, axes, d, x) - 1, a.dim[d] - 2);
                x0 = g[origin];
                x1 = g[origin + 1];
            }
            const T e = O::sub(x, x0), h = O::sub(x1, x0);
            if constexpr (sizeof(T) == 8) {
                // the cell's reciprocal width is tabulated: exact_div instead of the IEEE division
                t[d] = a.rect_fast_div ? exact_div(e, h, axes[a.rc_off[d] + origin], true) : O::div(e, h);
            } else {
                t[d] = O::div(e, h);
            }
        } else {
            int iloc = 0;
            ok = floor_cell(x, a.start[d], a.step[d], a.rstep[d], a.fast_div != 0, iloc) && ok;
            origin = clamp_cell(iloc, a.dim[d] - 2);
            // x0: fused under the fma feature by the flattened structs (N <= 6), not by the recursive twins
            T x0 = muladd<(N <= 6)>(a.step[d], O::from_int(origin), a.start[d]);
            t[d] = exact_div(O::sub(x, x0), a.step[d], a.rstep[d], a.fast_div != 0);
        }
        base += static_cast<I>(origin) * stride[d];
    }
    return ok;
}

// Nearest: flat index of the chosen node (ref: nearest/regular.rs:259-293, nearest/rectilinear.rs:213-239).
template <class T, int N, bool RECT, class I>
__device__ __forceinline__ bool nearest_locate(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&xs)[N],
                                               I& idx) {
    const I(&stride)[N] = strides_of<I>(a);
    using O = Ops<T>;
    const T half = T(0.5);
    bool ok = true;
    idx = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const T x = xs[d];
        int origin;
        T dt;
        if constexpr (RECT) {
            T x0, x1;
            origin = rect_cell_locate<T, N>(a, axes, d, x, x0, x1);
            const T e = O::sub(x, x0), h = O::sub(x1, x0);
            if (a.rect_fast_div) {  // division-free and exact: device_math.cuh nearest_upper (f64 and f32 thresholds)
                const T tau = O::mul(h, sizeof(T) == 8 ? T(0x1p-54) : T(0x1p-25));
                idx += static_cast<I>(origin + (nearest_upper(e, O::mul(h, T(0.5)), tau) ? 1 : 0)) * stride[d];
                continue;
            }
            if constexpr (sizeof(T) == 8) dt = exact_div_slow(e, h);  // axes outside the guarded range only: out of line
            else dt = O::div(e, h);
        } else {
            int iloc = 0;
            ok = floor_cell(x, a.start[d], a.step[d], a.rstep[d], a.fast_div != 0, iloc) && ok;
            origin = clamp_cell(iloc, a.dim[d] - 2);
            T x0 = muladd(a.step[d], O::from_int(origin), a.start[d]);
            dt = exact_div(O::sub(x, x0), a.step[d], a.rstep[d], a.fast_div != 0);
        }
        const int off = (dt <= half) ? 0 : 1;  // tie -> lower index; NaN (rectilinear only) -> upper
        idx += static_cast<I>(origin + off) * stride[d];
    }
    return ok;
}

// Division-free twins of linear_locate / nearest_locate for regular grids (device_math.cuh fast_cell & co., f64 and
// f32). They return false when a point needs the exact path; `base` / `idx` stay in range either way.
template <class T, int N, class I>
__device__ __forceinline__ bool linear_locate_fast(const EvalArgs<T, N>& a, const T (&xs)[N], T (&t)[N], I& base) {
    using O = Ops<T>;
    const I(&stride)[N] = strides_of<I>(a);
    bool sure = a.fast_div != 0;
    base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        int origin;
        T od, dd;
        sure = fast_cell(xs[d], a.start[d], a.step[d], a.rstep[d], static_cast<T>(a.lim[d]), a.dim[d], origin, od, dd) && sure;
        const T x0 = muladd<(N <= 6)>(a.step[d], od, a.start[d]);
        const T e = O::sub(xs[d], x0);
        sure = markstein_operand_ok(e) && sure;
        t[d] = markstein_div(e, a.step[d], a.rstep[d]);
        base += static_cast<I>(origin) * stride[d];
    }
    return sure;
}

template <class T, int N, class I>
__device__ __forceinline__ bool nearest_locate_fast(const EvalArgs<T, N>& a, const T (&xs)[N], I& idx) {
    using O = Ops<T>;
    const I(&stride)[N] = strides_of<I>(a);
    bool sure = a.fast_div != 0;
    idx = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        int origin;
        T od, dd;
        sure = fast_cell(xs[d], a.start[d], a.step[d], a.rstep[d], static_cast<T>(a.lim[d]), a.dim[d], origin, od, dd) && sure;
        const T x0 = muladd(a.step[d], od, a.start[d]);
        const T e = O::sub(xs[d], x0);
        const int off = nearest_upper(e, static_cast<T>(a.hstep[d]), static_cast<T>(a.tau[d])) ? 1 : 0;
        idx += static_cast<I>(origin + off) * stride[d];
    }
    return sure;
}

// One point of the streaming kernels: fast path where it exists, exact path otherwise.
template <class T, int N, bool RECT, class I, bool CELL = false>
__device__ __forceinline__ bool linear_locate_any(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&xs)[N],
                                                  T (&t)[N], I& base) {
    if constexpr (!RECT) {
        if (linear_locate_fast<T, N, I>(a, xs, t, base)) return true;
    }
    return linear_locate<T, N, RECT, I, CELL>(a, axes, xs, t, base);
}
template <class T, int N, bool RECT, class I>
__device__ __forceinline__ bool nearest_locate_any(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&xs)[N],
                                                   I& idx) {
    if constexpr (!RECT) {
        if (nearest_locate_fast<T, N, I>(a, xs, idx)) return true;
    }
    return nearest_locate<T, N, RECT, I>(a, axes, xs, idx);
}

// The streaming kernels (multilinear, nearest) are a chain  DRAM load -> locate -> gather -> store
// per point; with one point per thread the SM runs out of warps long before HBM runs out of
// bandwidth. Each thread therefore owns P consecutive points: one 16/32-byte vector load per
// coordinate array, P independent locate/gather chains in flight, one vector store. The host picks
// P > 1 only when every coordinate array and `out` are P*sizeof(T)-aligned (launch_common.cuh); the
// n % P tail is evaluated one point per thread.
template <class T, int N, bool RECT, int WL, int P, class I, bool AXSM>
__device__ __forceinline__ void linear_body(const EvalArgs<T, N>& a) {
    using O = Ops<T>;
    const I(&stride)[N] = strides_of<I>(a);
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes_as<AXSM, T, N>(a);
    const unsigned long long gtid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long ngroups = a.n / P;
    BlockSchedule sched;
    while (sched.next(a.work, ngroups)) {
        const unsigned long long g = sched.blk * blockDim.x + threadIdx.x;
        if (g >= ngroups) continue;
        const unsigned long long i0 = g * P;
        T xs[P][N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T v[P];
            load_query_vec<T, P>(a.obs[d] + i0, v);
#pragma unroll
            for (int p = 0; p < P; ++p) xs[p][d] = v[p];
        }
        T res[P];
        bool ok[P];
        bool all_ok = true;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            T t[N];
            I base;
            ok[p] = linear_locate_any<T, N, RECT, I, (WL != 0)>(a, axes, xs[p], t, base);
            all_ok = all_ok && ok[p];
            if (!ok[p]) base = 0;  // keep the gather in range; the value is discarded
            res[p] = linear_tree<T, N, WL, I>(a.vals, a.win, base, stride, t);
        }
        if (all_ok) {
            store_result_vec<T, P>(a.out + i0, res);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (ok[p]) store_result(a.out + i0 + p, res[p]);
                else report_bad(a, i0 + p);
            }
        }
    }
    if constexpr (P > 1) {
        const unsigned long long i = ngroups * P + gtid;
        if (i < a.n) {
            T xs[N];
#pragma unroll
            for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
            T t[N];
            I base;
            if (linear_locate_any<T, N, RECT, I, (WL != 0)>(a, axes, xs, t, base)) {
                store_result(a.out + i, linear_tree<T, N, WL, I>(a.vals, a.win, base, stride, t));
            } else {
                report_bad(a, i);
            }
        }
    }
}

template <class T, int N, bool RECT, int WL, int P, class I>
__global__ void __launch_bounds__(kBlock) linear_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    if constexpr (RECT) {
        if (a.axes_in_smem) linear_body<T, N, RECT, WL, P, I, true>(a);
        else linear_body<T, N, RECT, WL, P, I, false>(a);
    } else {
        linear_body<T, N, RECT, WL, P, I, false>(a);
    }
}

// ---------------------------------------------------------------------------------------------
// Multilinear N = 4..6 on a grid beyond L2 (C3-linear, C4): the hypercube layout,
//   hwin[f*16 + v] = vals[f + sum_b bit_b(v) * stride_{N-4+b}],
// — the 2^4 corners of the cell at flat index f over the LAST FOUR dimensions as one aligned 128-byte block (f64; 64 bytes
// in f32), a 16-fold copy of the grid (C3: 2.1 GB, C4: 24.5 GB; 180 GB of HBM is what it is for). HBM serves random aligned
// 128-byte lines at 39 G/s = 5 TB/s when a line is ONE request — four lanes, one LDG.256 each, in one instruction — against
// 50 G/s for 32-byte sectors requested separately (profiles/r2_microbench_b200.json). A point therefore costs 2^(N-4) line
// requests (the corners of the leading dimensions) instead of 2^(N-1) row gathers: no sort, no slab passes, and the
// gather no longer sits on the L1 wavefront rate. (One thread loading its block with four instructions was measured
// at 10.3 G points/s on C3-linear: the four sectors travel as four requests.)
// Work split: thread i owns point i (coalesced coordinate loads, cell location, the last two lerp levels, coalesced store);
// the quad works through its four points: lane j loads sector j of each of the point's blocks — bits (0,1) of v, i.e.
// dimensions N-4 and N-3, inside the sector; bits (2,3) = j, dimensions N-2 and N-1 — reduces the leading dimensions
// between the blocks (dimension 0 first) and then dimensions N-4, N-3 inside the sector, with the owner's t (shuffles); the
// four partial results reach the owner through the skewed transposition buffer of cubic_quad4.cuh, and the owner finishes
// with dimensions N-2 and N-1. Every lerp and their order are the reference's (multilinear/regular.rs:362-388).
// ---------------------------------------------------------------------------------------------
constexpr int kHyperXposeQuad = 20;
template <class T>
__host__ __device__ constexpr size_t linear_hyper_smem_bytes() {
    return static_cast<size_t>(kBlock / 32) * 8 * kHyperXposeQuad * sizeof(T);
}

template <class T, int N, bool RECT, bool AXSM>
__device__ __forceinline__ void linear_hyper_body(const EvalArgs<T, N>& a) {
    static_assert(N >= 4 && N <= 6, "the hypercube layout covers N = 4..6");
    using O = Ops<T>;
    constexpr int L = N - 4;  // leading dimensions: 2^L blocks per point
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* axes = nullptr;
    size_t axes_bytes = 0;
    if constexpr (RECT) {
        axes = stage_axes_as<AXSM, T, N>(a);
        if constexpr (AXSM) axes_bytes = (static_cast<size_t>(a.axes_total) * sizeof(T) + 15) / 16 * 16;
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, j = lane & 3u, quad = lane >> 2;
    const int qb = static_cast<int>(lane & ~3u);
    T* xq = reinterpret_cast<T*>(smem_raw + axes_bytes) + (warp * 8 + quad) * kHyperXposeQuad;
    const unsigned long long nblocks = (a.n + blockDim.x - 1) / blockDim.x;
    for (unsigned long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const unsigned long long i = blk * blockDim.x + threadIdx.x;
        const bool valid = i < a.n;
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
        T t[N];
        int base;
        const bool ok = linear_locate_any<T, N, RECT, int, true>(a, axes, xs, t, base);
        if (!ok) base = 0;  // keep the gather in range; the value is discarded
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int bp = __shfl_sync(0xffffffffu, base, qb + p);
            T tp[L + 2];  // the owner's t of the leading dimensions and of the two in-sector dimensions
#pragma unroll
            for (int d = 0; d < L + 2; ++d) tp[d] = __shfl_sync(0xffffffffu, t[d], qb + p);
            T v[1 << L][4];
#pragma unroll
            for (int c = 0; c < (1 << L); ++c) {
                int off = 0;
#pragma unroll
                for (int d = 0; d < L; ++d) off += ((c >> d) & 1) ? a.istride[d] : 0;
                load_row<T, 4, true, long long>(nullptr, a.win, static_cast<long long>(bp + off) * 4 + j, v[c]);
            }
            // leading dimensions between the blocks, dimension 0 (bit 0 of c) first
#pragma unroll
            for (int d = 0; d < L; ++d) {
#pragma unroll
                for (int c = 0; c < ((1 << L) >> (d + 1)); ++c) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[c][e] = muladd(tp[d], O::sub(v[2 * c + 1][e], v[2 * c][e]), v[2 * c][e]);
                }
            }
            const T a0 = muladd(tp[L], O::sub(v[0][1], v[0][0]), v[0][0]);
            const T a1 = muladd(tp[L], O::sub(v[0][3], v[0][2]), v[0][2]);
            xq[j * 5 + p] = muladd(tp[L + 1], O::sub(a1, a0), a0);
        }
        __syncwarp();
        const T w0 = xq[j], w1 = xq[5 + j], w2 = xq[10 + j], w3 = xq[15 + j];  // this lane's point: the four sectors' results
        __syncwarp();  // the next iteration overwrites the buffer
        const T b0 = muladd(t[N - 2], O::sub(w1, w0), w0);
        const T b1 = muladd(t[N - 2], O::sub(w3, w2), w2);
        const T res = muladd(t[N - 1], O::sub(b1, b0), b0);
        if (valid) {
            if (ok) store_result(a.out + i, res);
            else report_bad(a, i);
        }
    }
}

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) linear_hyper_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    if constexpr (RECT) {
        if (a.axes_in_smem) linear_hyper_body<T, N, RECT, true>(a);
        else linear_hyper_body<T, N, RECT, false>(a);
    } else {
        linear_hyper_body<T, N, RECT, false>(a);
    }
}

// N = 3: the block is the whole 2^3 footprint (64 bytes in f64, an 8-fold copy), hwin[f*8 + v] with bit b of v = offset along
// dimension b, read by a PAIR of lanes (sector s = bit 2 = dimension 2; dimensions 0, 1 inside the sector). A quad handles its
// four points in two steps of two points; the owner finishes with dimension 2.
template <class T, bool RECT, bool AXSM>
__device__ __forceinline__ void linear_hyper3_body(const EvalArgs<T, 3>& a) {
    using O = Ops<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* axes = nullptr;
    size_t axes_bytes = 0;
    if constexpr (RECT) {
        axes = stage_axes_as<AXSM, T, 3>(a);
        if constexpr (AXSM) axes_bytes = (static_cast<size_t>(a.axes_total) * sizeof(T) + 15) / 16 * 16;
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, j = lane & 3u, quad = lane >> 2;
    const int qb = static_cast<int>(lane & ~3u);
    const int sct = static_cast<int>(j & 1u), half = static_cast<int>(j >> 1);
    T* xq = reinterpret_cast<T*>(smem_raw + axes_bytes) + (warp * 8 + quad) * kHyperXposeQuad;
    const unsigned long long nblocks = (a.n + blockDim.x - 1) / blockDim.x;
    for (unsigned long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const unsigned long long i = blk * blockDim.x + threadIdx.x;
        const bool valid = i < a.n;
        T xs[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) xs[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
        T t[3];
        int base;
        const bool ok = linear_locate_any<T, 3, RECT, int, true>(a, axes, xs, t, base);
        if (!ok) base = 0;  // keep the gather in range; the value is discarded
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = 2 * u + half;  // this lane pair's point of the step
            const int bp = __shfl_sync(0xffffffffu, base, qb + p);
            const T t0 = __shfl_sync(0xffffffffu, t[0], qb + p), t1 = __shfl_sync(0xffffffffu, t[1], qb + p);
            T v[4];
            load_row<T, 4, true, long long>(nullptr, a.win, static_cast<long long>(bp) * 2 + sct, v);
            const T a0 = muladd(t0, O::sub(v[1], v[0]), v[0]);
            const T a1 = muladd(t0, O::sub(v[3], v[2]), v[2]);
            xq[sct * 5 + p] = muladd(t1, O::sub(a1, a0), a0);
        }
        __syncwarp();
        const T w0 = xq[j], w1 = xq[5 + j];  // this lane's point: the results of its two sectors
        __syncwarp();  // the next iteration overwrites the buffer
        const T res = muladd(t[2], O::sub(w1, w0), w0);
        if (valid) {
            if (ok) store_result(a.out + i, res);
            else report_bad(a, i);
        }
    }
}

template <class T, bool RECT>
__global__ void __launch_bounds__(kBlock) linear_hyper3_kernel(const __grid_constant__ EvalArgs<T, 3> a) {
    if constexpr (RECT) {
        if (a.axes_in_smem) linear_hyper3_body<T, RECT, true>(a);
        else linear_hyper3_body<T, RECT, false>(a);
    } else {
        linear_hyper3_body<T, RECT, false>(a);
    }
}

// Slab passes — multilinear straight from `vals` on a grid a little beyond L2 (C3: 134 MB against 126 MB), whose
// footprints are too few rows for the bin-swept path to pay. The batch is evaluated in a few launches; each takes only
// the points whose dimension-0 coordinate lies between two nodes of axis 0, so that all of a launch's gathers fall in
// one slab of `vals` that IS L2-resident (every footprint row is otherwise a 32-byte DRAM sector fetch, 417 B/point
// against 41 algorithmic: profiles/r1_p2_c3_linear4d_rect64_ncu.json).
// Skipping inside a warp would save nothing (the lanes that skip wait for the ones that gather), so each warp first
// sifts a tile of kSlabTile points — one coalesced load of coordinate 0 and two compares per point — into a dense list
// in shared memory, then evaluates the list 32 points at a time with the ordinary locate + lerp tree.
// The sift only has to PARTITION the batch among the launches (every x belongs to exactly one: e_lo < x <= e_hi, the
// first launch also takes x <= e_lo and NaN, the last everything above); which cell the point really falls in is
// decided by the ordinary locate, so results and error reports are those of linear_kernel bit for bit.
#ifndef IB200_SLAB_TILE
#define IB200_SLAB_TILE 256
#endif
constexpr int kSlabTile = IB200_SLAB_TILE;
#ifndef IB200_SLAB_ROWS
#define IB200_SLAB_ROWS 0  // 0: two scalar loads per row, 1: load_row_aligned (measured 3 % slower on C3-linear)
#endif
#ifndef IB200_SLAB_CELL
#define IB200_SLAB_CELL 1  // rectilinear axes: 1 = cell tables (three shared-memory loads per dimension), 0 = bucket search
#endif
#ifndef IB200_SLAB_MINB
#define IB200_SLAB_MINB 4
#endif
template <class T, int N, bool RECT, class I>
__global__ void __launch_bounds__(kBlock, IB200_SLAB_MINB) linear_slab_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    using O = Ops<T>;
    const I(&stride)[N] = strides_of<I>(a);
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    __shared__ unsigned short s_list[kBlock / 32][kSlabTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool first = a.slab_lo == 0, last = a.slab_hi < 0;
    T e_lo, e_hi;
    if constexpr (RECT) {
        e_lo = axes[a.axis_off[0] + a.slab_lo];
        e_hi = axes[a.axis_off[0] + (last ? 0 : a.slab_hi)];
    } else {
        e_lo = O::add(a.start[0], O::mul(a.step[0], O::from_int(a.slab_lo)));
        e_hi = O::add(a.start[0], O::mul(a.step[0], O::from_int(last ? 0 : a.slab_hi)));
    }
    unsigned short* list = s_list[warp];
    const unsigned long long ntiles = (a.n + kSlabTile - 1) / kSlabTile;
    const unsigned long long nwarps = static_cast<unsigned long long>(gridDim.x) * (kBlock / 32);
    for (unsigned long long tile = static_cast<unsigned long long>(blockIdx.x) * (kBlock / 32) + warp; tile < ntiles; tile += nwarps) {
        const unsigned long long i0 = tile * kSlabTile;
        const T* x0 = a.obs[0] + i0;
        const int m = static_cast<int>(min(static_cast<unsigned long long>(kSlabTile), a.n - i0));
        int count = 0;
#pragma unroll 4
        for (int k = 0; k < kSlabTile / 32; ++k) {
            const int j = k * 32 + lane;
            bool mine = false;
            if (j < m) {
                const T x = __ldg(x0 + j);
                mine = (first || x > e_lo) && (last || !(x > e_hi));
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, mine);
            if (mine) list[count + __popc(ballot & ((1u << lane) - 1u))] = static_cast<unsigned short>(j);
            count += __popc(ballot);
        }
        __syncwarp();
        for (int q = lane; q < count; q += 32) {
            const unsigned long long i = i0 + list[q];
            T xs[N];
#pragma unroll
            for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
            T t[N];
            I base;
            if (linear_locate_any<T, N, RECT, I, IB200_SLAB_CELL != 0>(a, axes, xs, t, base)) {
                store_result(a.out + i, linear_tree<T, N, IB200_SLAB_ROWS, I>(a.vals, a.win, base, stride, t));
            } else {
                report_bad(a, i);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Nearest (ref: nearest/regular.rs:234-295, nearest/rectilinear.rs:193-241)
// ---------------------------------------------------------------------------------------------

template <class T, int N, bool RECT, int P, class I, bool AXSM>
__device__ __forceinline__ void nearest_body(const EvalArgs<T, N>& a) {
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes_as<AXSM, T, N>(a);
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long gtid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned long long ngroups = a.n / P;
    // The coordinates of the thread's NEXT group are requested before the current group is located and gathered, so their
    // DRAM latency overlaps the search / gather / store chain (ncu on 2-D rectilinear f32: 17 long-scoreboard stall cycles
    // per issue at 33 % issue activity — each iteration began with a full DRAM round trip, profiles/r2_c5_n2rect_f32_ncu.json).
    // Measured (gpurun_out/r2_exp8, G points/s with / without): regular 2-D 198 / 194 (f32 229 / 221), 3-D 173 / 161 (f32 214 /
    // 208), rectilinear f32 2-D 96.6 / 87.7, 3-D 126.6 / 122.8 — but rectilinear f64 114 / 128 and 93 / 104 (the extra
    // registers cost it a resident CTA), which therefore keeps the plain loop.
    constexpr bool kPrefetch = !(RECT && sizeof(T) == 8);
    T nx[N][P];
    if (kPrefetch && gtid < ngroups) {
#pragma unroll
        for (int d = 0; d < N; ++d) load_query_vec<T, P>(a.obs[d] + gtid * P, nx[d]);
    }
    for (unsigned long long g = gtid; g < ngroups; g += gstride) {
        const unsigned long long i0 = g * P;
        T xs[P][N];
        if constexpr (kPrefetch) {
#pragma unroll
            for (int d = 0; d < N; ++d) {
#pragma unroll
                for (int p = 0; p < P; ++p) xs[p][d] = nx[d][p];
            }
            if (g + gstride < ngroups) {
#pragma unroll
                for (int d = 0; d < N; ++d) load_query_vec<T, P>(a.obs[d] + (g + gstride) * P, nx[d]);
            }
        } else {
#pragma unroll
            for (int d = 0; d < N; ++d) {
                T v[P];
                load_query_vec<T, P>(a.obs[d] + i0, v);
#pragma unroll
                for (int p = 0; p < P; ++p) xs[p][d] = v[p];
            }
        }
        I idx[P];
        bool ok[P];
        bool all_ok = true;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ok[p] = nearest_locate_any<T, N, RECT, I>(a, axes, xs[p], idx[p]);
            all_ok = all_ok && ok[p];
            if (!ok[p]) idx[p] = 0;
        }
        T res[P];
#pragma unroll
        for (int p = 0; p < P; ++p) res[p] = __ldg(a.vals + idx[p]);
        if (all_ok) {
            store_result_vec<T, P>(a.out + i0, res);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (ok[p]) store_result(a.out + i0 + p, res[p]);
                else report_bad(a, i0 + p);
            }
        }
    }
    if constexpr (P > 1) {
        const unsigned long long i = ngroups * P + gtid;
        if (i < a.n) {
            T xs[N];
#pragma unroll
            for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + i);
            I idx;
            if (nearest_locate_any<T, N, RECT, I>(a, axes, xs, idx)) store_result(a.out + i, __ldg(a.vals + idx));
            else report_bad(a, i);
        }
    }
}

template <class T, int N, bool RECT, int P, class I>
__global__ void __launch_bounds__(kBlock) nearest_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    if constexpr (RECT) {
        if (a.axes_in_smem) nearest_body<T, N, RECT, P, I, true>(a);
        else nearest_body<T, N, RECT, P, I, false>(a);
    } else {
        nearest_body<T, N, RECT, P, I, false>(a);
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

// One 1-D cubic step. The five-way saturation switch of the reference is evaluated without
// divergent branches: both centred slopes are always formed, the operands of the active formula
// are selected, and the arithmetic that follows is the reference's own sequence for that case, so
// every lane's result is bit-identical to its branch. `all_none` is warp-uniform (every lane of
// the warp is in the interior on this dimension) and skips the selects.
template <class T>
__device__ __forceinline__ T cubic_regular_step(T v0, T v1, T v2, T v3, const CubicRegDim<T>& c, bool all_none) {
    using O = Ops<T>;
    const T half = T(0.5);  // `/ two` is exact and equals `* 0.5` bit-for-bit
    const T two = T(2);
    const T sa = O::mul(O::sub(v2, v0), half);  // (v2 - v0) / 2
    const T sb = O::mul(O::sub(v3, v1), half);  // (v3 - v1) / 2
    if (all_none) return hermite(c.tt, v1, O::sub(v2, v1), sa, sb);
    const bool low = c.mode == kModeLow, high = c.mode == kModeHigh;
    const T y0 = high ? v2 : v1;
    const T y1 = low ? v0 : (high ? v3 : v2);
    const T k0 = low ? -sa : (high ? sb : sa);
    const T dy = O::sub(y1, y0);
    const T knat = O::sub(O::mul(two, dy), k0);  // natural-spline end condition
    const T k1 = (low || high) ? knat : sb;
    const T cub = hermite(c.tt, y0, dy, k0, k1);
    const T lin = muladd(k1, c.ttm1, y1);
    return c.lin ? lin : cub;
}

template <class T>
__device__ __forceinline__ bool cubic_regular_locate(T x, T start, T step, T rstep, bool fast, int dim, int linearize,
                                                     int& origin, CubicRegDim<T>& c) {
    using O = Ops<T>;
    // f = floor((x - start)/step); the reference's iloc is f - 1 (multicubic/regular.rs:438-440).
    int f = 0;
    bool ok = floor_cell(x, start, step, rstep, fast, f);
    origin = min(max(f, 1) - 1, dim - 4);  // clamp(iloc, 0, dim - 4) without overflow at the saturation ends
    bool outside;
    if (f < 0) { c.mode = kModeLow; outside = true; }              // iloc < -1
    else if (f == 0) { c.mode = kModeLow; outside = false; }       // iloc == -1
    else if (f > dim - 2) { c.mode = kModeHigh; outside = true; }  // iloc > n - 3
    else if (f == dim - 2) { c.mode = kModeHigh; outside = false; }
    else { c.mode = kModeNone; outside = false; }
    // t is relative to footprint index 1 and its origin coordinate is never fused
    // (ref: multicubic/regular.rs:356-360).
    T x1 = O::add(start, O::mul(step, O::from_int(origin + 1)));
    T t = exact_div(O::sub(x, x1), step, rstep, fast);
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
    T rdiv0;   // RN(1 / div0), for exact_div
    T wa1, wc1, div1, rdiv1;  // same for k1 (interior cells only)
    int mode;
    bool lin;
    bool fast;  // div0 and div1 are in exact_div's exponent range
};

template <class T>
__device__ __forceinline__ void cubic_rect_locate(T x, const T* __restrict__ g, int pp, int dim, int linearize, int& origin,
                                                  CubicRectDim<T>& c) {  // pp = partition_point(g < x)
    using O = Ops<T>;
    const T one = T(1);
    const int n = dim;
    const int iloc = pp - 2;
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
    c.rdiv0 = O::div(one, c.div0);
    c.rdiv1 = O::div(one, c.div1);
    c.fast = exact_div_divisor_ok(c.div0) && exact_div_divisor_ok(c.div1);
    c.ttm1 = O::sub(c.tt, one);
    c.lin = outside && linearize;
}

// centered_difference_nonuniform(y0,y1,y2,h01,h12) = a*b + c*d with the unit-spacing divisions
// (x / 1.0, exact) dropped (ref: multicubic/mod.rs:103-117, rectilinear.rs:449-450). Like the
// regular-grid step this is select-based: the one data-dependent quotient of k0 picks its
// numerator by case; k1's quotient exists only in the interior case.
template <bool RECURSIVE = false, class T>
__device__ __forceinline__ T cubic_rect_step(T v0, T v1, T v2, T v3, const CubicRectDim<T>& c, bool all_none) {
    using O = Ops<T>;
    const T two = T(2);
    const T d10 = O::sub(v1, v0), d21 = O::sub(v2, v1), d32 = O::sub(v3, v2);
    if (all_none) {
        // k0: h01 = r, h12 = 1 -> b = (v2-v1)/1, d = (v1-v0)/r;  k1: h01 = 1, h12 = s -> b = (v3-v2)/s, d = (v2-v1)/1
        T k0 = cdn_sum(c.wa, d21, c.wc, exact_div(d10, c.div0, c.rdiv0, c.fast));
        T k1 = cdn_sum(c.wa1, exact_div(d32, c.div1, c.rdiv1, c.fast), c.wc1, d21);
        return hermite(c.tt, v1, d21, k0, k1);
    }
    const bool low = c.mode == kModeLow, high = c.mode == kModeHigh, none = !(low || high);
    const T q0 = exact_div(none ? d10 : d21, c.div0, c.rdiv0, c.fast);
    // None: wa*(v2-v1) + wc*((v1-v0)/r)   Low: -(wa*((v2-v1)/q) + wc*(v1-v0))   High: wa*(v3-v2) + wc*((v2-v1)/p)
    const T pa = none ? d21 : (low ? q0 : d32);
    const T pc = low ? d10 : q0;
    const T k0r = cdn_sum(c.wa, pa, c.wc, pc);
    const T k0 = low ? -k0r : k0r;
    const T y0 = high ? v2 : v1;
    const T y1 = low ? v0 : (high ? v3 : v2);
    const T dy = none ? d21 : (low ? O::sub(v0, v1) : d32);  // not -d10: equal neighbours must give +0, like the reference
    T k1 = O::sub(O::mul(two, dy), k0);
    if (none) k1 = cdn_sum(c.wa1, exact_div(d32, c.div1, c.rdiv1, c.fast), c.wc1, d21);
    const T cub = hermite(c.tt, y0, dy, k0, k1);
    const T lin = muladd<RECURSIVE>(k1, c.ttm1, y1);  // fused by the recursive twin only (rectilinear_recursive.rs)
    return c.lin ? lin : cub;
}

template <class T, bool RECT>
struct CubicDimOf {
    using type = CubicRegDim<T>;
};
template <class T>
struct CubicDimOf<T, true> {
    using type = CubicRectDim<T>;
};

// RECURSIVE: called from the looping tree of N >= 5, the reference's recursive structs (fma-flavour quirks).
template <class T, bool RECT, bool RECURSIVE = false>
__device__ __forceinline__ T cubic_step(T v0, T v1, T v2, T v3, const typename CubicDimOf<T, RECT>::type& c,
                                        bool all_none) {
    if constexpr (RECT) return cubic_rect_step<RECURSIVE>(v0, v1, v2, v3, c, all_none);
    else return cubic_regular_step(v0, v1, v2, v3, c, all_none);
}

// Fully unrolled 4^N footprint (N <= 4, the reference's "flattened" range), row by row: reduces
// dimensions 0..D-1 of the sub-block at `idx` for all four positions of the last (contiguous)
// dimension at once, dimension 0 innermost like the reference (multicubic/regular.rs:383-412).
template <int D, class T, int N, bool RECT, bool WIN>
__device__ __forceinline__ void cubic_rows(const T* __restrict__ vals, const T* __restrict__ win, long long idx,
                                           const long long (&stride)[N],
                                           const typename CubicDimOf<T, RECT>::type (&c)[N], unsigned none_mask,
                                           T (&out)[4]) {
    if constexpr (D == 0) {
        load_row<T, 4, WIN>(vals, win, idx, out);
    } else {
        T sub[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            cubic_rows<D - 1, T, N, RECT, WIN>(vals, win, idx + k * stride[D - 1], stride, c, none_mask, sub[k]);
        const bool all_none = (none_mask >> (D - 1)) & 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            out[j] = cubic_step<T, RECT>(sub[0][j], sub[1][j], sub[2][j], sub[3][j], c[D - 1], all_none);
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
                const T(&s)[4] = store[j - 1];
                store[j][slot] = cubic_step<T, RECT, true>(s[0], s[1], s[2], s[3], c[j - 1], false);
            }
        }
    }
    const T(&s)[4] = store[N - 1];
    return cubic_step<T, RECT, true>(s[0], s[1], s[2], s[3], c[N - 1], false);
}

// One query point of the cubic evaluation. Must be called by all 32 lanes of a warp (the
// per-dimension "every lane is interior" votes are warp-wide); lanes without a point pass a copy of
// a valid one and discard the result. Returns false for an unrepresentable coordinate.
template <class T, int N, bool RECT, bool WIN>
__device__ __forceinline__ bool cubic_point(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&xs)[N], T& res) {
    typename CubicDimOf<T, RECT>::type c[N];
    long long base = 0;
    bool ok = true;
    unsigned none_mask = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const T x = xs[d];
        int origin;
        if constexpr (RECT) {
            cubic_rect_locate(x, axes + a.axis_off[d], rect_lower_bound<T, N>(a, axes, d, x), a.dim[d], a.linearize, origin, c[d]);
        } else {
            ok = cubic_regular_locate(x, a.start[d], a.step[d], a.rstep[d], a.fast_div != 0, a.dim[d], a.linearize, origin,
                                      c[d]) && ok;
        }
        base += static_cast<long long>(origin) * a.stride[d];
        if constexpr (N <= 4) none_mask |= __all_sync(0xffffffffu, c[d].mode == kModeNone) ? (1u << d) : 0u;
    }
    if (!ok) return false;
    if constexpr (N <= 4) {
        T r[4];
        cubic_rows<N - 1, T, N, RECT, WIN>(a.vals, a.win, base, a.stride, c, none_mask, r);
        res = cubic_step<T, RECT>(r[0], r[1], r[2], r[3], c[N - 1], (none_mask >> (N - 1)) & 1u);
    } else {
        res = cubic_tree_loop<T, N, RECT>(a.vals + base, a.stride, c);
    }
    return true;
}

// MINB = CTAs per SM the register allocation must allow (launch_common.cuh cubic_min_blocks).
template <class T, int N, bool RECT, bool WIN, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) cubic_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    // Whole warps iterate together (lanes past the end are masked) so the votes are warp-uniform.
    BlockSchedule sched;
    while (sched.next(a.work, a.n)) {
        const unsigned long long i = sched.blk * blockDim.x + threadIdx.x;
        const bool valid = i < a.n;
        const unsigned long long il = valid ? i : a.n - 1;
        T xs[N];
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = load_query(a.obs[d] + il);
        T res;
        const bool ok = cubic_point<T, N, RECT, WIN>(a, axes, xs, res);
        if (valid) {
            if (ok) store_result(a.out + i, res);
            else report_bad(a, i);
        }
    }
}

}  // namespace ib200
