// cubic_quad4.cuh — multicubic N = 2..4 on a regular grid, second generation of the quad-cooperative kernel
// (kernels.cuh cubic_quad_kernel): a quad of four lanes evaluates FOUR query points per iteration.
//
// Why (profiles/r1_p4_c2_quad_ncu.json): the one-point-per-quad kernel executes 360 warp instructions per 8 points,
// only 130 of them FP64 — every lane repeats the final 1-D step of its quad (3 of 4 wasted), one lane in four
// repeats a cell location, parameters travel by 20 shuffles per point, and the five-way saturation switch costs
// 16 selects per 1-D step whenever one lane of the warp is near an edge. Here
//   * thread i OWNS point i: it loads the point's coordinates (coalesced), locates it on every dimension with the
//     division-free test of device_math.cuh (the FMA remainder proves floor((x-start)/step); rare points take the
//     IEEE division), evaluates the point's LAST 1-D step and stores the result (coalesced);
//   * in between, the quad works through its four points one after the other: lane j gathers the sector of
//     last-dimension node j from the cross-window layout (one LDG.256 brings four consecutive nodes of dimension
//     N-2) and reduces dimensions 0..N-2 on it;
//   * parameters (t per dimension, flat index, saturation flags) and the 4x4 transposition of the partial results
//     go through padded shared memory (two LDS.128 per point instead of 20 shuffles; conflict-free strides);
//   * the saturation switch becomes a PERMUTATION of the four inputs of a 1-D step — InsideLow/OutsideLow is the
//     interior formula on (v2, v1, v0), InsideHigh/OutsideHigh on (v1, v2, v3), with the natural end slope
//     k1 = 2 dy - k0 (multicubic/regular.rs:519-613) — and for dimensions whose four inputs come from four
//     different loads (dimensions 0..N-3) or from four different lanes (dimension N-1) the permutation is applied to
//     the load ADDRESS, which costs nothing per step. Only dimension N-2 (inside a sector) needs selects.
// The 1-D steps are the reference's operation sequence with three exactly-equivalent fusions (cubic_step_perm).
#pragma once
#include "kernels.cuh"

#ifndef IB200_QUAD4_UNROLLK
#define IB200_QUAD4_UNROLLK 1
#endif

namespace ib200 {

constexpr int kQuad4UnrollK = IB200_QUAD4_UNROLLK;  // outer row loop of a 4-D footprint (quad4_rows): 1 = rolled

// Per-dimension parameters of a 1-D step.
template <class T, bool RECT>
struct QuadDim {
    T tt;  // t (interior), -t (low end), t - 1 (high end)
};
// Rectilinear grids: the spacing ratios and centered_difference_nonuniform weights of the active formula, with wa/wc
// exchanged in a low end cell so that one expression serves all classes. They depend on the cell only and come from
// the per-axis table built at construction (capi.cu cubic_cell_table), staged in shared memory with the axes.
template <class T>
struct QuadDim<T, true> {
    T tt, wa, wc, div0, rdiv0, wa1, wc1, div1, rdiv1;
};
constexpr int kCubicCellRow = 12;  // elements per table row (capi.cu cubic_cell_table)

// What the owner of a point publishes to its quad. flags, four bits per dimension d: bits 4d..4d+1 = CubicMode,
// bit 4d+2 = linearized extrapolation applies, bit 4d+3 = the spacing ratios are within exact_div's range (rectilinear).
template <class T, int N, bool RECT>
struct alignas(16) QuadSlot {
    T tt[N];
    int base;  // flat index of the footprint's first corner
    int flags;
};
template <class T, int N>
struct alignas(16) QuadSlot<T, N, true> {
    T tt[N];
    int pp[N];  // partition_point(g < x) per dimension = row of the cell table
    int base;
    int flags;
};

template <class T, int N>
__device__ __forceinline__ QuadDim<T, false> quad4_dim(const EvalArgs<T, N>&, const T*, const QuadSlot<T, N, false>* sp, int d) {
    return QuadDim<T, false>{sp->tt[d]};
}
template <class T, int N>
__device__ __forceinline__ QuadDim<T, true> quad4_dim(const EvalArgs<T, N>& a, const T* __restrict__ axes,
                                                      const QuadSlot<T, N, true>* sp, int d) {
    const T* row = axes + a.ct_off[d] + sp->pp[d] * kCubicCellRow;
    QuadDim<T, true> c;
    c.tt = sp->tt[d];
    if constexpr (sizeof(T) == 8) {
        const double2 r0 = reinterpret_cast<const double2*>(row)[0], r1 = reinterpret_cast<const double2*>(row)[1];
        const double2 r2 = reinterpret_cast<const double2*>(row)[2], r3 = reinterpret_cast<const double2*>(row)[3];
        c.wa = r0.x; c.wc = r0.y; c.div0 = r1.x; c.rdiv0 = r1.y; c.wa1 = r2.x; c.wc1 = r2.y; c.div1 = r3.x; c.rdiv1 = r3.y;
    } else {
        const float4 r0 = reinterpret_cast<const float4*>(row)[0], r1 = reinterpret_cast<const float4*>(row)[1];
        c.wa = r0.x; c.wc = r0.y; c.div0 = r0.z; c.rdiv0 = r0.w; c.wa1 = r1.x; c.wc1 = r1.y; c.div1 = r1.z; c.rdiv1 = r1.w;
    }
    return c;
}

// v, or -0 when `cond` and v is a zero. The permuted low-end formulas produce v0-v2 (resp. a sum of sign-flipped
// terms) where the reference produces -(v2-v0) (resp. minus the sum): equal bit for bit except that a zero comes out
// as +0 instead of the reference's -0. Integer test and select, general (end-cell) path only.
__device__ __forceinline__ double neg_zero_if(double v, bool cond) {
    // one DSETP + one SEL on the high word (the low word of a zero is already 0)
    const int hi = (cond && v == 0.0) ? static_cast<int>(0x80000000u) : __double2hiint(v);
    return __hiloint2double(hi, __double2loint(v));
}
__device__ __forceinline__ float neg_zero_if(float v, bool cond) {
    return (cond && (__float_as_int(v) << 1) == 0) ? __int_as_float(static_cast<int>(0x80000000u)) : v;
}

template <class T>
__device__ __forceinline__ T hermite_fused(T t, T y0, T dy, T k0, T k1) {  // device_math.cuh hermite with c2 fused
    using O = Ops<T>;
    const T a = O::sub(k0, dy);
    const T b = O::sub(dy, k1);
    const T c1 = O::add(dy, a);
    const T c2 = O::fma(T(-2), a, b);
    const T c3 = O::sub(a, b);
    return muladd(muladd(muladd(c3, t, c2), t, c1), t, y0);
}

// One 1-D cubic step on inputs that are already permuted for the saturation class:
//   interior (v0,v1,v2,v3)   low end (v2,v1,v0,*)   high end (v1,v2,v3,*)
// fl bit 0|1 = this lane is in an end cell (k1 is the natural-spline slope), bit 2 = outside the grid with
// linearize_extrapolation; `all_none` is warp-uniform (no lane of the warp is in an end cell).
// Regular grids: operation sequence of multicubic/regular.rs:474-623 + mod.rs:72-91, with these fusions, each of which
// returns the bits of the two-operation original because its inner product is exact (a power-of-two scaling; needs
// the scaled value to stay normal, i.e. grid-value differences within [2^-1021, 2^1023]):
//   a  = (v2-v0)/2 - dy      -> fma(0.5, v2-v0, -dy)
//   b  = -(v3-v1)/2 + dy     -> fma(-0.5, v3-v1, dy)
//   c2 = b - (a+a)           -> fma(-2, a, b)
//   k1 = 2*dy - k0           -> fma(2, dy, -k0)
template <class T>
__device__ __forceinline__ T cubic_step_perm(T u0, T u1, T u2, T u3, const QuadDim<T, false>& c, int fl, bool all_none) {
    using O = Ops<T>;
    const T half = T(0.5), two = T(2);
    const T tt = c.tt;
    const T dy = O::sub(u2, u1);
    if (all_none) {
        const T d20 = O::sub(u2, u0);
        const T a = O::fma(half, d20, -dy);
        const T d31 = O::sub(u3, u1);
        const T b = O::fma(-half, d31, dy);
        const T c1 = O::add(dy, a);
        const T c2 = O::fma(-two, a, b);
        const T c3 = O::sub(a, b);
        return muladd(muladd(muladd(c3, tt, c2), tt, c1), tt, u1);
    }
    const T d20 = neg_zero_if(O::sub(u2, u0), (fl & 3) == kModeLow);  // low end: the reference's -(v2 - v0)
    const T a = O::fma(half, d20, -dy);
    const T k0 = O::mul(d20, half);
    const T knat = O::fma(two, dy, -k0);
    const T kint = O::mul(O::sub(u3, u1), half);
    const T k1 = (fl & 3) ? knat : kint;
    const T b = O::sub(dy, k1);
    const T c1 = O::add(dy, a);
    const T c2 = O::fma(-two, a, b);
    const T c3 = O::sub(a, b);
    const T cub = muladd(muladd(muladd(c3, tt, c2), tt, c1), tt, u1);
    const T linv = muladd(k1, O::sub(tt, T(1)), u2);  // fused under the fma feature (multicubic/regular.rs:553-564)
    return (fl & 4) ? linv : cub;
}

// Rectilinear grids (multicubic/rectilinear.rs:413-545, mod.rs:103-117). On permuted inputs every class has
//   k0 = wa*(u2-u1) + wc*((u1-u0)/div0),   dy = u2-u1,   y0 = u1
// (interior: as written; high end: the same expression on (v1,v2,v3); low end: -(wa*((v2-v1)/q) + wc*(v1-v0)) equals
// wc*(u2-u1) + wa*((u1-u0)/q) on (v2,v1,v0) because negation commutes with every rounding — hence the exchanged
// weights), and k1 is the interior slope wa1*((u3-u2)/div1) + wc1*(u2-u1) or the end slope 2*dy - k0.
// The two quotients q0 = (u1-u0)/div0 and q1 = (u3-u2)/div1 are passed in (cubic_steps_rect).
template <class T>
__device__ __forceinline__ T cubic_step_tail(T u1, T u2, T dy, T q0, T q1, const QuadDim<T, true>& c, int fl, bool all_none) {
    using O = Ops<T>;
    // centered_difference_nonuniform's a*b + c*d (cdn_sum). Under the fma feature the SECOND product is the rounded
    // one, and in a low end cell the exchanged weights exchange the products' roles: the reference's
    // -fma(a, (v2-v1)/q, c*(v1-v0)) is fma(wc, q0, wa*dy) on the permuted inputs.
    T k0s = cdn_sum(c.wa, dy, c.wc, q0);
    const T kint = cdn_sum(c.wa1, q1, c.wc1, dy);
    if (all_none) return hermite_fused(c.tt, u1, dy, k0s, kint);
    if constexpr (kArithFma) k0s = (fl & 3) == kModeLow ? cdn_sum(c.wc, q0, c.wa, dy) : k0s;
    const T k0 = neg_zero_if(k0s, (fl & 3) == kModeLow);  // low end: the reference negates the sum
    const T k1 = (fl & 3) ? O::fma(T(2), dy, -k0) : kint;
    const T cub = hermite_fused(c.tt, u1, dy, k0, k1);
    const T linv = O::add(u2, O::mul(k1, O::sub(c.tt, T(1))));  // not fused by the flattened structs (N <= 4)
    return (fl & 4) ? linv : cub;
}

// |v| in [2^-300, 2^301): exact_div's operand range, tested on the high word (three integer instructions).
__device__ __forceinline__ bool exact_div_operand_ok(double v) {
    return (static_cast<unsigned>(__double2hiint(v)) & 0x7fffffffu) - (723u << 20) < (601u << 20);
}

// M independent 1-D steps of one dimension (same parameters, inputs u[k][j]). f64: all 2M quotients take the
// five-instruction sequence of device_math.cuh exact_div unconditionally, their operand guards are accumulated into
// ONE predicate, and only a lane with an operand outside the guarded range redoes its quotients with the IEEE
// division — one branch per group instead of one per quotient (ncu: the per-quotient guards and their
// BSSY/BRA/BSYNC were a quarter of the instructions of the rectilinear kernel).
template <int M, class T>
__device__ __forceinline__ void cubic_steps_rect(const T (&u)[4][M], const QuadDim<T, true>& c, int fl, bool all_none,
                                                 T (&out)[M]) {
    using O = Ops<T>;
    T d10[M], dy[M], d32[M], q0[M], q1[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        d10[j] = O::sub(u[1][j], u[0][j]);
        dy[j] = O::sub(u[2][j], u[1][j]);
        d32[j] = O::sub(u[3][j], u[2][j]);
    }
    if constexpr (sizeof(T) == 8) {
        bool ok = (fl & 8) != 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            q0[j] = markstein_div_raw(d10[j], c.div0, c.rdiv0);
            q1[j] = markstein_div_raw(d32[j], c.div1, c.rdiv1);
            ok = ok && exact_div_operand_ok(d10[j]) && exact_div_operand_ok(d32[j]);
        }
        if (!ok) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                q0[j] = exact_div_slow(d10[j], c.div0);
                q1[j] = exact_div_slow(d32[j], c.div1);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < M; ++j) {
            q0[j] = O::div(d10[j], c.div0);
            q1[j] = O::div(d32[j], c.div1);
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) out[j] = cubic_step_tail(u[1][j], u[2][j], dy[j], q0[j], q1[j], c, fl, all_none);
}

template <class T>
__device__ __forceinline__ T cubic_step_perm(T u0, T u1, T u2, T u3, const QuadDim<T, true>& c, int fl, bool all_none) {
    const T u[4][1] = {{u0}, {u1}, {u2}, {u3}};
    T out[1];
    cubic_steps_rect<1, T>(u, c, fl, all_none, out);
    return out[0];
}

// Four steps on the rows of a sub-block (regular grids: four independent calls; the compiler interleaves them).
template <class T>
__device__ __forceinline__ void cubic_steps4(const T (&sub)[4][4], const QuadDim<T, false>& c, int fl, bool all_none, T (&out)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = cubic_step_perm(sub[0][j], sub[1][j], sub[2][j], sub[3][j], c, fl, all_none);
}
template <class T>
__device__ __forceinline__ void cubic_steps4(const T (&sub)[4][4], const QuadDim<T, true>& c, int fl, bool all_none, T (&out)[4]) {
    cubic_steps_rect<4, T>(sub, c, fl, all_none, out);
}

// Two steps (the stashed outer rows of a 4-D footprint are reduced two in-sector positions at a time).
template <class T>
__device__ __forceinline__ void cubic_steps2(const T (&u)[4][2], const QuadDim<T, false>& c, int fl, bool all_none, T (&out)[2]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) out[j] = cubic_step_perm(u[0][j], u[1][j], u[2][j], u[3][j], c, fl, all_none);
}
template <class T>
__device__ __forceinline__ void cubic_steps2(const T (&u)[4][2], const QuadDim<T, true>& c, int fl, bool all_none, T (&out)[2]) {
    cubic_steps_rect<2, T>(u, c, fl, all_none, out);
}

// The same step on inputs in natural order: the permutation is done with selects (dimension N-2, whose four inputs
// sit in one sector).
template <class T, bool RECT>
__device__ __forceinline__ T cubic_step_sel(T v0, T v1, T v2, T v3, const QuadDim<T, RECT>& c, int fl, bool all_none) {
    if (all_none) return cubic_step_perm(v0, v1, v2, v3, c, fl, true);
    const bool low = (fl & 3) == kModeLow, high = (fl & 3) == kModeHigh;
    const T u0 = low ? v2 : (high ? v1 : v0);
    const T u1 = high ? v2 : v1;
    const T u2 = low ? v0 : (high ? v3 : v2);
    return cubic_step_perm(u0, u1, u2, v3, c, fl, false);
}

// Row (or lane) order of the permuted inputs: interior 0,1,2,3; low end 2,1,0,3; high end 1,2,3,3.
__device__ __forceinline__ void cubic_perm(int mode, int (&k)[4]) {
    const bool low = mode == kModeLow, high = mode == kModeHigh;
    k[0] = low ? 2 : (high ? 1 : 0);
    k[1] = high ? 2 : 1;
    k[2] = low ? 0 : (high ? 3 : 2);
    k[3] = 3;
}

// The exact twin of the fast location below for the rare coordinate it cannot prove (within 2^-20 of a node from
// below, beyond 2^30 cells, NaN/inf, a numerator outside exact_div's range): IEEE divisions, out of line.
template <class T>
struct Quad4Exact {
    T t;
    int f;
    int ok;
};
template <class T>
static __device__ __noinline__ Quad4Exact<T> quad4_locate_exact(T x, T start, T step, T rstep, bool fast_div, int dim) {
    using O = Ops<T>;
    Quad4Exact<T> r;
    r.ok = floor_cell(x, start, step, rstep, fast_div, r.f);
    const int origin = min(max(r.f, 1) - 1, dim - 4);
    const T x1 = O::add(start, O::mul(step, O::from_int(origin + 1)));
    r.t = O::div(O::sub(x, x1), step);
    return r;
}

// Cell location of one point on every dimension, regular grid (ref: multicubic/regular.rs:432-469 and :356-360):
// f = floor((x - start)/step) (the reference's iloc + 1), footprint origin clamp(f - 1, 0, dim - 4), saturation class
// from f, and t = (x - x1)/step relative to footprint node 1 (x1 = start + step*(origin + 1), never fused).
// f64: f~ = floor(RN(d * RN(1/step))) is proven by the FMA remainder 0 <= d - f~*step <= step*(1 - 2^-20)
// (device_math.cuh fast_cell; the unclamped f is needed here, so the remainder is taken against f~ itself) and t takes
// the Markstein sequence; one accumulated predicate sends the rare unproven coordinate to quad4_locate_exact.
template <class T, int N>
__device__ __forceinline__ bool quad4_locate(const EvalArgs<T, N>& a, const T*, const T (&x)[N], QuadSlot<T, N, false>& s) {
    using O = Ops<T>;
    bool ok = true;
    int base = 0, flags = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int dim = a.dim[d];
        int f = 0;
        T t = T(0);
        bool good = false;
        {
            // |f| bound of the remainder proof: 2^30 in f64, 2^12 in f32 (device_math.cuh, the f32 twins)
            constexpr unsigned kSane = sizeof(T) == 8 ? (1u << 30) : (1u << 12);
            const T dd = O::sub(x[d], a.start[d]);
            f = O::floor_sat(O::mul(dd, a.rstep[d]));
            const T r = O::fma(-O::from_int(f), a.step[d], dd);
            const int origin = min(max(f, 1) - 1, dim - 4);
            const T x1 = O::add(a.start[d], O::mul(a.step[d], O::from_int(origin + 1)));
            const T e = O::sub(x[d], x1);
            t = markstein_div(e, a.step[d], a.rstep[d]);
            good = a.fast_div != 0 && r >= T(0) && r <= static_cast<T>(a.lim[d]) &&
                   static_cast<unsigned>(f) + kSane <= 2u * kSane && markstein_operand_ok(e);
        }
        if (!good) {
            const Quad4Exact<T> ex = quad4_locate_exact<T>(x[d], a.start[d], a.step[d], a.rstep[d], a.fast_div != 0, dim);
            f = ex.f;
            t = ex.t;
            ok = ex.ok && ok;
        }
        const int origin = min(max(f, 1) - 1, dim - 4);
        // tested in the reference's order: f < 0 OutsideLow, f == 0 InsideLow, f > dim-2 OutsideHigh, f == dim-2 InsideHigh
        const bool low = f <= 0, high = f >= dim - 2, outside = f < 0 || f > dim - 2;
        const int mode = low ? kModeLow : (high ? kModeHigh : kModeNone);
        s.tt[d] = low ? -t : (high ? O::sub(t, T(1)) : t);
        base += origin * a.istride[d];
        flags |= (mode | ((outside && a.linearize) ? 4 : 0)) << (4 * d);
    }
    s.base = base;
    s.flags = flags;
    return ok;
}

// Rectilinear grid (ref: multicubic/rectilinear.rs:366-408): never fails, NaN lands in the first cell. The cell's
// class, reference node and width come from the cell table; t = +-(x - gref)/href is the reference's division
// (kernels.cuh cubic_rect_locate) through the tabulated reciprocal (device_math.cuh exact_div).
template <class T, int N>
__device__ __forceinline__ bool quad4_locate(const EvalArgs<T, N>& a, const T* __restrict__ axes, const T (&x)[N],
                                             QuadSlot<T, N, true>& s) {
    using O = Ops<T>;
    int base = 0, flags = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int pp = rect_lower_bound<T, N>(a, axes, d, x[d]);
        const T* row = axes + a.ct_off[d] + pp * kCubicCellRow;
        const T gref = row[8], href = row[9], rhref = row[10];
        int rf;
        if constexpr (sizeof(T) == 8) rf = __double2loint(row[11]);
        else rf = __float_as_int(row[11]);
        const T e = O::sub(x[d], gref);
        s.tt[d] = exact_div((rf & 3) == kModeLow ? -e : e, href, rhref, (rf & 16) != 0);
        s.pp[d] = pp;
        base += clamp_cell(pp - 2, a.dim[d] - 4) * a.istride[d];
        flags |= ((rf & 3) | (((rf & 4) && a.linearize) ? 4 : 0) | (rf & 8)) << (4 * d);
    }
    s.base = base;
    s.flags = flags;
    return true;
}

// k-th input of a step in permuted order: interior 0,1,2,3; low end 2,1,0,3; high end 1,2,3,3.
__device__ __forceinline__ int cubic_perm_k(int mode, int k) {
    return mode == kModeLow ? (k < 3 ? 2 - k : 3) : (mode == kModeHigh ? min(k + 1, 3) : k);
}

// Reduces dimensions 0..D-1 (address dimensions, D <= N-2) of the sub-block at sector index `idx` for the four
// in-sector positions at once; the rows of each dimension are visited in the permuted order of its saturation class.
// The per-dimension parameters are read from the owner's slot (and the cell table) where they are used, so they are
// live only for the steps of their dimension.
// 16-byte vector of two T (f64: double2, f32: two pairs are wasted bandwidth-wise but keep one code path: float4 holds 4).
template <class T> struct Stash2;
template <> struct Stash2<double> { using V = double2; };
template <> struct Stash2<float> { using V = float2; };

#ifndef IB200_QUAD4_STASH
#define IB200_QUAD4_STASH 1
#endif
// N = 4: the partial results of the outer row loop (four rows of dimension 1 x four in-sector positions = 16 values per
// lane) are parked in shared memory instead of registers. With them in registers the 4-D kernels needed 122-128
// registers (2 CTAs per SM, 24 % of the warps: profiles/r1_p5_x4_ncu.json) and sat on load latency; per-thread columns
// of `stash` ([8 vectors of two values][kBlock threads], consecutive threads in consecutive vectors: conflict-free).
constexpr bool kQuad4Stash = IB200_QUAD4_STASH != 0;
#ifndef IB200_QUAD4_INNER2
#define IB200_QUAD4_INNER2 0
#endif
constexpr bool kQuad4Inner2 = IB200_QUAD4_INNER2 != 0;
// Measured and dropped (round 2, gpurun_out/r2_exp3, commit "prefetch variant"): issuing the four sector loads of row
// k+1 before the steps of row k from a second register block (software pipeline, 128 registers with ~150 bytes of
// spills) ran 4-D regular 32^4 at 4.65 instead of 6.49 G points/s and 4-D rectilinear 32^4 at 1.51 instead of 2.37.

template <int D, class T, int N, bool RECT>
__device__ __forceinline__ void quad4_rows(const EvalArgs<T, N>& a, const T* __restrict__ axes, int idx,
                                           const QuadSlot<T, N, RECT>* sp, int flags, unsigned none_mask,
                                           typename Stash2<T>::V* __restrict__ stash, T (&out)[4]) {
    if constexpr (D == 0) {
        load_row<T, 4, true, int>(nullptr, a.win, idx, out);
    } else {
        const int fl = flags >> (4 * (D - 1));
        const bool all_none = (none_mask >> (D - 1)) & 1u;
        const int mode = all_none ? 0 : (fl & 3), stride = a.istride[D - 1];
        if constexpr (D >= 2 && kQuad4Stash) {
            using V = typename Stash2<T>::V;
            // Not unrolled (code size: a 4-D footprint unrolled 16 ways did not fit the instruction cache).
#pragma unroll(kQuad4UnrollK)
            for (int k = 0; k < 4; ++k) {
                T rk[4];
                quad4_rows<D - 1, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, k) * stride, sp, flags, none_mask, stash, rk);
                V lo, hi;
                lo.x = rk[0]; lo.y = rk[1]; hi.x = rk[2]; hi.y = rk[3];
                stash[(2 * k) * kBlock] = lo;
                stash[(2 * k + 1) * kBlock] = hi;
            }
            const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, sp, D - 1);
            // two in-sector positions at a time: half the live inputs and temporaries of four interleaved steps
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                T u[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const V v = stash[(2 * k + h) * kBlock];
                    u[k][0] = v.x; u[k][1] = v.y;
                }
                T o2[2];
                if (all_none) cubic_steps2(u, c, fl, true, o2);
                else cubic_steps2(u, c, fl, false, o2);
                out[2 * h] = o2[0]; out[2 * h + 1] = o2[1];
            }
        } else {
            T sub[4][4];
            if constexpr (D >= 2) {
                // The four partial rows are shifted through `sub` so that no register array is indexed dynamically.
#pragma unroll(kQuad4UnrollK)
                for (int k = 0; k < 4; ++k) {
                    T rk[4];
                    quad4_rows<D - 1, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, k) * stride, sp, flags, none_mask, stash, rk);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sub[0][j] = sub[1][j]; sub[1][j] = sub[2][j]; sub[2][j] = sub[3][j]; sub[3][j] = rk[j];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    quad4_rows<D - 1, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, k) * stride, sp, flags, none_mask, stash, sub[k]);
            }
            const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, sp, D - 1);
            if constexpr (N >= 4 && kQuad4Inner2) {
                // 4-D: two steps at a time (fewer live temporaries; the kernel is bound by registers, not by ILP)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    T u[4][2], o2[2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { u[k][0] = sub[k][2 * h]; u[k][1] = sub[k][2 * h + 1]; }
                    if (all_none) cubic_steps2(u, c, fl, true, o2);
                    else cubic_steps2(u, c, fl, false, o2);
                    out[2 * h] = o2[0]; out[2 * h + 1] = o2[1];
                }
            } else {
                if (all_none) cubic_steps4(sub, c, fl, true, out);  // one warp-uniform branch around the four independent steps
                else cubic_steps4(sub, c, fl, false, out);
            }
        }
    }
}

template <class T, int N, bool RECT>
__host__ __device__ constexpr int quad4_slot_warp_bytes() {  // one slot per lane + 16 bytes of padding per quad (bank spreading)
    return 32 * static_cast<int>(sizeof(QuadSlot<T, N, RECT>)) + 8 * 16;
}
constexpr int kQuad4XposeQuad = 20;  // transposition buffer [quad][lane j][point p]: quad stride 16 + 4 elements
template <class T, int N>
__host__ __device__ constexpr size_t quad4_stash_bytes() {  // eight 2-element vectors per thread, N = 4 only
    return (N >= 4 && kQuad4Stash) ? static_cast<size_t>(kBlock) * 8 * 2 * sizeof(T) : 0;
}
template <class T, int N, bool RECT>
__host__ __device__ constexpr size_t quad4_smem_bytes() {  // beyond the staged axes
    return static_cast<size_t>(kBlock / 32) * (quad4_slot_warp_bytes<T, N, RECT>() + 8 * kQuad4XposeQuad * sizeof(T)) +
           quad4_stash_bytes<T, N>();
}

template <class T, int N, bool RECT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) cubic_quad4_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    static_assert(N >= 2 && N <= 4, "quad-cooperative cubic covers N = 2..4");
    using Slot = QuadSlot<T, N, RECT>;
    constexpr int kWarps = kBlock / 32;
#ifndef IB200_QUAD4_UNROLL3
#define IB200_QUAD4_UNROLL3 4
#endif
    constexpr int kUnrollP = (!RECT && N <= 3) ? IB200_QUAD4_UNROLL3 : 1;
    constexpr int kSlotWarpBytes = quad4_slot_warp_bytes<T, N, RECT>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* axes = nullptr;
    size_t axes_bytes = 0;
    if constexpr (RECT) {
        axes = stage_axes<T, N>(a);
        if (a.axes_in_smem) axes_bytes = (static_cast<size_t>(a.axes_total) * sizeof(T) + 15) / 16 * 16;
    }
    unsigned char* s_slots = smem_raw + axes_bytes;
    T* s_xpose = reinterpret_cast<T*>(s_slots + kWarps * kSlotWarpBytes);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, b = lane & 3u, quad = lane >> 2;
    unsigned char* wslots = s_slots + warp * kSlotWarpBytes + quad * 16;
    Slot* myslot = reinterpret_cast<Slot*>(wslots) + lane;
    const Slot* qslots = reinterpret_cast<const Slot*>(wslots) + (lane & ~3u);
    T* xq = s_xpose + (warp * 8 + quad) * kQuad4XposeQuad;
    // per-thread column of the stash (16-byte aligned: every region before it is a multiple of 16 bytes)
    typename Stash2<T>::V* stash = reinterpret_cast<typename Stash2<T>::V*>(s_xpose + kWarps * 8 * kQuad4XposeQuad) + threadIdx.x;

    // The coordinates of the NEXT block of points are requested after the last gather of the current one has been
    // consumed (its registers are free again), so their DRAM latency overlaps the transposition, the final step
    // and the store instead of stalling the next cell location.
    BlockSchedule sched;
    if (!sched.next(a.work, a.n)) return;
    unsigned long long i = sched.blk * blockDim.x + threadIdx.x;
    bool valid = i < a.n;
    T x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
    for (;;) {
        bool ok;
        unsigned edges[N];  // lanes (= points) of the warp that are in an end cell of dimension d
        {
            Slot mine;
            ok = quad4_locate<T, N>(a, axes, x, mine);
            if (!ok) mine.base = 0;  // keep the gathers in range; the result is discarded
            *myslot = mine;
#pragma unroll
            for (int d = 0; d < N; ++d) edges[d] = __ballot_sync(0xffffffffu, ((mine.flags >> (4 * d)) & 3) != 0);
        }
        __syncwarp();

        // The four points of the quad, one after the other. Each lane's partial result goes straight into the
        // transposition buffer [lane j][point p]; unrolled only where the body is small (IB200_QUAD4_UNROLL).
#pragma unroll(kUnrollP)
        for (int p = 0; p < 4; ++p) {
            const Slot* sp = qslots + p;
            const int flags = sp->flags;
            unsigned none_mask = 0;
#pragma unroll
            for (int d = 0; d < N; ++d) none_mask |= (edges[d] & (0x11111111u << p)) == 0u ? (1u << d) : 0u;
            T r[4];
            quad4_rows<N - 2, T, N, RECT>(a, axes, sp->base + static_cast<int>(b), sp, flags, none_mask, stash, r);
            const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, sp, N - 2);
            xq[b * 4 + p] = cubic_step_sel<T, RECT>(r[0], r[1], r[2], r[3], c, flags >> (4 * (N - 2)), (none_mask >> (N - 2)) & 1u);
        }
        const unsigned long long i_cur = i;
        const bool valid_cur = valid;
        const bool more = sched.next(a.work, a.n);
        if (more) {
            i = sched.blk * blockDim.x + threadIdx.x;
            valid = i < a.n;
#pragma unroll
            for (int d = 0; d < N; ++d) x[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
        }
        // Transposition: lane j holds the partial results of its last-dimension node for points 0..3; the owner of
        // point b needs the four nodes' results of point b, in the permuted order of its saturation class.
        __syncwarp();
        const int fl = myslot->flags >> (4 * (N - 1));  // re-read from the slot: not kept live across the point loop
        int k4[4];
        cubic_perm(fl & 3, k4);
        const T w0 = xq[k4[0] * 4 + b], w1 = xq[k4[1] * 4 + b], w2 = xq[k4[2] * 4 + b], w3 = xq[12 + b];
        const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, myslot, N - 1);
        const T res = cubic_step_perm(w0, w1, w2, w3, c, fl, edges[N - 1] == 0u);
        __syncwarp();  // the next iteration overwrites both buffers
        if (valid_cur) {
            if (ok) store_result(a.out + i_cur, res);
            else report_bad(a, i_cur);
        }
        if (!more) break;
    }
}

}  // namespace ib200
