// cubic_quad4.cuh — multicubic N = 2..4, third generation: a quad of four lanes evaluates FOUR query points per
// iteration over the COEFFICIENT layout of the grid.
//
// First-level hoisting. The reference reduces a 4^N footprint dimension 0 first (multicubic/regular.rs:368-412: vertex
// i's offset along dimension k is bit field k of i), so 4^(N-1) of the (4^N - 1)/3 one-dimensional steps of a point —
// 16 of 21 in 3-D, 64 of 85 in 4-D — take four RAW grid values along dimension 0. Everything such a step computes before
// its Horner polynomial (the differences, the centred / non-uniform slopes with their two IEEE quotients on rectilinear
// axes, the natural-spline end slope, the coefficients c1, c2, c3 of normalized_hermite_spline, multicubic/mod.rs:72-91)
// depends on the grid values and on the CELL of dimension 0 only, not on t. It is therefore computed once per
// (cell of dimension 0, node of the other dimensions) when the interpolator is built — by cubic_coef_perm below, the
// same IEEE operations in the same order as the step itself — and stored as one 32-byte sector (y0, c1, c2, c3):
//   cwin[(slot * S0 + r) * 4 + 0..3],  r = flat index over dimensions 1..N-1,  S0 = stride of dimension 0,
//   slot 1 + c for cell c = clamp(floor index, 0, dim0 - 2) (cell 0 = low end class, dim0 - 2 = high end class),
//   slot 0 / dim0 = (y1, k1, 0, 0) of the linearized extrapolation below / above the grid.
// A first-level step is then ONE sector load and three multiply-adds (6 FP64 instructions; 3 in the fma flavour)
// instead of four loads and 14 (regular) or 30 (rectilinear) instructions, bit for bit the reference's result. The
// table is four times the grid (x (dim0 + 1)/dim0), the size of the window copy of round 1 that it replaces.
//
// Work split (unchanged from the second generation, profiles/r1_p5_c2_quad4_ncu.json):
//   * thread i OWNS point i: it loads the point's coordinates (coalesced), locates it on every dimension with the
//     division-free test of device_math.cuh (the FMA remainder proves floor((x-start)/step); rare points take the
//     IEEE division), evaluates the point's LAST 1-D step and stores the result (coalesced);
//   * in between, the quad works through its four points one after the other: lane j reduces dimensions 0..N-2 at
//     last-dimension node j — sectors of consecutive last-dimension nodes are adjacent, so a quad's load is 128
//     contiguous bytes;
//   * parameters (t per dimension, flat index, saturation flags) and the 4x4 transposition of the partial results
//     go through padded shared memory (two LDS.128 per point; conflict-free strides);
//   * the saturation switch of the run-time steps is a PERMUTATION of the four inputs of a 1-D step — InsideLow /
//     OutsideLow is the interior formula on (v2, v1, v0), InsideHigh / OutsideHigh on (v1, v2, v3), with the natural
//     end slope k1 = 2 dy - k0 (multicubic/regular.rs:519-613). Every dimension's four inputs now come from four
//     different loads (dimensions 1..N-2) or four different lanes (dimension N-1), so the permutation is applied to
//     the load ADDRESS / lane index and costs nothing per step; dimension 0's class is the table slot.
// The run-time 1-D steps are the reference's operation sequence with three exactly-equivalent fusions (cubic_step_perm).
#pragma once
#include "interp_internal.h"
#include "kernels.cuh"

namespace ib200 {

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
// Elements per row of the cell table (capi.cu cubic_cell_table; interp_internal.h cubic_cell_row_stride): 12 used, f64
// rows padded to 14 so that the 16-byte chunks of eight random rows fall into eight different bank groups.
template <class T>
constexpr int kCubicCellRow = cubic_cell_row_stride(static_cast<int>(sizeof(T)));

// What the owner of a point publishes to its quad. flags, four bits per dimension d: bits 4d..4d+1 = CubicMode,
// bit 4d+2 = linearized extrapolation applies, bit 4d+3 = the spacing ratios are within exact_div's range (rectilinear).
template <class T, int N, bool RECT>
struct QuadSlot {
    T tt[N];
    int pp[RECT ? N : 1];  // rectilinear: partition_point(g < x) per dimension = row of the cell table
    int base;              // sector index of the footprint's first corner in the coefficient layout
    int flags;
};

// Per-warp parameter block in shared memory, a structure of arrays so that both sides are conflict-free
// (profiles/r2_c2_coef_ncu.json: with one 32-byte struct per lane the owner's 64-bit accesses took 4 wavefronts instead
// of 2): the owner of point e (= its lane) accesses element e of every array — consecutive lanes, consecutive words —
// and the four lanes of a quad read element 4q + p. Arrays of 64-bit elements leave one pad element between the two
// half-warps (a 64-bit access is served per half-warp; without it quads q and q+4 meet in the same banks).
template <class T, int N, bool RECT>
struct QuadParams {
    static constexpr int kTStride = 34;                 // elements per T array: 32 + the pad (+1: keeps 16-byte multiples)
    static constexpr int kInts = 2 + (RECT ? N : 0);    // base, flags, pp[N]
    static constexpr int kBytes = (N * kTStride * static_cast<int>(sizeof(T)) + kInts * 32 * 4 + 15) / 16 * 16;
    unsigned char* w;
    __device__ __forceinline__ T* tt(int d) const { return reinterpret_cast<T*>(w) + d * kTStride; }
    __device__ __forceinline__ int* ints(int k) const { return reinterpret_cast<int*>(w + N * kTStride * sizeof(T)) + k * 32; }
    static __device__ __forceinline__ int pad(int e) { return sizeof(T) == 8 ? e + (e >> 4) : e; }
    __device__ __forceinline__ void publish(int e, const QuadSlot<T, N, RECT>& m) const {
#pragma unroll
        for (int d = 0; d < N; ++d) tt(d)[pad(e)] = m.tt[d];
        ints(0)[e] = m.base;
        ints(1)[e] = m.flags;
        if constexpr (RECT) {
#pragma unroll
            for (int d = 0; d < N; ++d) ints(2 + d)[e] = m.pp[d];
        }
    }
};
// One point's parameters as seen by a reader.
template <class T, int N, bool RECT>
struct QuadRef {
    QuadParams<T, N, RECT> P;
    int e;
    __device__ __forceinline__ T tt(int d) const { return P.tt(d)[P.pad(e)]; }
    __device__ __forceinline__ int base() const { return P.ints(0)[e]; }
    __device__ __forceinline__ int flags() const { return P.ints(1)[e]; }
    __device__ __forceinline__ int pp(int d) const { return P.ints(2 + d)[e]; }
};

template <class T, int N>
__device__ __forceinline__ QuadDim<T, false> quad4_dim(const EvalArgs<T, N>&, const T*, const QuadRef<T, N, false>& r, int d) {
    return QuadDim<T, false>{r.tt(d)};
}
template <class T, int N>
__device__ __forceinline__ QuadDim<T, true> quad4_dim(const EvalArgs<T, N>& a, const T* __restrict__ axes,
                                                      const QuadRef<T, N, true>& r, int d) {
    const T* row = axes + a.ct_off[d] + r.pp(d) * kCubicCellRow<T>;
    QuadDim<T, true> c;
    c.tt = r.tt(d);
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

// |v| in [2^-300, 2^301) — exact_div's operand range, tested on the high word — or a zero, whose quotient by a positive
// divisor is the zero itself (every term of the sequence is a zero; the caller's copysign restores -0).
__device__ __forceinline__ bool exact_div_operand_ok(double v) {
    const unsigned h = static_cast<unsigned>(__double2hiint(v)) & 0x7fffffffu;
    return (h - (723u << 20) < (601u << 20)) || (h | static_cast<unsigned>(__double2loint(v))) == 0u;
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
        // q1 feeds the interior slope only: in an end cell its numerator is u3 - u2 with u3 a stand-in (permuted order
        // 1,2,3,3 makes it exactly zero at the high end), so it must not send the lane to the IEEE division — that was
        // 17 out-of-line divisions per 32 points on a batch with 10 % of the points outside the grid
        // (profiles/r2_x4rect_coef_ncu.json). The spacing ratios are positive, so copysign only matters for a zero numerator.
        const bool end = !all_none && (fl & 3) != 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            q0[j] = copysign(markstein_div_raw(d10[j], c.div0, c.rdiv0), d10[j]);
            q1[j] = copysign(markstein_div_raw(d32[j], c.div1, c.rdiv1), d32[j]);
            ok = ok && exact_div_operand_ok(d10[j]) && (end || exact_div_operand_ok(d32[j]));
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

// ---------------------------------------------------------------------------------------------
// First-level coefficients (see the header). CubicCoef of a step on PERMUTED inputs: (y0, c1, c2, c3) of its Horner
// polynomial y0 + tt*(c1 + tt*(c2 + tt*c3)), and (y1, k1) of the linearized extrapolation y1 + k1*(tt - 1) of an end
// cell. Same operations in the same order as cubic_step_perm / cubic_step_tail up to the polynomial.
// ---------------------------------------------------------------------------------------------
template <class T>
struct CubicCoef {
    T y0, c1, c2, c3, y1, k1;
};

template <class T>
__device__ __forceinline__ CubicCoef<T> cubic_coef_perm(T u0, T u1, T u2, T u3, const QuadDim<T, false>&, int mode) {
    using O = Ops<T>;
    const T half = T(0.5), two = T(2);
    CubicCoef<T> r;
    const T dy = O::sub(u2, u1);
    T a, b;
    if (mode == kModeNone) {
        const T d20 = O::sub(u2, u0);
        a = O::fma(half, d20, -dy);
        const T d31 = O::sub(u3, u1);
        b = O::fma(-half, d31, dy);
        r.k1 = O::mul(d31, half);  // interior cells are never extrapolated from: not stored
    } else {
        const T d20 = neg_zero_if(O::sub(u2, u0), mode == kModeLow);  // low end: the reference's -(v2 - v0)
        a = O::fma(half, d20, -dy);
        const T k0 = O::mul(d20, half);
        r.k1 = O::fma(two, dy, -k0);
        b = O::sub(dy, r.k1);
    }
    r.y0 = u1;
    r.y1 = u2;
    r.c1 = O::add(dy, a);
    r.c2 = O::fma(-two, a, b);
    r.c3 = O::sub(a, b);
    return r;
}

// Rectilinear: the two quotients are plain IEEE divisions here (what exact_div returns by construction).
template <class T>
__device__ __forceinline__ CubicCoef<T> cubic_coef_perm(T u0, T u1, T u2, T u3, const QuadDim<T, true>& c, int mode) {
    using O = Ops<T>;
    const T d10 = O::sub(u1, u0), dy = O::sub(u2, u1), d32 = O::sub(u3, u2);
    const T q0 = O::div(d10, c.div0), q1 = O::div(d32, c.div1);
    T k0 = cdn_sum(c.wa, dy, c.wc, q0);
    T k1 = cdn_sum(c.wa1, q1, c.wc1, dy);
    if (mode != kModeNone) {
        if constexpr (kArithFma) k0 = mode == kModeLow ? cdn_sum(c.wc, q0, c.wa, dy) : k0;
        k0 = neg_zero_if(k0, mode == kModeLow);  // low end: the reference negates the sum
        k1 = O::fma(T(2), dy, -k0);
    }
    CubicCoef<T> r;
    const T a = O::sub(k0, dy);  // hermite_fused
    const T b = O::sub(dy, k1);
    r.y0 = u1;
    r.y1 = u2;
    r.k1 = k1;
    r.c1 = O::add(dy, a);
    r.c2 = O::fma(T(-2), a, b);
    r.c3 = O::sub(a, b);
    return r;
}

// A first-level step from its sector: the Horner polynomial, or — `lin`, the point lies outside the grid on dimension 0
// and linearize_extrapolation is set — y1 + k1*(tt - 1) from the extrapolation slot (fused under the fma feature by the
// regular-grid struct only: multicubic/regular.rs:553-564 against rectilinear.rs:480-540). LIN is a compile-time copy of
// the warp-uniform "some lane of this point group is linearized": the caller branches ONCE per point between the two
// instantiations of the whole reduction (quad4_reduce) — left as a run-time flag around these three instructions, ptxas
// predicated them and every sector paid for both formulas and two selects (ncu: 17 of ~75 instructions per four sectors).
template <bool RECT, bool LIN, class T>
__device__ __forceinline__ T cubic_coef_eval(const T (&s)[4], T tt, bool lin) {
    using O = Ops<T>;
    const T cub = muladd(muladd(muladd(s[3], tt, s[2]), tt, s[1]), tt, s[0]);
    if constexpr (!LIN) {
        return cub;
    } else {
        const T linv = muladd<!RECT>(s[1], O::sub(tt, T(1)), s[0]);
        return lin ? linv : cub;
    }
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
        const bool lin = outside && a.linearize;
        // dimension 0 indexes the coefficient table by slot (header): 1 + cell, or the extrapolation slots 0 / dim
        const int pos = d == 0 ? (lin ? (low ? 0 : dim) : min(max(f, 0), dim - 2) + 1) : origin;
        base += pos * a.istride[d];
        flags |= (mode | (lin ? 4 : 0)) << (4 * d);
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
        const T* row = axes + a.ct_off[d] + pp * kCubicCellRow<T>;
        T gref, href, rhref;
        int rf;
        if constexpr (sizeof(T) == 8) {  // two 16-byte loads (four 8-byte loads of 32 random rows cost 11 wavefronts each)
            const double2 g = reinterpret_cast<const double2*>(row)[4], h = reinterpret_cast<const double2*>(row)[5];
            gref = g.x; href = g.y; rhref = h.x;
            rf = __double2loint(h.y);
        } else {
            const float4 g = reinterpret_cast<const float4*>(row)[2];
            gref = g.x; href = g.y; rhref = g.z;
            rf = __float_as_int(g.w);
        }
        const T e = O::sub(x[d], gref);
        s.tt[d] = exact_div((rf & 3) == kModeLow ? -e : e, href, rhref, (rf & 16) != 0);
        s.pp[d] = pp;
        const bool lin = (rf & 4) && a.linearize;
        // dimension 0 indexes the coefficient table by slot (header): 1 + cell, or the extrapolation slots 0 / dim
        const int pos = d == 0 ? (lin ? ((rf & 3) == kModeLow ? 0 : a.dim[d]) : clamp_cell(pp - 1, a.dim[d] - 2) + 1)
                               : clamp_cell(pp - 2, a.dim[d] - 4);
        base += pos * a.istride[d];
        flags |= ((rf & 3) | (lin ? 4 : 0) | (rf & 8)) << (4 * d);
    }
    s.base = base;
    s.flags = flags;
    return true;
}

// k-th input of a step in permuted order: interior 0,1,2,3; low end 2,1,0,3; high end 1,2,3,3.
__device__ __forceinline__ int cubic_perm_k(int mode, int k) {
    return mode == kModeLow ? (k < 3 ? 2 - k : 3) : (mode == kModeHigh ? min(k + 1, 3) : k);
}

// Reduces dimensions 0..D-1 (D >= 1) of the sub-block whose first-level sector index is `idx`: dimension 0 from the
// coefficient sectors, dimensions 1..D-1 by run-time steps whose four inputs are visited in the permuted order of the
// dimension's saturation class. Per-dimension parameters are read from the owner's slot (and the cell table) where
// they are used, so they are live only for the step of their dimension.
#ifndef IB200_QUAD4_UNROLL_OUTER
#define IB200_QUAD4_UNROLL_OUTER 1  // outermost loop of a 4-D footprint: 1 = rolled (four inner groups of 4 loads), 4 = unrolled
#endif
constexpr int kQuad4UnrollOuter = IB200_QUAD4_UNROLL_OUTER;
template <int D, bool LIN, class T, int N, bool RECT>
__device__ __forceinline__ T quad4_reduce(const EvalArgs<T, N>& a, const T* __restrict__ axes, int idx,
                                          const QuadRef<T, N, RECT>& sp, T tt0, int flags, unsigned none_mask) {
    if constexpr (D == 1) {
        T s[4];
        load_row<T, 4, true, int>(nullptr, a.win, idx, s);
        return cubic_coef_eval<RECT, LIN>(s, tt0, (flags & 4) != 0);
    } else {
        const int fl = flags >> (4 * (D - 1));
        const bool all_none = (none_mask >> (D - 1)) & 1u;
        const int mode = all_none ? 0 : (fl & 3), stride = a.istride[D - 1];
        T u0, u1, u2, u3;
        if constexpr (D >= 3) {
            u0 = u1 = u2 = u3 = T(0);
            // the four sub-results are shifted through u0..u3 so that no register array is indexed dynamically
#pragma unroll(kQuad4UnrollOuter)
            for (int k = 0; k < 4; ++k) {
                const T v = quad4_reduce<D - 1, LIN, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, k) * stride, sp, tt0, flags, none_mask);
                u0 = u1; u1 = u2; u2 = u3; u3 = v;
            }
        } else {
            u0 = quad4_reduce<D - 1, LIN, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, 0) * stride, sp, tt0, flags, none_mask);
            u1 = quad4_reduce<D - 1, LIN, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, 1) * stride, sp, tt0, flags, none_mask);
            u2 = quad4_reduce<D - 1, LIN, T, N, RECT>(a, axes, idx + cubic_perm_k(mode, 2) * stride, sp, tt0, flags, none_mask);
            u3 = quad4_reduce<D - 1, LIN, T, N, RECT>(a, axes, idx + 3 * stride, sp, tt0, flags, none_mask);
        }
        const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, sp, D - 1);
        if (all_none) return cubic_step_perm(u0, u1, u2, u3, c, fl, true);  // one warp-uniform branch
        return cubic_step_perm(u0, u1, u2, u3, c, fl, false);
    }
}

// Measured and dropped (round 2, gpurun_out/r2_exp6): a software pipeline over groups of four first-level sectors — the
// next group's loads issued as soon as the current group's sectors were reduced by their Horner polynomials, before its
// run-time steps — needed 80 (regular) to 126 (rectilinear) registers and changed nothing: C2 30.7 against 30.5,
// 3-D rectilinear 18.0 against 18.3, 4-D regular 8.76 against 8.51, 4-D rectilinear 5.79 against 6.06 G points/s. The
// kernels are bound by the L2 -> L1 sector rate (C2, 4-D regular) or by issue + FP64 (rectilinear), not by the load latency
// of one warp: the other resident warps already cover it.
// Experiment (round 2, north_star's "TMA or cp.async staged coordinate loads"): IB200_QUAD4_TMA=1 replaces the register
// prefetch of the next block's coordinates by a per-warp double-buffered bulk-copy pipeline — one elected lane issues N
// cp.async.bulk (1-D TMA) copies of 32 coordinates into shared memory, completion on an mbarrier with expect_tx; the warp
// waits on the barrier's phase where it used to wait on the loads' scoreboard. Blocks whose coordinates cannot be copied in
// bulk (pointers not 16-byte aligned, the ragged last block) take the plain loads. Bit-identical (the GPU parity file passes
// with this build) and SLOWER everywhere (gpurun_out/r2_tma; profiles/r2_c2_tma_variant_ncu.json): C2 29.7 against 32.5 G
// points/s, 4-D regular 7.03 against 8.55, 3-D rectilinear 16.0 against 18.4. The coordinates were never the problem — 3 of
// the 19 load instructions per 32 points, already requested a block ahead — and the detour adds 28 % to the shared-memory
// wavefronts of a kernel whose binding units are the L1 data pipe and the L2 -> L1 sector rate, plus the barrier traffic.
// Kept as a compile-time variant (default off) so that the measurement can be repeated.
#ifndef IB200_QUAD4_TMA
#define IB200_QUAD4_TMA 0
#endif
constexpr bool kQuad4Tma = IB200_QUAD4_TMA != 0;
template <class T, int N>
__host__ __device__ constexpr int quad4_tma_warp_bytes() {  // [2 buffers][N][32 coordinates] + two mbarriers
    return kQuad4Tma ? 2 * N * 32 * static_cast<int>(sizeof(T)) + 16 : 0;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kQuad4XposeQuad = 20;  // transposition buffer: element (node j, point p) of quad q at q*20 + 5j + p, see the kernel
template <class T, int N, bool RECT>
__host__ __device__ constexpr size_t quad4_smem_bytes() {  // beyond the staged axes
    return static_cast<size_t>(kBlock / 32) * (QuadParams<T, N, RECT>::kBytes + 8 * kQuad4XposeQuad * sizeof(T) + quad4_tma_warp_bytes<T, N>());
}

// AXSM: the rectilinear axes blob (axes, bucket tables, cell tables) is staged in shared memory — the loads of the
// tables are then LDS, not generic loads — else it is read from global memory through L1 (blobs beyond the budget).
template <class T, int N, bool RECT, int MINB, bool AXSM = true>
__global__ void __launch_bounds__(kBlock, MINB) cubic_quad4_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    static_assert(N >= 2 && N <= 4, "quad-cooperative cubic covers N = 2..4");
    using Params = QuadParams<T, N, RECT>;
    constexpr int kWarps = kBlock / 32;
#ifndef IB200_QUAD4_UNROLL3
#define IB200_QUAD4_UNROLL3 4
#endif
    constexpr int kUnrollP = (!RECT && N <= 3) ? IB200_QUAD4_UNROLL3 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* axes = nullptr;
    size_t axes_bytes = 0;
    if constexpr (RECT) {
        if constexpr (AXSM) {
            T* s_axes = reinterpret_cast<T*>(smem_raw);
            for (int k = threadIdx.x; k < a.axes_total; k += blockDim.x) s_axes[k] = a.axes[k];
            __syncthreads();
            axes = s_axes;
            axes_bytes = (static_cast<size_t>(a.axes_total) * sizeof(T) + 15) / 16 * 16;
        } else {
            axes = a.axes;
        }
    }
    unsigned char* s_params = smem_raw + axes_bytes;
    T* s_xpose = reinterpret_cast<T*>(s_params + kWarps * Params::kBytes);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, b = lane & 3u, quad = lane >> 2;
    const Params params{s_params + warp * Params::kBytes};
    const QuadRef<T, N, RECT> mine_ref{params, static_cast<int>(lane)};
    // Transposition buffer of the quad: element (last-dimension node j, point p) at 5j + p, quads 20 elements apart. With
    // this skew both sides are conflict-free for 64-bit elements: the writers of one point (lanes j, fixed p) fall on
    // bank pairs 4q + 5j, the readers of one node (lanes p, fixed j) on 4q + p — all distinct within a half-warp (the
    // former 4j + p layout made every write an 8-way conflict, 43 % of the kernel's shared-memory wavefronts).
    T* xq = s_xpose + (warp * 8 + quad) * kQuad4XposeQuad;

    // The coordinates of the NEXT block of points are requested after the last gather of the current one has been
    // consumed (its registers are free again), so their DRAM latency overlaps the transposition, the final step
    // and the store instead of stalling the next cell location.
    // (IB200_QUAD4_TMA) per-warp bulk-copy pipeline of the coordinates: buffers [2][N][32] and two mbarriers
    T* sx = nullptr;
    unsigned long long* mbar = nullptr;
    bool obs_aligned = false, cur_tma = false;
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    auto tma_issue = [&](int bf, unsigned long long first) {  // the warp's 32 coordinates of every dimension, one elected lane
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(mbar + bf, N * 32 * static_cast<unsigned>(sizeof(T)));
#pragma unroll
            for (int d = 0; d < N; ++d) bulk_copy_g2s(sx + (bf * N + d) * 32, a.obs[d] + first, 32 * static_cast<unsigned>(sizeof(T)), mbar + bf);
        }
    };
    if constexpr (kQuad4Tma) {
        unsigned char* tma_base = reinterpret_cast<unsigned char*>(s_xpose + kWarps * 8 * kQuad4XposeQuad) + warp * quad4_tma_warp_bytes<T, N>();
        sx = reinterpret_cast<T*>(tma_base);
        mbar = reinterpret_cast<unsigned long long*>(tma_base + 2 * N * 32 * sizeof(T));
        if (lane == 0) {
            mbar_init(mbar, 1);
            mbar_init(mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        unsigned long long bits = 0;
#pragma unroll
        for (int d = 0; d < N; ++d) bits |= reinterpret_cast<unsigned long long>(a.obs[d]);
        obs_aligned = (bits & 15ull) == 0ull;
        __syncwarp();
    }
    BlockSchedule sched;
    if (!sched.next(a.work, a.n)) return;
    unsigned long long i = sched.blk * blockDim.x + threadIdx.x;
    bool valid = i < a.n;
    T x[N];
    if (kQuad4Tma && obs_aligned && (i - lane) + 32 <= a.n) {
        tma_issue(0, i - lane);
        cur_tma = true;
    } else {
#pragma unroll
        for (int d = 0; d < N; ++d) x[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
    }
    for (;;) {
        if (kQuad4Tma && cur_tma) {  // warp-uniform
            mbar_wait(mbar + buf, buf ? phase1 : phase0);
            if (buf) phase1 ^= 1u; else phase0 ^= 1u;
#pragma unroll
            for (int d = 0; d < N; ++d) x[d] = sx[(buf * N + d) * 32 + lane];
        }
        bool ok;
        unsigned edges[N];  // d >= 1: lanes (= points) of the warp in an end cell of dimension d; d = 0: linearized on dimension 0
        {
            QuadSlot<T, N, RECT> mine;
            ok = quad4_locate<T, N>(a, axes, x, mine);
            if (!ok) mine.base = 0;  // keep the gathers in range; the result is discarded
            params.publish(static_cast<int>(lane), mine);
            edges[0] = __ballot_sync(0xffffffffu, (mine.flags & 4) != 0);
#pragma unroll
            for (int d = 1; d < N; ++d) edges[d] = __ballot_sync(0xffffffffu, ((mine.flags >> (4 * d)) & 3) != 0);
        }
        __syncwarp();

        // The four points of the quad, one after the other. Each lane's partial result goes straight into the
        // transposition buffer; unrolled only where the body is small (IB200_QUAD4_UNROLL3).
#pragma unroll(kUnrollP)
        for (int p = 0; p < 4; ++p) {
            const QuadRef<T, N, RECT> sp{params, static_cast<int>(lane & ~3u) + p};
            const int flags = sp.flags();
            unsigned none_mask = 0;
#pragma unroll
            for (int d = 1; d < N; ++d) none_mask |= (edges[d] & (0x11111111u << p)) == 0u ? (1u << d) : 0u;
            const bool lin_any = (edges[0] & (0x11111111u << p)) != 0u;  // warp-uniform: one branch per point group
            // (measured, G points/s with / without the second instantiation: C2 32.5 / 30.7, 4-D rectilinear 6.24 / 6.06, but 4-D
            // regular 8.18 / 8.50 — at 90 % of the L2 -> L1 rate it only pays for the larger code — which keeps one copy)
            constexpr bool kSplitLin = N <= 3 || RECT;
            T v;
            if (kSplitLin && !lin_any) v = quad4_reduce<N - 1, false, T, N, RECT>(a, axes, sp.base() + static_cast<int>(b), sp, sp.tt(0), flags, none_mask);
            else v = quad4_reduce<N - 1, true, T, N, RECT>(a, axes, sp.base() + static_cast<int>(b), sp, sp.tt(0), flags, none_mask);
            xq[b * 5 + p] = v;
        }
        const unsigned long long i_cur = i;
        const bool valid_cur = valid;
        const bool more = sched.next(a.work, a.n);
        if (more) {
            i = sched.blk * blockDim.x + threadIdx.x;
            valid = i < a.n;
            if (kQuad4Tma && obs_aligned && (i - lane) + 32 <= a.n) {
                __syncwarp();  // every lane has read its coordinates of the buffer that is refilled now (two iterations ago)
                buf ^= 1;
                tma_issue(buf, i - lane);
                cur_tma = true;
            } else {
                cur_tma = false;
#pragma unroll
                for (int d = 0; d < N; ++d) x[d] = load_query(a.obs[d] + (valid ? i : a.n - 1));
            }
        }
        // Transposition: lane j holds the partial results of its last-dimension node for points 0..3; the owner of
        // point b needs the four nodes' results of point b, in the permuted order of its saturation class.
        __syncwarp();
        const int fl = mine_ref.flags() >> (4 * (N - 1));  // re-read: not kept live across the point loop
        int k4[4];
        cubic_perm(fl & 3, k4);
        const T w0 = xq[k4[0] * 5 + b], w1 = xq[k4[1] * 5 + b], w2 = xq[k4[2] * 5 + b], w3 = xq[15 + b];
        const QuadDim<T, RECT> c = quad4_dim<T, N>(a, axes, mine_ref, N - 1);
        const T res = cubic_step_perm(w0, w1, w2, w3, c, fl, edges[N - 1] == 0u);
        __syncwarp();  // the next iteration overwrites both buffers
        if (valid_cur) {
            if (ok) store_result(a.out + i_cur, res);
            else report_bad(a, i_cur);
        }
        if (!more) break;
    }
}

// ---------------------------------------------------------------------------------------------
// Builder of the coefficient layout (launch_cubic_build.cu): one thread per sector.
// ---------------------------------------------------------------------------------------------
template <class T, bool RECT>
__global__ void __launch_bounds__(kBlock) build_coef_window_kernel(const T* __restrict__ vals, T* __restrict__ cwin, int dim0,
                                                                   unsigned long long s0, const T* __restrict__ table0) {
    const unsigned long long total = static_cast<unsigned long long>(dim0 + 1) * s0;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < total; k += gstride) {
        const int slot = static_cast<int>(k / s0);
        const unsigned long long r = k - static_cast<unsigned long long>(slot) * s0;
        const bool lin = slot == 0 || slot == dim0;
        const int cell = slot == 0 ? 0 : (slot == dim0 ? dim0 - 2 : slot - 1);
        const int mode = cell == 0 ? kModeLow : (cell == dim0 - 2 ? kModeHigh : kModeNone);
        const int origin = min(max(cell, 1) - 1, dim0 - 4);
        T u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = vals[static_cast<unsigned long long>(origin + cubic_perm_k(mode, j)) * s0 + r];
        QuadDim<T, RECT> c{};
        if constexpr (RECT) {
            // row pp = cell + 1 of axis 0's cell table (the in-grid variant of an end cell; the constants are the same)
            const T* row = table0 + static_cast<size_t>(cell + 1) * kCubicCellRow<T>;
            c.wa = row[0]; c.wc = row[1]; c.div0 = row[2]; c.rdiv0 = row[3];
            c.wa1 = row[4]; c.wc1 = row[5]; c.div1 = row[6]; c.rdiv1 = row[7];
        }
        const CubicCoef<T> q = cubic_coef_perm(u[0], u[1], u[2], u[3], c, mode);
        T* dst = cwin + k * 4;
        dst[0] = lin ? q.y1 : q.y0;
        dst[1] = lin ? q.k1 : q.c1;
        dst[2] = lin ? T(0) : q.c2;
        dst[3] = lin ? T(0) : q.c3;
    }
}

}  // namespace ib200
