// interpn_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch C++ restatement of the arithmetic of interpn 0.8.2 (jlogan03/interpn),
// operation for operation, so that the CUDA path can be compared bit-for-bit.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this. The product library (interpn_b200/csrc) never includes it.
//
// PARITY PINNING: the reference is pure Rust and no Rust toolchain exists in the
// build image, so the reference binary itself cannot be run here. The oracle is
// pinned instead by re-running every known-answer / analytic test the reference
// holds for this path (tests/test_oracle_reference_suite.py lists them with
// file:line) in BOTH arithmetic modes (strict = crate default features; fma =
// the `fma` cargo feature used by the Python wheel), and by requiring the
// "flattened" and "recursive" evaluation orders to agree bit-for-bit in strict mode.
// There are NO stored golden vectors in the reference: ULP-level parity
// against the Rust binary is therefore pinned by this restatement only.
//
// Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off  (rustc never contracts a*b+c;
// .cargo/config.toml:1-2 sets target-cpu=x86-64-v3).
//
// All `ref:` citations are relative to /root/reference/src.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace oracle {

constexpr int MAXDIMS = 8;

// Error codes; messages mirror the reference's &'static str literals (see oracle_strerror).
enum Status : int {
    OK = 0,
    DIM_MISMATCH = 1,          // "Dimension mismatch"
    MIN_TWO = 2,               // "All grids must have at least two entries"      (regular linear/nearest)
    MIN_2 = 3,                 // "All grids must have at least 2 entries"        (rectilinear linear/nearest)
    MIN_FOUR = 4,              // "All grids must have at least four entries"     (regular cubic)
    MIN_4 = 5,                 // "All grids must have at least 4 entries"        (rectilinear cubic)
    NOT_MONOTONIC = 6,         // "All grids must be monotonically increasing"
    UNREPRESENTABLE = 7,       // "Unrepresentable coordinate value"
    MAXDIM_8 = 8,              // "Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions."
    MAXDIM_6 = 9,              // "Dimension exceeds maximum (6)."
    LENGTH_MISMATCH = 10,      // "Length mismatch"          (one_dim)
    UNREPRESENTABLE_NUM = 11,  // "Unrepresentable number"   (one_dim)
};

inline const char* strerror(int s) {
    switch (s) {
        case OK: return "";
        case DIM_MISMATCH: return "Dimension mismatch";
        case MIN_TWO: return "All grids must have at least two entries";
        case MIN_2: return "All grids must have at least 2 entries";
        case MIN_FOUR: return "All grids must have at least four entries";
        case MIN_4: return "All grids must have at least 4 entries";
        case NOT_MONOTONIC: return "All grids must be monotonically increasing";
        case UNREPRESENTABLE: return "Unrepresentable coordinate value";
        case MAXDIM_8:
            return "Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions.";
        case MAXDIM_6: return "Dimension exceeds maximum (6).";
        case LENGTH_MISMATCH: return "Length mismatch";
        case UNREPRESENTABLE_NUM: return "Unrepresentable number";
        default: return "unknown";
    }
}

// ---------------------------------------------------------------------------------------------
// Primitive semantics supplied to the reference by third-party crates.
// ---------------------------------------------------------------------------------------------

// num-traits 0.2.19 `<isize as NumCast>::from(float)`: Some iff -2^63 <= f < 2^63 (NaN -> None).
// Call sites: multilinear/regular.rs:418, multicubic/regular.rs:438, nearest/regular.rs:309,
// one_dim/mod.rs:110.
template <class T>
inline bool float_to_isize(T f, int64_t& out) {
    const T lo = static_cast<T>(-9223372036854775808.0);
    const T hi = static_cast<T>(9223372036854775808.0);
    if (f >= lo && f < hi) {
        out = static_cast<int64_t>(f);
        return true;
    }
    return false;
}

// core::slice::partition_point(|x| *x < v) on an ascending axis == lower bound.
// Call sites: multilinear/rectilinear.rs:363, multicubic/rectilinear.rs:377,
// nearest/rectilinear.rs:258, one_dim/mod.rs:158.
template <class T>
inline int64_t partition_point_lt(const T* g, size_t n, T v) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (g[mid] < v) lo = mid + 1; else hi = mid;
    }
    return static_cast<int64_t>(lo);
}

template <bool FMA, class T>
inline T muladd(T a, T b, T c) {  // a*b + c, fused only when the reference's `fma` feature fuses it
    if (FMA) return std::fma(a, b, c);
    return a * b + c;
}

// C-order strides, ref: multilinear/regular.rs:315-326 (dimprod[N-1]=1, dimprod[k]=prod dims[k+1..]).
inline void c_strides(int n, const size_t* dims, size_t* dimprod) {
    size_t acc = 1;
    for (int i = 0; i < n; ++i) {
        if (i > 0) acc *= dims[n - i];
        dimprod[n - i - 1] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Cell location
// ---------------------------------------------------------------------------------------------

// ref: multilinear/regular.rs:414-425, nearest/regular.rs:305-316 (footprint 2)
template <class T>
inline bool regular_loc2(T v, T start, T step, size_t dim, size_t& loc) {
    T floc = std::floor((v - start) / step);
    int64_t iloc;
    if (!float_to_isize(floc, iloc)) return false;
    int64_t n = static_cast<int64_t>(dim);
    int64_t dimmax = n - 2 > 0 ? n - 2 : 0;
    int64_t l = iloc > 0 ? iloc : 0;
    l = l < dimmax ? l : dimmax;
    loc = static_cast<size_t>(l);
    return true;
}

// ref: multilinear/rectilinear.rs:353-370, nearest/rectilinear.rs:248-265
template <class T>
inline size_t rect_loc2(T v, const T* g, size_t dim) {
    int64_t iloc = partition_point_lt(g, dim, v) - 1;
    int64_t n = static_cast<int64_t>(dim);
    int64_t dimmax = n - 2 > 0 ? n - 2 : 0;
    int64_t l = iloc > 0 ? iloc : 0;
    l = l < dimmax ? l : dimmax;
    return static_cast<size_t>(l);
}

// ref: multicubic/mod.rs:59-66
enum Saturation : int { SAT_NONE = 0, SAT_INSIDE_LOW = 1, SAT_OUTSIDE_LOW = 2, SAT_INSIDE_HIGH = 3, SAT_OUTSIDE_HIGH = 4 };

// ref: multicubic/regular.rs:432-469 (twin regular_recursive.rs:386-423)
template <class T>
inline bool cubic_regular_loc(T v, T start, T step, size_t dim, size_t& loc, Saturation& sat) {
    T floc = std::floor((v - start) / step);
    int64_t iloc;
    if (!float_to_isize(floc, iloc)) return false;
    iloc -= 1;  // i64::MIN - 1 would panic under overflow-checks in the reference; unreachable for finite grids
    int64_t n = static_cast<int64_t>(dim);
    int64_t dimmax = n - 4 > 0 ? n - 4 : 0;
    int64_t l = iloc > 0 ? iloc : 0;
    l = l < dimmax ? l : dimmax;
    loc = static_cast<size_t>(l);
    if (iloc < -1) sat = SAT_OUTSIDE_LOW;
    else if (iloc == -1) sat = SAT_INSIDE_LOW;
    else if (iloc > n - 3) sat = SAT_OUTSIDE_HIGH;
    else if (iloc == n - 3) sat = SAT_INSIDE_HIGH;
    else sat = SAT_NONE;
    return true;
}

// ref: multicubic/rectilinear.rs:366-408 (twin rectilinear_recursive.rs:294-336)
template <class T>
inline void cubic_rect_loc(T v, const T* g, size_t dim, size_t& loc, Saturation& sat) {
    int64_t iloc = partition_point_lt(g, dim, v) - 2;
    int64_t n = static_cast<int64_t>(dim);
    int64_t dimmax = n - 4 > 0 ? n - 4 : 0;
    int64_t l = iloc > 0 ? iloc : 0;
    l = l < dimmax ? l : dimmax;
    loc = static_cast<size_t>(l);
    if (iloc == -2) sat = SAT_OUTSIDE_LOW;
    else if (iloc == -1) sat = SAT_INSIDE_LOW;
    else if (iloc == n - 2) sat = SAT_OUTSIDE_HIGH;
    else if (iloc == n - 3) sat = SAT_INSIDE_HIGH;
    else sat = SAT_NONE;
}

// ---------------------------------------------------------------------------------------------
// Cubic 1-D building blocks
// ---------------------------------------------------------------------------------------------

// ref: multicubic/mod.rs:72-91
template <bool FMA, class T>
inline T hermite(T t, T y0, T dy, T k0, T k1) {
    T a = k0 - dy;
    T b = -k1 + dy;
    T c1 = dy + a;
    T c2 = b - (a + a);
    T c3 = a - b;
    if (FMA) return std::fma(std::fma(std::fma(c3, t, c2), t, c1), t, y0);
    return y0 + t * (c1 + t * (c2 + t * c3));
}

// ref: multicubic/mod.rs:103-117
template <bool FMA, class T>
inline T centered_difference_nonuniform(T y0, T y1, T y2, T h01, T h12) {
    T a = h01 / (h01 + h12);
    T b = (y2 - y1) / h12;
    T c = h12 / (h12 + h01);
    T d = (y1 - y0) / h01;
    if (FMA) return std::fma(a, b, c * d);
    return a * b + c * d;
}

// Regular-grid 1-D cubic step.
// ref: multicubic/regular.rs:474-623 (flattened) and regular_recursive.rs:470-610 (recursive).
// FMA-mode quirk: the recursive twin computes OutsideLow's k1 unfused (regular_recursive.rs:536).
template <bool FMA, bool RECURSIVE, class T>
inline T cubic_regular_inner(const T* v, T t, Saturation sat, bool linearize) {
    const T one = T(1);
    const T two = one + one;
    switch (sat) {
        case SAT_NONE: {
            T y0 = v[1];
            T dy = v[2] - v[1];
            T k0 = (v[2] - v[0]) / two;
            T k1 = (v[3] - v[1]) / two;
            return hermite<FMA>(t, y0, dy, k0, k1);
        }
        case SAT_INSIDE_LOW: {
            T tt = -t;
            T y0 = v[1];
            T dy = v[0] - v[1];
            T k0 = -(v[2] - v[0]) / two;
            T k1 = FMA ? std::fma(two, dy, -k0) : two * dy - k0;
            return hermite<FMA>(tt, y0, dy, k0, k1);
        }
        case SAT_OUTSIDE_LOW: {
            T tt = -t;
            T y0 = v[1];
            T y1 = v[0];
            T dy = v[0] - v[1];
            T k0 = -(v[2] - v[0]) / two;
            T k1 = (FMA && !RECURSIVE) ? std::fma(two, dy, -k0) : two * dy - k0;
            if (linearize) return muladd<FMA>(k1, tt - one, y1);
            return hermite<FMA>(tt, y0, dy, k0, k1);
        }
        case SAT_INSIDE_HIGH: {
            T tt = t - one;
            T y0 = v[2];
            T dy = v[3] - v[2];
            T k0 = (v[3] - v[1]) / two;
            T k1 = FMA ? std::fma(two, dy, -k0) : two * dy - k0;
            return hermite<FMA>(tt, y0, dy, k0, k1);
        }
        default: {  // SAT_OUTSIDE_HIGH
            T tt = t - one;
            T y0 = v[2];
            T y1 = v[3];
            T dy = v[3] - v[2];
            T k0 = (v[3] - v[1]) / two;
            T k1 = FMA ? std::fma(two, dy, -k0) : two * dy - k0;
            if (linearize) return muladd<FMA>(k1, tt - one, y1);
            return hermite<FMA>(tt, y0, dy, k0, k1);
        }
    }
}

// Rectilinear-grid 1-D cubic step.
// ref: multicubic/rectilinear.rs:413-545 (flattened: no fma sites of its own) and
// rectilinear_recursive.rs:385-539 (recursive: fma on k1 and on the linearized extrapolation).
// hermite / centered_difference_nonuniform carry their own fma sites in both.
template <bool FMA, bool RECURSIVE, class T>
inline T cubic_rect_inner(const T* v, const T* g, T x, Saturation sat, bool linearize) {
    const T one = T(1);
    const T two = one + one;
    constexpr bool F = FMA && RECURSIVE;
    switch (sat) {
        case SAT_NONE: {
            T y0 = v[1];
            T dy = v[2] - v[1];
            T h01 = g[1] - g[0];
            T h12 = g[2] - g[1];
            T h23 = g[3] - g[2];
            T k0 = centered_difference_nonuniform<FMA>(v[0], v[1], v[2], h01 / h12, one);
            T k1 = centered_difference_nonuniform<FMA>(v[1], v[2], v[3], one, h23 / h12);
            T t = (x - g[1]) / h12;
            return hermite<FMA>(t, y0, dy, k0, k1);
        }
        case SAT_INSIDE_LOW:
        case SAT_OUTSIDE_LOW: {
            T y0 = v[1];
            T y1 = v[0];
            T dy = v[0] - v[1];
            T h01 = g[1] - g[0];
            T h12 = g[2] - g[1];
            T k0 = -centered_difference_nonuniform<FMA>(v[0], v[1], v[2], one, h12 / h01);
            T k1 = F ? std::fma(two, dy, -k0) : two * dy - k0;
            T t = -(x - g[1]) / h01;
            if (sat == SAT_OUTSIDE_LOW && linearize) return muladd<F>(k1, t - one, y1);
            return hermite<FMA>(t, y0, dy, k0, k1);
        }
        default: {  // SAT_INSIDE_HIGH / SAT_OUTSIDE_HIGH
            T y0 = v[2];
            T y1 = v[3];
            T dy = v[3] - v[2];
            T h12 = g[2] - g[1];
            T h23 = g[3] - g[2];
            T k0 = centered_difference_nonuniform<FMA>(v[1], v[2], v[3], h12 / h23, one);
            T k1 = F ? std::fma(two, dy, -k0) : two * dy - k0;
            T t = (x - g[2]) / h23;
            if (sat == SAT_OUTSIDE_HIGH && linearize) return muladd<F>(k1, t - one, y1);
            return hermite<FMA>(t, y0, dy, k0, k1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// N-D reduction trees.
//
// "Flattened" order (ref: multilinear/regular.rs:347-403, multicubic/regular.rs:368-420):
// vertex v in [0, FP^N); offset along dim k is base-FP digit k of v; whenever a group of FP
// completes at level j it is reduced along dim j-1; final reduce along dim N-1.
// "Recursive" order (ref: multilinear/regular_recursive.rs:348-389,
// multicubic/regular_recursive.rs:427-465): populate(dim) reduces dim-1, leaves at dim 0.
// Both visit the same arithmetic DAG; they are kept separate so the tests can assert that.
// ---------------------------------------------------------------------------------------------

// NC > 0: the dimensionality is a compile-time constant (the reference's structs are const-generic and unroll these loops
// with crunchy::unroll!, multilinear/regular.rs:347-403), so the timed CPU baseline is not slowed by index arithmetic
// the crate never executes. Same operations on the same values in the same order as NC = 0; tests compare the two.
constexpr size_t ipow(size_t b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

template <int FP, int NC = 0, class T, class Reduce>
inline __attribute__((always_inline)) T tree_flattened(int n_rt, const size_t* origin, const size_t* dimprod, const T* vals, Reduce reduce) {
    const int n = NC ? NC : n_rt;
    T store[MAXDIMS][FP];
    size_t nverts = 1;
    for (int i = 0; i < n; ++i) nverts *= FP;
    if (NC) nverts = ipow(FP, NC);
    // fully unrolled up to 64 vertices; a 4-D cubic footprint (256) is unrolled by 64 — the dimension-3 bookkeeping stays
    // in the loop, everything below it is constant (keeps the build of this checker to a minute)
#pragma GCC unroll 64
    for (size_t i = 0; i < nverts; ++i) {
        size_t idx = 0, rem = i;
        for (int k = 0; k < n; ++k) {
            idx += (origin[k] + rem % FP) * dimprod[k];
            rem /= FP;
        }
        store[0][i % FP] = vals[idx];
        size_t q = 1;
        for (int j = 1; j < n; ++j) {
            q *= FP;
            if ((i + 1) % q == 0) {
                size_t p = ((i + 1) / q - 1) % FP;
                store[j][p] = reduce(store[j - 1], j - 1);
            }
        }
    }
    return reduce(store[n - 1], n - 1);
}

template <int FP, class T, class Reduce>
inline T tree_recursive(int dim, const size_t* origin, size_t* loc, const size_t* dimprod, int n, const T* vals,
                        Reduce& reduce) {
    if (dim == 0) {
        size_t idx = 0;
        for (int j = 0; j < n; ++j) idx += loc[j] * dimprod[j];
        return vals[idx];
    }
    int next = dim - 1;
    T v[FP];
    for (int i = 0; i < FP; ++i) {
        loc[next] = origin[next] + i;
        v[i] = tree_recursive<FP>(next, origin, loc, dimprod, n, vals, reduce);
    }
    loc[next] = origin[next];
    return reduce(v, next);
}

template <int FP, bool RECURSIVE, int NC = 0, class T, class Reduce>
inline __attribute__((always_inline)) T tree(int n, const size_t* origin, const size_t* dimprod, const T* vals, Reduce reduce) {
    if (RECURSIVE) {
        size_t loc[MAXDIMS];
        for (int i = 0; i < n; ++i) loc[i] = origin[i];
        return tree_recursive<FP>(n, origin, loc, dimprod, n, vals, reduce);
    }
    return tree_flattened<FP, NC>(n, origin, dimprod, vals, reduce);
}

// ---------------------------------------------------------------------------------------------
// Per-point evaluators (one per reference interp_one)
// ---------------------------------------------------------------------------------------------

// ref: multilinear/regular.rs:296-404; recursive twin regular_recursive.rs:274-323
// (the twin never fuses index_zero_loc, regular_recursive.rs:310-313).
template <bool FMA, bool RECURSIVE, int NC = 0, class T>
inline bool linear_regular_one(int n_rt, const size_t* dims, const T* starts, const T* steps, const T* vals, const T* x,
                               T& out) {
    const int n = NC ? NC : n_rt;
    size_t origin[MAXDIMS], dimprod[MAXDIMS];
    T dts[MAXDIMS];
    c_strides(n, dims, dimprod);
    for (int i = 0; i < n; ++i) {
        if (!regular_loc2(x[i], starts[i], steps[i], dims[i], origin[i])) return false;
        T origin_f = static_cast<T>(origin[i]);
        T zero_loc = muladd<(FMA && !RECURSIVE)>(steps[i], origin_f, starts[i]);
        dts[i] = (x[i] - zero_loc) / steps[i];
    }
    out = tree<2, RECURSIVE, NC>(n, origin, dimprod, vals, [&](const T* s, int d) {
        T y0 = s[0];
        T dy = s[1] - y0;
        return muladd<FMA>(dts[d], dy, y0);
    });
    return true;
}

// ref: multilinear/rectilinear.rs:244-346; recursive twin rectilinear_recursive.rs:224-336
template <bool FMA, bool RECURSIVE, int NC = 0, class T>
inline void linear_rect_one(int n_rt, const size_t* dims, const T* const* grids, const T* vals, const T* x, T& out) {
    const int n = NC ? NC : n_rt;
    size_t origin[MAXDIMS], dimprod[MAXDIMS];
    c_strides(n, dims, dimprod);
    for (int i = 0; i < n; ++i) origin[i] = rect_loc2(x[i], grids[i], dims[i]);
    out = tree<2, RECURSIVE, NC>(n, origin, dimprod, vals, [&](const T* s, int d) {
        T x0 = grids[d][origin[d]];
        T x1 = grids[d][origin[d] + 1];
        T step = x1 - x0;
        T t = (x[d] - x0) / step;
        T y0 = s[0];
        T dy = s[1] - y0;
        return muladd<FMA>(t, dy, y0);
    });
}

// ref: multicubic/regular.rs:325-422; recursive twin regular_recursive.rs:322-376
template <bool FMA, bool RECURSIVE, int NC = 0, class T>
inline bool cubic_regular_one(int n_rt, const size_t* dims, const T* starts, const T* steps, const T* vals,
                              bool linearize, const T* x, T& out) {
    const int n = NC ? NC : n_rt;
    size_t origin[MAXDIMS], dimprod[MAXDIMS];
    Saturation sat[MAXDIMS];
    T dts[MAXDIMS];
    c_strides(n, dims, dimprod);
    for (int i = 0; i < n; ++i) {
        if (!cubic_regular_loc(x[i], starts[i], steps[i], dims[i], origin[i], sat[i])) return false;
        T one_loc = starts[i] + steps[i] * static_cast<T>(origin[i] + 1);  // never fused (regular.rs:356-359)
        dts[i] = (x[i] - one_loc) / steps[i];
    }
    out = tree<4, RECURSIVE, NC>(n, origin, dimprod, vals, [&](const T* s, int d) {
        return cubic_regular_inner<FMA, RECURSIVE>(s, dts[d], sat[d], linearize);
    });
    return true;
}

// ref: multicubic/rectilinear.rs:265-356; recursive twin rectilinear_recursive.rs:242-284
template <bool FMA, bool RECURSIVE, int NC = 0, class T>
inline void cubic_rect_one(int n_rt, const size_t* dims, const T* const* grids, const T* vals, bool linearize,
                           const T* x, T& out) {
    const int n = NC ? NC : n_rt;
    size_t origin[MAXDIMS], dimprod[MAXDIMS];
    Saturation sat[MAXDIMS];
    c_strides(n, dims, dimprod);
    for (int i = 0; i < n; ++i) cubic_rect_loc(x[i], grids[i], dims[i], origin[i], sat[i]);
    out = tree<4, RECURSIVE, NC>(n, origin, dimprod, vals, [&](const T* s, int d) {
        return cubic_rect_inner<FMA, RECURSIVE>(s, grids[d] + origin[d], x[d], sat[d], linearize);
    });
}

// ref: nearest/regular.rs:234-295
template <bool FMA, class T>
inline bool nearest_regular_one(int n, const size_t* dims, const T* starts, const T* steps, const T* vals, const T* x,
                                T& out) {
    size_t dimprod[MAXDIMS];
    c_strides(n, dims, dimprod);
    const T two = T(1) + T(1);
    const T half = T(1) / two;
    size_t idx = 0;
    for (int i = 0; i < n; ++i) {
        size_t origin;
        if (!regular_loc2(x[i], starts[i], steps[i], dims[i], origin)) return false;
        T origin_f = static_cast<T>(origin);
        T zero_loc = muladd<FMA>(steps[i], origin_f, starts[i]);
        T dt = (x[i] - zero_loc) / steps[i];
        size_t off = (dt <= half) ? 0 : 1;  // tie -> lower; NaN cannot reach here (loc fails first)
        idx += (origin + off) * dimprod[i];
    }
    out = vals[idx];
    return true;
}

// ref: nearest/rectilinear.rs:193-241
template <class T>
inline void nearest_rect_one(int n, const size_t* dims, const T* const* grids, const T* vals, const T* x, T& out) {
    size_t dimprod[MAXDIMS];
    c_strides(n, dims, dimprod);
    const T two = T(1) + T(1);
    const T half = T(1) / two;
    size_t idx = 0;
    for (int i = 0; i < n; ++i) {
        size_t origin = rect_loc2(x[i], grids[i], dims[i]);
        T x0 = grids[i][origin];
        T x1 = grids[i][origin + 1];
        T step = x1 - x0;
        T dt = (x[i] - x0) / step;
        size_t off = (dt <= half) ? 0 : 1;  // NaN query: comparison false -> origin+1
        idx += (origin + off) * dimprod[i];
    }
    out = vals[idx];
}

// ---------------------------------------------------------------------------------------------
// one_dim (ref: one_dim/mod.rs, linear.rs, hold.rs)
// ---------------------------------------------------------------------------------------------

enum Extrap : int { INSIDE = 0, OUTSIDE_LOW = 1, OUTSIDE_HIGH = 2 };
enum Kind1D : int { LINEAR = 0, LINEAR_HOLD_LAST = 1, LEFT = 2, RIGHT = 3, NEAREST = 4 };

template <class T>
struct GridSample {
    T x0, y0, x1, y1;
    Extrap extrap;
};

// ref: one_dim/mod.rs:85-138 (RegularGrid1D::new/index/at). `stop` is computed once in new().
template <class T>
inline bool regular_1d_at(T start, T stop, T step, const T* vals, size_t nvals, T loc, GridSample<T>& s) {
    Extrap e = INSIDE;
    if (loc > stop) e = OUTSIDE_HIGH;
    else if (loc < start) e = OUTSIDE_LOW;
    T fi = std::floor((loc - start) / step);
    int64_t ii;
    if (!float_to_isize(fi, ii)) return false;
    int64_t hi = static_cast<int64_t>(nvals - 2);
    ii = ii > 0 ? ii : 0;
    ii = ii < hi ? ii : hi;
    size_t i = static_cast<size_t>(ii);
    s.x0 = start + step * static_cast<T>(i);
    s.x1 = s.x0 + step;
    s.y0 = vals[i];
    s.y1 = vals[i + 1];
    s.extrap = e;
    return true;
}

// ref: one_dim/mod.rs:147-187 (RectilinearGrid1D::index/at)
template <class T>
inline void rect_1d_at(const T* grid, const T* vals, size_t n, T loc, GridSample<T>& s) {
    int64_t pp = partition_point_lt(grid, n, loc) - 1;
    pp = pp > 0 ? pp : 0;
    size_t i = static_cast<size_t>(pp);
    if (i > n - 2) i = n - 2;
    Extrap e = INSIDE;
    if (loc < grid[0]) e = OUTSIDE_LOW;
    else if (loc > grid[n - 1]) e = OUTSIDE_HIGH;
    s.x0 = grid[i];
    s.x1 = grid[i + 1];
    s.y0 = vals[i];
    s.y1 = vals[i + 1];
    s.extrap = e;
}

// ref: one_dim/linear.rs:24-37, 58-85; one_dim/hold.rs:23-39, 58-74, 94-107
template <bool FMA, class T>
inline T eval_1d(int kind, const GridSample<T>& s, T loc) {
    switch (kind) {
        case LINEAR: {
            T slope = (s.y1 - s.y0) / (s.x1 - s.x0);
            T dx = loc - s.x0;
            return muladd<FMA>(slope, dx, s.y0);
        }
        case LINEAR_HOLD_LAST: {
            if (s.extrap == OUTSIDE_LOW) return s.y0;
            if (s.extrap == OUTSIDE_HIGH) return s.y1;
            T slope = (s.y1 - s.y0) / (s.x1 - s.x0);
            T dx = loc - s.x0;
            return muladd<FMA>(slope, dx, s.y0);
        }
        case LEFT: return s.extrap == OUTSIDE_HIGH ? s.y1 : s.y0;
        case RIGHT: return s.extrap == OUTSIDE_LOW ? s.y0 : s.y1;
        default: {  // NEAREST: tie -> left
            T dx0 = std::fabs(loc - s.x0);
            T dx1 = std::fabs(loc - s.x1);
            return (dx1 >= dx0) ? s.y0 : s.y1;
        }
    }
}

}  // namespace oracle
