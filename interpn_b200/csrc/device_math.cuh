// device_math.cuh — per-point arithmetic of the InterpN hot path for sm_100a.
//
// Bit-parity rules:
//  * every add / sub / mul / div goes through the round-to-nearest intrinsics (__dadd_rn, ...),
//    which nvcc never contracts into FMAs, so the operation sequence equals the Rust crate built
//    with default features (no `fma`), independent of -fmad;
//  * divisions are true IEEE divisions (__ddiv_rn / __fdiv_rn), never reciprocal-multiplies,
//    except where the provably identical exact_div sequence below replaces them;
//  * the reduction tree runs dimension 0 first and dimension N-1 last
//    (ref: multilinear/regular.rs:362-388, multicubic/regular.rs:383-420).
//
// `ref:` citations are relative to the reference's src/ directory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ib200 {

constexpr int kMaxDims = 8;
constexpr unsigned long long kNoBad = ~0ull;

template <class T>
struct Ops;

template <>
struct Ops<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    // Only where the fused result provably equals the two-operation sequence (exact inner product).
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double from_int(int i) { return __int2double_rn(i); }
    // floor(a) as i32, saturating at INT_MIN / INT_MAX, NaN -> 0 (one F2I.S32.F64.FLOOR)
    static __device__ __forceinline__ int floor_sat(double a) { return __double2int_rd(a); }
};

template <>
struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float floor(float a) { return ::floorf(a); }
    static __device__ __forceinline__ float abs(float a) { return ::fabsf(a); }
    static __device__ __forceinline__ float from_int(int i) { return __int2float_rn(i); }
    static __device__ __forceinline__ int floor_sat(float a) { return __float2int_rd(a); }
};

// Saturation class of one dimension (ref: multicubic/mod.rs:59-66), split into the two facts the
// 1-D cubic step needs: which end-cell formula applies, and whether the point is outside the grid.
enum CubicMode : int { kModeNone = 0, kModeLow = 1, kModeHigh = 2 };

__device__ __forceinline__ int clamp_cell(int iloc, int dimmax) {
    // iloc.max(0).min(dimmax)  (ref: multilinear/regular.rs:420-422)
    return min(max(iloc, 0), dimmax);
}

// Number of axis entries strictly below v == slice::partition_point(|x| *x < v) on an ascending
// axis (ref: multilinear/rectilinear.rs:363). Branch-free; NaN compares false everywhere -> 0.
template <class T>
__device__ __forceinline__ int lower_bound(const T* __restrict__ g, int n, T v) {
    int lo = 0;
    int len = n;
    while (len > 1) {
        int half = len >> 1;
        lo += (g[lo + half - 1] < v) ? half : 0;
        len -= half;
    }
    lo += (g[lo] < v) ? 1 : 0;
    return lo;
}

// ---------------------------------------------------------------------------------------------
// exact_div: a / b, correctly rounded, for a divisor whose correctly rounded reciprocal
// rb = RN(1/b) is already known (grid steps; spacing ratios reused by every 1-D cubic step of one
// dimension). q0 = RN(a*rb) is within 1.5 ulp of a/b; one residual correction
// q1 = RN(q0 + (a - q0*b)*rb) makes it faithful, and by Markstein's theorem (rb = RN(1/b), q1
// faithful, residual exact under FMA) the second correction yields RN(a/b) — the same bits as the
// IEEE division the reference performs — in 5 FP64 instructions instead of the ~13 + MUFU of the
// general division sequence (measured 2.4x the throughput, profiles/r1_microbench_b200.json). The
// theorem needs every intermediate to stay normal, hence the exponent guards: operands outside the
// guarded range (zeros, subnormals, infinities, NaN, huge or tiny magnitudes) take the true
// division. tests/test_numeric_tricks.py re-checks the sequence against `/` on adversarial operand
// pairs on the CPU, and the GPU parity suite compares whole evaluations bit for bit.
// f32 keeps the true division (cheap on the FP32 pipe).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool exact_div_exponent_ok(double v) {
    const unsigned e = (static_cast<unsigned>(__double2hiint(v)) >> 20) & 0x7ffu;
    return e - 723u <= 600u;  // 2^-300 <= |v| < 2^301
}
__device__ __forceinline__ bool exact_div_divisor_ok(double b) { return exact_div_exponent_ok(b); }
__device__ __forceinline__ bool exact_div_divisor_ok(float) { return false; }

// The IEEE division of the rare operands outside the guarded range, out of line: inlined, its ~40 instructions per
// call site made the cubic kernels overflow the instruction cache (ncu: 6 "no instruction" stall cycles per issue).
static __device__ __noinline__ double exact_div_slow(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double exact_div(double a, double b, double rb, bool divisor_ok) {
    const double q0 = __dmul_rn(a, rb);
    const double e0 = __fma_rn(-q0, b, a);
    const double q1 = __fma_rn(e0, rb, q0);
    const double e1 = __fma_rn(-q1, b, a);
    double q = __fma_rn(e1, rb, q1);
    if (!(divisor_ok && exact_div_exponent_ok(a))) q = exact_div_slow(a, b);
    return q;
}
__device__ __forceinline__ float exact_div(float a, float b, float, bool) { return __fdiv_rn(a, b); }

// floor((v - start) / step) with the reference's representability check (ref:
// multilinear/regular.rs:415-418; num-traits NumCast: Some iff -2^63 <= floor(q) < 2^63, NaN -> None;
// both bounds are integers, so the test on q itself is equivalent). The cell index is returned
// saturated to i32: grid dimensions are < 2^31 - 8 (capi.cu), so every clamp and every saturation
// classification computed from the saturated value equals the one computed from the i64.
// `rstep` = RN(1/step) and `fast` (step within exact_div's exponent range) come from the host; the
// quotient is the IEEE quotient either way.
template <class T>
__device__ __forceinline__ bool floor_cell(T v, T start, T step, T rstep, bool fast, int& iloc) {
    using O = Ops<T>;
    const T q = exact_div(O::sub(v, start), step, rstep, fast);
    iloc = O::floor_sat(q);
    return q >= T(-9223372036854775808.0) && q < T(9223372036854775808.0);
}

// ---------------------------------------------------------------------------------------------
// Division-free cell location for regular f64 grids (the streaming kernels are instruction-issue
// bound). All three pieces return the reference's bits or report "not sure", in which
// case the caller re-evaluates the point with the true divisions above. Preconditions checked on the
// host (launch_common.cuh make_args -> fast_div): 2^-300 <= step < 2^301 and dim < 2^30.
//
//  * fast_cell: the approximate quotient q~ = RN(d * RN(1/step)) is within 2^-52 |Q| of Q = d/step, so
//    f~ = floor(q~) can differ from the reference's floor(RN(Q)) only when Q is that close to an
//    integer. For an unclamped f~ the FMA remainder r = RN(d - f~*step) (sign-exact, one rounding)
//    proves the cell: 0 <= r <= step*(1 - 2^-20) implies f~ <= Q <= f~ + 1 - 2^-20, and since
//    |RN(Q) - Q| <= 2^-53 * 2^31, floor(RN(Q)) = f~. A clamped f~ (f~ < 0 or f~ > dim-2, |f~| <= 2^30)
//    needs no proof: the true floor is within 1 of f~, hence on the same side of the clamp, and
//    linear / nearest evaluation uses the clamped origin only. NaN and +-inf never pass
//    (r is NaN, or f~ saturates beyond 2^30), so the caller's exact path reports them.
//  * nearest_upper: the reference picks the upper node iff !(RN(e/step) <= 0.5). RN is monotone and
//    0.5 has an even significand, so RN(y) <= 0.5 <=> y <= 0.5 + 2^-54, i.e. e - step/2 <= step*2^-54.
//    e - step/2 is exact whenever e is within a factor 2 of step/2 (Sterbenz) and far from the
//    threshold otherwise, so the FP comparison decides exactly like the division.
//  * markstein_div: exact_div's sequence without the branch; valid (== RN(a/b)) when
//    markstein_operand_ok(a) (normal range, or zero — the copysign keeps -0/b = -0).
// tests/test_numeric_tricks.py re-checks all three against IEEE division on the CPU.
// ---------------------------------------------------------------------------------------------
struct FastDim {
    double hstep;  // step / 2
    double tau;    // step * 2^-54
    double lim;    // step * (1 - 2^-20)
};

__device__ __forceinline__ bool fast_cell(double x, double start, double step, double rstep, double lim, int dim,
                                          int& origin, double& od, double& d) {
    d = __dsub_rn(x, start);
    const double q = __dmul_rn(d, rstep);
    const int f = __double2int_rd(q);
    origin = min(max(f, 0), dim - 2);
    od = __int2double_rn(origin);
    const double r = __fma_rn(-od, step, d);
    const bool proven = r >= 0.0 && r <= lim;
    const bool sane = static_cast<unsigned>(f) + (1u << 30) <= (1u << 31);
    return origin == f ? proven : sane;
}

__device__ __forceinline__ bool nearest_upper(double e, double hstep, double tau) {
    return !(__dsub_rn(e, hstep) <= tau);
}

__device__ __forceinline__ bool markstein_operand_ok(double a) {
    const unsigned hi = static_cast<unsigned>(__double2hiint(a));
    const unsigned e = (hi >> 20) & 0x7ffu;
    return (e - 723u <= 600u) || ((hi << 1 | static_cast<unsigned>(__double2loint(a))) == 0u);
}

// The bare sequence (no sign fix-up for a zero numerator, no guard): callers test the operands themselves.
__device__ __forceinline__ double markstein_div_raw(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double e0 = __fma_rn(-q0, b, a);
    const double q1 = __fma_rn(e0, rb, q0);
    const double e1 = __fma_rn(-q1, b, a);
    return __fma_rn(e1, rb, q1);
}
__device__ __forceinline__ float markstein_div_raw(float a, float b, float) { return __fdiv_rn(a, b); }

__device__ __forceinline__ double markstein_div(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double e0 = __fma_rn(-q0, b, a);
    const double q1 = __fma_rn(e0, rb, q0);
    const double e1 = __fma_rn(-q1, b, a);
    return copysign(__fma_rn(e1, rb, q1), a);
}

// ---------------------------------------------------------------------------------------------
// Arithmetic flavour. The reference has two builds: default features (every a*b+c is two rounded operations) and
// the `fma` cargo feature (Cargo.toml; the Python wheel is built with it, pyproject.toml:72), which calls mul_add at a
// fixed list of sites. The library is compiled once per flavour (csrc/Makefile: libinterpn_b200.so and
// libinterpn_b200_fma.so, -DIB200_ARITH_FMA=1); muladd() is a*b+c the way the flavour's reference build computes it
// at such a site. FUSE = false marks a site the reference leaves unfused even with the feature (the recursive twins'
// x0, multilinear/regular_recursive.rs:310-313; the flattened rectilinear cubic's linearized extrapolation,
// multicubic/rectilinear.rs:480-540).
// ---------------------------------------------------------------------------------------------
#ifndef IB200_ARITH_FMA
#define IB200_ARITH_FMA 0
#endif
constexpr bool kArithFma = IB200_ARITH_FMA != 0;

template <bool FUSE = true, class T>
__device__ __forceinline__ T muladd(T a, T b, T c) {
    if constexpr (kArithFma && FUSE) return Ops<T>::fma(a, b, c);
    else return Ops<T>::add(Ops<T>::mul(a, b), c);
}
// centered_difference_nonuniform's a*b + c*d (ref: multicubic/mod.rs:111-116; fma feature: mul_add(a, b, c*d))
template <class T>
__device__ __forceinline__ T cdn_sum(T a, T b, T c, T d) {
    return muladd(a, b, Ops<T>::mul(c, d));
}

// ---- f32 twins of the three division-free pieces -------------------------------------------------------------
// Same arguments with the f32 constants: the approximate quotient is within 2^-23 |Q| of Q = d/step, so an unclamped
// f~ is proven by 0 <= fmaf(-f~, step, d) <= step*(1 - 2^-11) when |f~| < 2^12 (|RN(Q) - Q| <= 2^-24 * 2^12), a clamped
// one needs |f~| <= 2^22; 0.5f has an even significand and ulp 2^-24 above it, so RN(y) <= 0.5 <=> y <= 0.5 + 2^-25;
// Markstein's theorem is precision-independent (guard: operands within [2^-60, 2^60), or a zero numerator).
// Host preconditions (launch_common.cuh make_args -> fast_div): 2^-60 <= step < 2^60 and every dim <= 4096.
__device__ __forceinline__ bool fast_cell(float x, float start, float step, float rstep, float lim, int dim, int& origin,
                                          float& od, float& d) {
    d = __fsub_rn(x, start);
    const float q = __fmul_rn(d, rstep);
    const int f = __float2int_rd(q);
    origin = min(max(f, 0), dim - 2);
    od = __int2float_rn(origin);
    const float r = __fmaf_rn(-od, step, d);
    const bool proven = r >= 0.0f && r <= lim;
    const bool sane = static_cast<unsigned>(f) + (1u << 22) <= (1u << 23);
    return origin == f ? proven : sane;
}
__device__ __forceinline__ bool nearest_upper(float e, float hstep, float tau) { return !(__fsub_rn(e, hstep) <= tau); }
__device__ __forceinline__ bool markstein_operand_ok(float a) {
    const unsigned bits = static_cast<unsigned>(__float_as_int(a));
    const unsigned e = (bits >> 23) & 0xffu;
    return (e - 67u <= 119u) || ((bits << 1) == 0u);  // 2^-60 <= |a| < 2^60, or +-0
}
__device__ __forceinline__ float markstein_div(float a, float b, float rb) {
    const float q0 = __fmul_rn(a, rb);
    const float e0 = __fmaf_rn(-q0, b, a);
    const float q1 = __fmaf_rn(e0, rb, q0);
    const float e1 = __fmaf_rn(-q1, b, a);
    return copysignf(__fmaf_rn(e1, rb, q1), a);
}

// ref: multicubic/mod.rs:72-91. c2 = b - (a+a) is issued as fma(-2, a, b): a+a is exact, so the bits are the same
// in both flavours. The polynomial is y0 + t*(c1 + t*(c2 + t*c3)), three chained mul_add under the fma feature.
template <class T>
__device__ __forceinline__ T hermite(T t, T y0, T dy, T k0, T k1) {
    using O = Ops<T>;
    T a = O::sub(k0, dy);
    T b = O::add(-k1, dy);
    T c1 = O::add(dy, a);
    T c2 = O::fma(T(-2), a, b);
    T c3 = O::sub(a, b);
    return muladd(muladd(muladd(c3, t, c2), t, c1), t, y0);
}

}  // namespace ib200
