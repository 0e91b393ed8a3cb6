// device_math.cuh — per-point arithmetic of the InterpN hot path for sm_100a.
//
// Bit-parity rules:
//  * every add / sub / mul / div goes through the round-to-nearest intrinsics (__dadd_rn, ...),
//    which nvcc never contracts into FMAs, so the operation sequence equals the Rust crate built
//    with default features (no `fma`), independent of -fmad;
//  * divisions are true IEEE divisions (__ddiv_rn / __fdiv_rn), never reciprocal-multiplies,
//    except where a provably exact replacement is used (exact_div.cuh);
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
    static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
    static __device__ __forceinline__ double abs(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double from_int(long long i) { return __ll2double_rn(i); }
    static __device__ __forceinline__ long long to_int(double a) { return __double2ll_rz(a); }
};

template <>
struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float floor(float a) { return ::floorf(a); }
    static __device__ __forceinline__ float abs(float a) { return ::fabsf(a); }
    static __device__ __forceinline__ float from_int(long long i) { return __ll2float_rn(i); }
    static __device__ __forceinline__ long long to_int(float a) { return __float2ll_rz(a); }
};

// Saturation class of one dimension (ref: multicubic/mod.rs:59-66), split into the two facts the
// 1-D cubic step needs: which end-cell formula applies, and whether the point is outside the grid.
enum CubicMode : int { kModeNone = 0, kModeLow = 1, kModeHigh = 2 };

// floor((v - start) / step) as a checked i64 (ref: multilinear/regular.rs:415-418; num-traits
// NumCast: Some iff -2^63 <= f < 2^63, NaN -> None).
template <class T>
__device__ __forceinline__ bool floor_cell(T v, T start, T step, long long& iloc) {
    using O = Ops<T>;
    T fl = O::floor(O::div(O::sub(v, start), step));
    if (!(fl >= T(-9223372036854775808.0) && fl < T(9223372036854775808.0))) return false;
    iloc = O::to_int(fl);
    return true;
}

__device__ __forceinline__ int clamp_cell(long long iloc, int dimmax) {
    // iloc.max(0).min(dimmax)  (ref: multilinear/regular.rs:420-422)
    long long l = iloc > 0 ? iloc : 0;
    return static_cast<int>(l < dimmax ? l : dimmax);
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

// ref: multicubic/mod.rs:72-91 (strict arithmetic)
template <class T>
__device__ __forceinline__ T hermite(T t, T y0, T dy, T k0, T k1) {
    using O = Ops<T>;
    T a = O::sub(k0, dy);
    T b = O::add(-k1, dy);
    T c1 = O::add(dy, a);
    T c2 = O::sub(b, O::add(a, a));
    T c3 = O::sub(a, b);
    return O::add(y0, O::mul(t, O::add(c1, O::mul(t, O::add(c2, O::mul(t, c3))))));
}

}  // namespace ib200
