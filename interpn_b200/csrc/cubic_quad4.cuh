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

namespace ib200 {

template <class T, int N>
struct alignas(16) QuadSlot {
    T tt[N];    // per dimension: t (interior), -t (low end), t - 1 (high end)
    int base;   // flat index of the footprint's first corner
    int flags;  // per dimension d: bits 3d..3d+1 = CubicMode, bit 3d+2 = linearized extrapolation applies
};

// One 1-D cubic step on inputs that are already permuted for the saturation class:
//   interior (v0,v1,v2,v3)   low end (v2,v1,v0,*)   high end (v1,v2,v3,*)
// `edge` = this lane is in an end cell (k1 is the natural-spline slope), `lin` = outside the grid with
// linearize_extrapolation; `all_none` is warp-uniform (no lane of the warp has `edge`).
// Operation sequence of multicubic/regular.rs:474-623 + mod.rs:72-91, with these fusions, each of which returns the
// bits of the two-operation original because its inner product is exact (a power-of-two scaling; needs the scaled
// value to stay normal, i.e. grid-value differences within [2^-1021, 2^1023]):
//   a  = (v2-v0)/2 - dy      -> fma(0.5, v2-v0, -dy)
//   b  = -(v3-v1)/2 + dy     -> fma(-0.5, v3-v1, dy)
//   c2 = b - (a+a)           -> fma(-2, a, b)
//   k1 = 2*dy - k0           -> fma(2, dy, -k0)
template <class T>
__device__ __forceinline__ T cubic_step_perm(T u0, T u1, T u2, T u3, T tt, bool edge, bool lin, bool all_none) {
    using O = Ops<T>;
    const T half = T(0.5), two = T(2);
    const T d20 = O::sub(u2, u0);
    const T dy = O::sub(u2, u1);
    const T a = O::fma(half, d20, -dy);
    if (all_none) {
        const T d31 = O::sub(u3, u1);
        const T b = O::fma(-half, d31, dy);
        const T c1 = O::add(dy, a);
        const T c2 = O::fma(-two, a, b);
        const T c3 = O::sub(a, b);
        return O::add(u1, O::mul(tt, O::add(c1, O::mul(tt, O::add(c2, O::mul(tt, c3))))));
    }
    const T k0 = O::mul(d20, half);
    const T knat = O::fma(two, dy, -k0);
    const T kint = O::mul(O::sub(u3, u1), half);
    const T k1 = edge ? knat : kint;
    const T b = O::sub(dy, k1);
    const T c1 = O::add(dy, a);
    const T c2 = O::fma(-two, a, b);
    const T c3 = O::sub(a, b);
    const T cub = O::add(u1, O::mul(tt, O::add(c1, O::mul(tt, O::add(c2, O::mul(tt, c3))))));
    const T linv = O::add(u2, O::mul(k1, O::sub(tt, T(1))));
    return lin ? linv : cub;
}

// The same step on inputs in natural order: the permutation is done with selects (dimension N-2, whose four inputs
// sit in one sector).
template <class T>
__device__ __forceinline__ T cubic_step_sel(T v0, T v1, T v2, T v3, T tt, int mode, bool lin, bool all_none) {
    if (all_none) return cubic_step_perm(v0, v1, v2, v3, tt, false, false, true);
    const bool low = mode == kModeLow, high = mode == kModeHigh;
    const T u0 = low ? v2 : (high ? v1 : v0);
    const T u1 = high ? v2 : v1;
    const T u2 = low ? v0 : (high ? v3 : v2);
    return cubic_step_perm(u0, u1, u2, v3, tt, low || high, lin, false);
}

// Row (or lane) order of the permuted inputs: interior 0,1,2,3; low end 2,1,0,3; high end 1,2,3,3.
__device__ __forceinline__ void cubic_perm(int mode, int (&k)[4]) {
    const bool low = mode == kModeLow, high = mode == kModeHigh;
    k[0] = low ? 2 : (high ? 1 : 0);
    k[1] = high ? 2 : 1;
    k[2] = low ? 0 : (high ? 3 : 2);
    k[3] = 3;
}

// Cell location of one point on every dimension (ref: multicubic/regular.rs:432-469 and :356-360).
template <class T, int N>
__device__ __forceinline__ bool quad4_locate(const EvalArgs<T, N>& a, const T (&x)[N], QuadSlot<T, N>& s) {
    using O = Ops<T>;
    bool ok = true;
    int base = 0, flags = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        int f = 0;
        bool proven = false;
        if constexpr (sizeof(T) == 8) {
            // f~ = floor(RN(d * RN(1/step))) with the FMA remainder as proof (device_math.cuh fast_cell; here the
            // unclamped f is needed, so the remainder is always taken against f~ itself).
            const double dd = __dsub_rn(x[d], a.start[d]);
            f = __double2int_rd(__dmul_rn(dd, a.rstep[d]));
            const double r = __fma_rn(-__int2double_rn(f), a.step[d], dd);
            proven = a.fast_div != 0 && r >= 0.0 && r <= a.lim[d] && static_cast<unsigned>(f) + (1u << 30) <= (1u << 31);
        }
        if (!proven) ok = floor_cell(x[d], a.start[d], a.step[d], a.rstep[d], a.fast_div != 0, f) && ok;
        const int dim = a.dim[d];
        const int origin = min(max(f, 1) - 1, dim - 4);
        int mode;
        bool outside;
        if (f < 0) { mode = kModeLow; outside = true; }
        else if (f == 0) { mode = kModeLow; outside = false; }
        else if (f > dim - 2) { mode = kModeHigh; outside = true; }
        else if (f == dim - 2) { mode = kModeHigh; outside = false; }
        else { mode = kModeNone; outside = false; }
        const T x1 = O::add(a.start[d], O::mul(a.step[d], O::from_int(origin + 1)));
        const T t = exact_div(O::sub(x[d], x1), a.step[d], a.rstep[d], a.fast_div != 0);
        s.tt[d] = mode == kModeNone ? t : (mode == kModeLow ? -t : O::sub(t, T(1)));
        base += origin * a.istride[d];
        flags |= (mode | ((outside && a.linearize) ? 4 : 0)) << (3 * d);
    }
    s.base = base;
    s.flags = flags;
    return ok;
}

// Reduces dimensions 0..D-1 (address dimensions, D <= N-2) of the sub-block at sector index `idx` for the four
// in-sector positions at once. ro[d][k] = offset of the k-th permuted row of dimension d.
template <int D, class T, int N>
__device__ __forceinline__ void quad4_rows(const T* __restrict__ win, int idx, const int (&ro)[N][4], const T (&tt)[N],
                                           int flags, unsigned none_mask, T (&out)[4]) {
    if constexpr (D == 0) {
        load_row<T, 4, true, int>(nullptr, win, idx, out);
    } else {
        T sub[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) quad4_rows<D - 1, T, N>(win, idx + ro[D - 1][k], ro, tt, flags, none_mask, sub[k]);
        const int fl = flags >> (3 * (D - 1));
        const bool edge = (fl & 3) != 0, lin = (fl & 4) != 0, all_none = (none_mask >> (D - 1)) & 1u;
        if (all_none) {  // one warp-uniform branch around the four independent steps: they interleave
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = cubic_step_perm(sub[0][j], sub[1][j], sub[2][j], sub[3][j], tt[D - 1], false, false, true);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = cubic_step_perm(sub[0][j], sub[1][j], sub[2][j], sub[3][j], tt[D - 1], edge, lin, false);
        }
    }
}

template <class T, int N, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) cubic_quad4_kernel(const __grid_constant__ EvalArgs<T, N> a) {
    static_assert(N >= 2 && N <= 4, "quad-cooperative cubic covers N = 2..4");
    using Slot = QuadSlot<T, N>;
    constexpr int kWarps = kBlock / 32;
    // Parameter slots: one per lane, 16 bytes of padding per quad so that the eight quads of a warp read eight
    // different bank groups. Transposition buffer: [quad][lane j][point p], quad stride 16 + 4 elements.
    constexpr int kSlotWarpBytes = 32 * static_cast<int>(sizeof(Slot)) + 8 * 16;
    constexpr int kXposeQuad = 20;
    __shared__ __align__(16) unsigned char s_slots[kWarps * kSlotWarpBytes];
    __shared__ __align__(16) T s_xpose[kWarps * 8 * kXposeQuad];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, b = lane & 3u, quad = lane >> 2;
    unsigned char* wslots = s_slots + warp * kSlotWarpBytes;
    Slot* myslot = reinterpret_cast<Slot*>(wslots + lane * sizeof(Slot) + quad * 16);
    T* xq = s_xpose + (warp * 8 + quad) * kXposeQuad;

    BlockSchedule sched;
    while (sched.next(a.work, a.n)) {
        const unsigned long long i = sched.blk * blockDim.x + threadIdx.x;
        const bool valid = i < a.n;
        const unsigned long long il = valid ? i : a.n - 1;
        T x[N];
#pragma unroll
        for (int d = 0; d < N; ++d) x[d] = load_query(a.obs[d] + il);
        Slot mine;
        const bool ok = quad4_locate<T, N>(a, x, mine);
        if (!ok) mine.base = 0;  // keep the gathers in range; the result is discarded
        *myslot = mine;
        unsigned edges[N];  // lanes (= points) of the warp that are in an end cell of dimension d
#pragma unroll
        for (int d = 0; d < N; ++d) edges[d] = __ballot_sync(0xffffffffu, ((mine.flags >> (3 * d)) & 3) != 0);
        __syncwarp();

        T s[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const Slot sp = *reinterpret_cast<const Slot*>(wslots + ((lane & ~3u) + p) * sizeof(Slot) + quad * 16);
            unsigned none_mask = 0;
#pragma unroll
            for (int d = 0; d < N; ++d) none_mask |= (edges[d] & (0x11111111u << p)) == 0u ? (1u << d) : 0u;
            int ro[N][4];
#pragma unroll
            for (int d = 0; d + 2 < N; ++d) {
                if ((none_mask >> d) & 1u) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ro[d][k] = k * a.istride[d];
                } else {
                    int k4[4];
                    cubic_perm((sp.flags >> (3 * d)) & 3, k4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ro[d][k] = k4[k] * a.istride[d];
                }
            }
            T r[4];
            quad4_rows<N - 2, T, N>(a.win, sp.base + static_cast<int>(b), ro, sp.tt, sp.flags, none_mask, r);
            const int fl = sp.flags >> (3 * (N - 2));
            s[p] = cubic_step_sel(r[0], r[1], r[2], r[3], sp.tt[N - 2], fl & 3, (fl & 4) != 0, (none_mask >> (N - 2)) & 1u);
        }
        // Transposition: lane j holds the partial results of its last-dimension node for points 0..3; the owner of
        // point b needs the four nodes' results of point b, in the permuted order of its saturation class.
#pragma unroll
        for (int p = 0; p < 4; ++p) xq[b * 4 + p] = s[p];
        __syncwarp();
        const int fl = mine.flags >> (3 * (N - 1));
        int k4[4];
        cubic_perm(fl & 3, k4);
        const T w0 = xq[k4[0] * 4 + b], w1 = xq[k4[1] * 4 + b], w2 = xq[k4[2] * 4 + b], w3 = xq[12 + b];
        const T res = cubic_step_perm(w0, w1, w2, w3, mine.tt[N - 1], (fl & 3) != 0, (fl & 4) != 0, edges[N - 1] == 0u);
        __syncwarp();  // the next iteration overwrites both buffers
        if (valid) {
            if (ok) store_result(a.out + i, res);
            else report_bad(a, i);
        }
    }
}

}  // namespace ib200
