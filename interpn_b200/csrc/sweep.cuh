// sweep.cuh — bin-swept evaluation for grids that do not fit in L2.
//
// Random query points on a grid larger than the 126 MB L2 turn every corner row into a DRAM
// sector fetch: a 6-D multilinear footprint is 32 rows, a 4-D multicubic one 64, and HBM serves
// random 32-byte sectors at ~50 G/s (profiles/r1_microbench_b200.json), far below what the same
// gathers reach from L2 (~288 G/s). The sweep restores locality by evaluating the batch in chunks
// of a few million points, each chunk in the order of a coarse cell key over the leading dimensions:
//
//   1. sweep_hist_kernel     key of every point (approximate cell of the leading dimensions), global histogram
//   2. sweep_scan_kernel     exclusive scan of the histogram -> start of every key's run
//   3. sweep_scatter_kernel  coordinates copied into key order (counting sort; every CTA reserves one
//                            contiguous run per key, so only nbins write frontiers are open at a time
//                            and L2 merges the 8-byte stores into full lines)
//   4. the ordinary evaluation kernel of the method (kernels.cuh) on the sorted coordinates: consecutive
//      CTAs now gather from the same few-megabyte slab of the grid, which is read from HBM once per chunk
//   5. sweep_unsort_kernel   out[i] = result[position of i]; the chunk's results (<= 32 MB) are still in L2
//
// Per-point arithmetic is the unmodified device code of the direct kernels, so results are
// bit-identical; the key only decides the order of evaluation and may be approximate. The smallest
// failing original index is still reported (EvalArgs::remap), and failing points are never written.
#pragma once
#include "launch_common.cuh"

namespace ib200 {

constexpr int kSweepKeyDims = 3;
constexpr int kSweepMaxBins = 4096;
// Histogram / scatter tile (16-bit tile-local indices) and the scatter CTA. The scatter is a chain of shared-memory phases
// separated by barriers with a DRAM round trip inside most of them, so what hides its latency is SEVERAL resident CTAs
// per SM, not a big one? Measured (round 2, gpurun_out/r2_exp1): NO — 8192-point tiles on one 1024-thread CTA per SM
// (C4 7.95 G points/s) beat 4096/512 x3 (7.27), 4096/1024 x2 (7.59) and 2048/256 x6 (6.48): smaller tiles shorten the
// runs of a (tile, key) pair (C4: 529 keys, 15 points per run at 8192) and that costs more than the extra overlap buys.
#ifndef IB200_SWEEP_TILE
#define IB200_SWEEP_TILE 8192
#endif
#ifndef IB200_SWEEP_SCATTER_BLOCK
#define IB200_SWEEP_SCATTER_BLOCK 1024
#endif
#ifndef IB200_SWEEP_SCATTER_MINB
#define IB200_SWEEP_SCATTER_MINB 1
#endif
constexpr unsigned kSweepTile = IB200_SWEEP_TILE;
constexpr int kSweepScatterBlock = IB200_SWEEP_SCATTER_BLOCK;
constexpr int kSweepScatterMinBlocks = IB200_SWEEP_SCATTER_MINB;
static_assert(kSweepTile <= 65536 && kSweepTile % kSweepScatterBlock == 0, "16-bit tile-local indices, whole elements per thread");
constexpr int kSweepSlots = 16;            // cursors per key: tile t uses slot t % 16 (same-address atomics serialise in L2)

struct SweepKey {
    int nkey;                    // leading dimensions that enter the key (1..3)
    float rgrp[kSweepKeyDims];   // 1 / (cells per key group)
    int ngrp[kSweepKeyDims];     // key groups along the dimension
    int mult[kSweepKeyDims];     // key = sum(group_d * mult_d), dimension 0 most significant
    int nbins;
};

// Approximate cell of x along dimension d (locality only: any deterministic value will do).
template <class T, int N, bool RECT>
__device__ __forceinline__ int sweep_cell(const EvalArgs<T, N>& a, const T* __restrict__ axes, int d, T x) {
    if constexpr (RECT) {
        return clamp_cell(rect_lower_bound<T, N>(a, axes, d, x) - 1, a.dim[d] - 2);
    } else {
        return clamp_cell(Ops<T>::floor_sat((x - a.start[d]) * a.rstep[d]), a.dim[d] - 2);
    }
}

template <class T, int N, bool RECT>
__global__ void __launch_bounds__(kBlock) sweep_hist_kernel(const __grid_constant__ EvalArgs<T, N> a,
                                                            const __grid_constant__ SweepKey k, int hist_off,
                                                            unsigned short* __restrict__ keys, unsigned* __restrict__ ghist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* axes = nullptr;
    if constexpr (RECT) axes = stage_axes<T, N>(a);
    unsigned* hist = reinterpret_cast<unsigned*>(smem_raw + hist_off);
    for (unsigned b = threadIdx.x; b < static_cast<unsigned>(k.nbins); b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    // Tiles of kSweepTile points, tile t -> CTA t % gridDim; gridDim is a multiple of
    // kSweepSlots, so all tiles of this CTA share the slot blockIdx % kSweepSlots (= t % kSweepSlots).
    const unsigned long long tile = kSweepTile;
    for (unsigned long long i0 = blockIdx.x * tile; i0 < a.n; i0 += gridDim.x * tile) {
        const unsigned long long iend = min(i0 + tile, a.n);
        for (unsigned long long i = i0 + threadIdx.x; i < iend; i += kBlock) {
            unsigned key = 0;
#pragma unroll
            for (int d = 0; d < kSweepKeyDims; ++d) {
                if (d < N && d < k.nkey) {
                    const int c = sweep_cell<T, N, RECT>(a, axes, d, __ldg(a.obs[d] + i));
                    const int grp = min(__float2int_rd(__int2float_rn(c) * k.rgrp[d]), k.ngrp[d] - 1);
                    key += static_cast<unsigned>(grp * k.mult[d]);
                }
            }
            keys[i] = static_cast<unsigned short>(key);
            atomicAdd(&hist[key], 1u);
        }
    }
    __syncthreads();
    const unsigned slot = blockIdx.x % kSweepSlots;
    for (unsigned b = threadIdx.x; b < static_cast<unsigned>(k.nbins); b += blockDim.x) {
        const unsigned c = hist[b];
        if (c) atomicAdd(&ghist[b * kSweepSlots + slot], c);
    }
}

// One CTA: cursor[b] = number of points with a key below b (in place).
static __global__ void __launch_bounds__(kBlock) sweep_scan_kernel(unsigned* __restrict__ cursor, int n) {
    __shared__ unsigned warp_tot[32];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int per = (n + nthr - 1) / nthr;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    unsigned sum = 0;
    for (int j = lo; j < hi; ++j) sum += cursor[j];
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        const unsigned w = tid < (nthr >> 5) ? warp_tot[tid] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
            if (tid >= o) winc += v;
        }
        warp_tot[tid] = winc - w;
    }
    __syncthreads();
    unsigned run = warp_tot[tid >> 5] + inc - sum;
    for (int j = lo; j < hi; ++j) {
        const unsigned c = cursor[j];
        cursor[j] = run;
        run += c;
    }
}

template <class T, int N>
struct SweepScatterArgs {
    const T* obs[N];
    T* sx[N];                     // coordinates in key order
    unsigned long long n;
    const unsigned short* keys;
    unsigned* cursor;             // [nbins][kSweepSlots] next free position of every (key, slot) run
    unsigned* orig;               // orig[p] = original index of the point at sorted position p
    unsigned* pos;                // pos[i]  = sorted position of original point i
    int nbins;
};

// Counting sort of one tile at a time through shared memory, so that both sides of the copy are
// coalesced: every coordinate array of the tile is staged in shared memory (read in original order)
// and written out in sorted order, consecutive threads to consecutive positions of a key's run.
template <class T, int N>
__global__ void __launch_bounds__(kSweepScatterBlock, kSweepScatterMinBlocks) sweep_scatter_kernel(const __grid_constant__ SweepScatterArgs<T, N> s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stage = reinterpret_cast<T*>(smem_raw);                       // [kSweepTile] one coordinate array of the tile
    unsigned* krank = reinterpret_cast<unsigned*>(stage + kSweepTile);  // [kSweepTile] key << 16 | rank, by tile-local index
    unsigned* sorted = krank + kSweepTile;                           // [kSweepTile] key << 16 | tile-local index, by sorted position
    unsigned* cnt = sorted + kSweepTile;                             // [nbins] counts, then global base of the tile's run
    unsigned* loff = cnt + s.nbins;                                  // [nbins] tile-local start of every key's run
    __shared__ unsigned warp_tot[32];
    const unsigned tid = threadIdx.x, nthr = blockDim.x;
    const unsigned long long ntiles = (s.n + kSweepTile - 1) / kSweepTile;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long i0 = t * kSweepTile;
        const unsigned nt = static_cast<unsigned>(min(static_cast<unsigned long long>(kSweepTile), s.n - i0));
        const unsigned slot = static_cast<unsigned>(t % kSweepSlots);
        for (unsigned b = tid; b < static_cast<unsigned>(s.nbins); b += nthr) cnt[b] = 0u;
        __syncthreads();
        {
            constexpr unsigned kPerK = kSweepTile / kSweepScatterBlock;
            unsigned key[kPerK];
#pragma unroll
            for (unsigned k = 0; k < kPerK; ++k) {
                const unsigned li = tid + k * kSweepScatterBlock;
                key[k] = li < nt ? s.keys[i0 + li] : 0u;
            }
#pragma unroll
            for (unsigned k = 0; k < kPerK; ++k) {
                const unsigned li = tid + k * kSweepScatterBlock;
                if (li < nt) krank[li] = (key[k] << 16) | atomicAdd(&cnt[key[k]], 1u);
            }
        }
        __syncthreads();
        // tile-local exclusive scan of the counts -> loff; reserve the tile's global run of every key
        {
            const int n = s.nbins, per = (n + static_cast<int>(nthr) - 1) / static_cast<int>(nthr);
            const int lo = min(static_cast<int>(tid) * per, n), hi = min(lo + per, n);
            unsigned sum = 0;
            for (int j = lo; j < hi; ++j) sum += cnt[j];
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                if ((tid & 31) >= static_cast<unsigned>(o)) inc += v;
            }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = inc;
            __syncthreads();
            if (tid < 32) {
                const unsigned w = tid < (nthr >> 5) ? warp_tot[tid] : 0u;
                unsigned winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
                    if (tid >= static_cast<unsigned>(o)) winc += v;
                }
                warp_tot[tid] = winc - w;
            }
            __syncthreads();
            unsigned run = warp_tot[tid >> 5] + inc - sum;
            for (int j = lo; j < hi; ++j) {
                const unsigned c = cnt[j];
                loff[j] = run;
                run += c;
                if (c) cnt[j] = atomicAdd(&s.cursor[static_cast<size_t>(j) * kSweepSlots + slot], c);
            }
        }
        __syncthreads();
        for (unsigned li = tid; li < nt; li += nthr) {
            const unsigned kr = krank[li];
            sorted[loff[kr >> 16] + (kr & 0xffffu)] = (kr & 0xffff0000u) | li;
        }
        __syncthreads();
        // pos in original order (coalesced): position = global base of the tile's run of the key + rank
        for (unsigned li = tid; li < nt; li += nthr) {
            const unsigned kr = krank[li];
            s.pos[i0 + li] = cnt[kr >> 16] + (kr & 0xffffu);
        }
        __syncthreads();
        // sorted[j] -> global position of sorted element j (kept in krank, no longer needed), and orig
        for (unsigned j = tid; j < nt; j += nthr) {
            const unsigned kl = sorted[j];
            const unsigned key = kl >> 16, li = kl & 0xffffu;
            const unsigned p = cnt[key] + (j - loff[key]);
            krank[j] = p;
            s.orig[p] = static_cast<unsigned>(i0 + li);
        }
        // A thread owns kPer = tile / block elements; all of its loads are issued before the first use so that the
        // kernel is not one DRAM round trip per element (ncu: 41 long-scoreboard stall cycles per issue before). The
        // sorted positions and source slots do not depend on the dimension: read once, kept in registers.
        constexpr unsigned kPer = kSweepTile / kSweepScatterBlock;
        unsigned dstp[kPer], srcl[kPer];
#pragma unroll
        for (unsigned k = 0; k < kPer; ++k) {
            const unsigned j = tid + k * kSweepScatterBlock;
            dstp[k] = j < nt ? krank[j] : 0u;
            srcl[k] = j < nt ? (sorted[j] & 0xffffu) : 0u;
        }
#pragma unroll 1
        for (int d = 0; d < N; ++d) {
            T v[kPer];
#pragma unroll
            for (unsigned k = 0; k < kPer; ++k) {
                const unsigned li = tid + k * kSweepScatterBlock;
                if (li < nt) v[k] = load_query(s.obs[d] + i0 + li);
            }
            __syncthreads();  // the previous dimension's readers are done with `stage`
#pragma unroll
            for (unsigned k = 0; k < kPer; ++k) {
                const unsigned li = tid + k * kSweepScatterBlock;
                if (li < nt) stage[li] = v[k];
            }
            __syncthreads();
#pragma unroll
            for (unsigned k = 0; k < kPer; ++k) v[k] = stage[srcl[k]];
#pragma unroll
            for (unsigned k = 0; k < kPer; ++k)
                if (tid + k * kSweepScatterBlock < nt) s.sx[d][dstp[k]] = v[k];
        }
        __syncthreads();
    }
}

// out[i] = res[pos[i]]. Points the evaluation refused (unrepresentable coordinate) are not written:
// only when the launch has recorded such a point (first_bad) is the exact predicate re-evaluated.
template <class T, int N>
__global__ void __launch_bounds__(kBlock) sweep_unsort_kernel(const __grid_constant__ EvalArgs<T, N> a, const T* __restrict__ res,
                                                              const unsigned* __restrict__ pos, bool regular) {
    const bool any_bad = regular && *reinterpret_cast<volatile unsigned long long*>(a.first_bad) != kNoBad;
    const unsigned long long gstride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long gtid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (!any_bad) {
        // common case: four dependent (pos -> res) gathers in flight per thread
        constexpr int U = 4;
        unsigned long long i = gtid;
        for (; i + (U - 1) * gstride < a.n; i += U * gstride) {
            unsigned p[U];
            T v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) p[k] = __ldg(pos + i + k * gstride);
#pragma unroll
            for (int k = 0; k < U; ++k) v[k] = __ldcg(res + p[k]);
#pragma unroll
            for (int k = 0; k < U; ++k) store_result(a.out + i + k * gstride, v[k]);
        }
        for (; i < a.n; i += gstride) store_result(a.out + i, __ldcg(res + __ldg(pos + i)));
        return;
    }
    for (unsigned long long i = gtid; i < a.n; i += gstride) {
        if (any_bad) {
            bool ok = true;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                int iloc;
                ok = floor_cell(__ldg(a.obs[d] + i), a.start[d], a.step[d], a.rstep[d], false, iloc) && ok;
            }
            if (!ok) continue;
        }
        store_result(a.out + i, __ldcg(res + __ldg(pos + i)));
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------

size_t sweep_env(const char* name, size_t fallback);
// The library's own stream-ordered memory pool on the current device (launch_misc.cu): the sweep scratch is cached there
// between calls, so the process-wide default pool — which torch, RMM or cuDF may share — keeps its own release policy.
cudaMemPool_t sweep_scratch_pool();

// Decide whether `n` points on grid `g` are worth sweeping and fill the key description. `fp` is the
// footprint width (2 linear, 4 cubic); `row_bytes_scale` the blow-up of the array the kernel gathers
// from (1 for vals, W for the window layout); `rows` the 32-byte sectors a point's footprint touches; `fp0` the
// footprint along dimension 0 of the gathered array (1 for the cubic coefficient layout: one slot per cell).
template <class T, int N>
inline bool plan_sweep(const DeviceGrid& g, size_t n, int fp, int row_bytes_scale, int rows, int fp0, bool axes_in_smem, SweepKey& s) {
    const size_t gathered = g.nvals * sizeof(T) * static_cast<size_t>(row_bytes_scale);
    const size_t min_bytes = sweep_env("INTERPN_B200_SWEEP_MIN_MB", 96) << 20;
    const size_t min_points = sweep_env("INTERPN_B200_SWEEP_MIN_POINTS", size_t(1) << 21);
    const size_t min_rows = sweep_env("INTERPN_B200_SWEEP_MIN_ROWS", 16);
    if (N < 2 || gathered <= min_bytes || n < min_points || static_cast<size_t>(rows) < min_rows) return false;
    if (index64(g) || (g.rect && !axes_in_smem)) return false;
    // Slab of the gathered array under one key. Multilinear evaluation is cheap next to the sort, and fewer, larger
    // slabs mean longer runs per (tile, key) in the scatter: 48 MB measured best on C4 (6.1 -> 7.2 G points/s; 96 MB
    // falls out of L2). The multicubic kernels dominate their sort and are flat between 6 and 24 MB.
    const size_t slab_kb = sweep_env("INTERPN_B200_SWEEP_SLAB_KB", fp == 2 ? 49152 : 12288);
    const double target = slab_kb ? static_cast<double>(slab_kb << 10) : 128.0;  // 0: as fine as it gets (tests)
    double slab = static_cast<double>(gathered);
    int bins = 1;
    s.nkey = 0;
    for (int d = 0; d < N - 1 && d < kSweepKeyDims; ++d) {
        if (slab <= target) break;
        const int cells = g.dim[d] - 1;  // sweep_cell() values 0 .. dim-2
        const int fpd = d == 0 ? fp0 : fp;
        int per = static_cast<int>(target * g.dim[d] / slab) - (fpd - 1);
        per = per < 1 ? 1 : (per > cells ? cells : per);
        int groups = (cells + per - 1) / per;
        if (bins * groups > kSweepMaxBins) {
            groups = kSweepMaxBins / bins;
            if (groups < 2) break;
            per = (cells + groups - 1) / groups;
            groups = (cells + per - 1) / per;
        }
        s.rgrp[d] = 1.0f / static_cast<float>(per);
        s.ngrp[d] = groups;
        slab = slab * (per + fpd - 1) / g.dim[d];
        bins *= groups;
        s.nkey = d + 1;
    }
    if (s.nkey == 0 || bins < 2) return false;
    int mult = 1;
    for (int d = s.nkey - 1; d >= 0; --d) {
        s.mult[d] = mult;
        mult *= s.ngrp[d];
    }
    s.nbins = bins;
    return true;
}

// `eval(obs_sorted, count, res, orig, index_base, work)` launches the method's ordinary kernel with dynamic
// block scheduling (`work`, a zeroed device counter): blocks of points are handed out in key order on demand,
// so the CTAs never drift apart along the key range.
template <class T, int N, bool RECT, class Eval>
inline cudaError_t launch_sweep(const DeviceGrid& g, int fp, int row_bytes_scale, int rows, int fp0, const T* const* obs, size_t n,
                                T* out, unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream,
                                Eval&& eval, bool& used) {
    used = false;
    if (n == 0) return cudaSuccess;
    EvalArgs<T, N> a = make_args<T, N>(g, obs, n, out, first_bad, index_base);
    SweepKey key{};
    if (!plan_sweep<T, N>(g, n, fp, row_bytes_scale, rows, fp0, a.axes_in_smem != 0, key)) return cudaSuccess;
    const size_t axes_bytes = a.axes_in_smem ? (static_cast<size_t>(g.axes_total) * sizeof(T) + 15) / 16 * 16 : 0;
    const size_t hist_smem = axes_bytes + static_cast<size_t>(key.nbins) * sizeof(unsigned);
    cudaError_t e;
    if (hist_smem > 48 * 1024) {
        e = cudaFuncSetAttribute(sweep_hist_kernel<T, N, RECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hist_smem));
        if (e != cudaSuccess) return e;
    }
    // Chunk size: the evaluation kernel keeps ~2000 points per SM in flight, i.e. a window of in_flight/chunk of
    // the key range; the part of the grid under that window (twice, for the halo) must fit in a fraction of L2.
    const size_t scatter_smem = kSweepTile * (sizeof(T) + 2 * sizeof(unsigned)) + 2 * static_cast<size_t>(key.nbins) * sizeof(unsigned);
    e = cudaFuncSetAttribute(sweep_scatter_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem));
    if (e != cudaSuccess) return e;
    size_t chunk = sweep_env("INTERPN_B200_SWEEP_CHUNK", 0);
    if (chunk == 0) {
        const double gathered = static_cast<double>(g.nvals) * sizeof(T) * row_bytes_scale;
        const double in_flight = 2048.0 * g.sm_count;
        chunk = static_cast<size_t>(in_flight * gathered * 2.0 / (24.0 * 1048576.0));
        if (chunk < (size_t(1) << 22)) chunk = size_t(1) << 22;
        if (chunk > (size_t(1) << 26)) chunk = size_t(1) << 26;
        const size_t parts = (n + chunk - 1) / chunk;  // equal parts
        chunk = (n + parts - 1) / parts;
    }
    chunk = chunk < 4096 ? 4096 : (chunk > (size_t(1) << 31) ? (size_t(1) << 31) : chunk);
    if (chunk > n) chunk = n;
    const size_t chunk_al = (chunk + 63) / 64 * 64;

    // Stream-ordered scratch from the library's own pool (kept cached between calls, launch_misc.cu sweep_scratch_pool).
    const size_t bytes = chunk_al * ((N + 1) * sizeof(T) + 2 * sizeof(unsigned) + sizeof(unsigned short)) +
                         static_cast<size_t>(key.nbins) * kSweepSlots * sizeof(unsigned) + 64;
    void* scratch = nullptr;
    e = cudaMallocFromPoolAsync(&scratch, bytes, sweep_scratch_pool(), stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cudaSuccess;  // no memory for the scratch: the direct kernel takes the call
    }
    unsigned char* p = static_cast<unsigned char*>(scratch);
    T* sx[N];
    for (int d = 0; d < N; ++d) sx[d] = reinterpret_cast<T*>(p) + static_cast<size_t>(d) * chunk_al;
    T* res = reinterpret_cast<T*>(p) + static_cast<size_t>(N) * chunk_al;
    unsigned* orig = reinterpret_cast<unsigned*>(res + chunk_al);
    unsigned* pos = orig + chunk_al;
    unsigned* cursor = pos + chunk_al;
    unsigned long long* work = reinterpret_cast<unsigned long long*>(cursor + static_cast<size_t>(key.nbins) * kSweepSlots);
    unsigned short* keys = reinterpret_cast<unsigned short*>(work + 8);

    auto body = [&]() -> cudaError_t {
        for (size_t c0 = 0; c0 < n; c0 += chunk) {
            const size_t cnt = n - c0 < chunk ? n - c0 : chunk;
            const T* cobs[N];
            for (int d = 0; d < N; ++d) cobs[d] = obs[d] + c0;
            EvalArgs<T, N> ca = make_args<T, N>(g, cobs, cnt, out + c0, first_bad, index_base + c0);
            cudaError_t err = cudaMemsetAsync(cursor, 0, static_cast<size_t>(key.nbins) * kSweepSlots * sizeof(unsigned) + 64, stream);
            if (err != cudaSuccess) return err;
            const size_t tiles = (cnt + kSweepTile - 1) / kSweepTile;
            unsigned hgrid = grid_for(tiles * kBlock, g.sm_count, 8);
            hgrid = (hgrid + kSweepSlots - 1) / kSweepSlots * kSweepSlots;  // see sweep_hist_kernel
            sweep_hist_kernel<T, N, RECT><<<hgrid, kBlock, hist_smem, stream>>>(ca, key, static_cast<int>(axes_bytes), keys, cursor);
            sweep_scan_kernel<<<1, kBlock, 0, stream>>>(cursor, key.nbins * kSweepSlots);
            SweepScatterArgs<T, N> sa{};
            for (int d = 0; d < N; ++d) {
                sa.obs[d] = cobs[d];
                sa.sx[d] = sx[d];
            }
            sa.n = cnt;
            sa.keys = keys;
            sa.cursor = cursor;
            sa.orig = orig;
            sa.pos = pos;
            sa.nbins = key.nbins;
            sweep_scatter_kernel<T, N><<<grid_for(tiles * kBlock, g.sm_count, kSweepScatterMinBlocks), kSweepScatterBlock, scatter_smem, stream>>>(sa);
            count_launch();
            count_launch();
            count_launch();
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
            err = eval(const_cast<const T* const*>(sx), cnt, res, orig, index_base + c0, work);
            if (err != cudaSuccess) return err;
            sweep_unsort_kernel<T, N><<<grid_for(cnt, g.sm_count, 8), kBlock, 0, stream>>>(ca, res, pos, !RECT);
            count_launch();
            count_swept_launch();
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
        }
        return cudaSuccess;
    };
    e = body();
    cudaError_t e2 = cudaFreeAsync(scratch, stream);
    used = true;
    return e != cudaSuccess ? e : e2;
}

}  // namespace ib200
