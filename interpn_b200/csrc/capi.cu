// capi.cu — the extern "C" boundary declared in include/interpn_b200.h.
//
// Host-side logic only: argument validation restating the reference's dispatchers and `new()`
// constructors, grid residency, the host<->device copy pipeline of the host-buffer entry points,
// and error mapping. All arithmetic on query points happens in the CUDA kernels; there is no CPU
// evaluation path in this library.
#include "../../include/interpn_b200.h"

#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "interp_internal.h"

namespace ib200 {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<uint64_t> g_swept{0};
void count_swept_launch() { g_swept.fetch_add(1, std::memory_order_relaxed); }

static thread_local char t_detail[512] = "";

static int cuda_fail(cudaError_t e, const char* what, int line) {
    snprintf(t_detail, sizeof(t_detail), "%s (%s) at capi.cu:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), line,
             what);
    cudaGetLastError();  // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? INTERPN_B200_ERR_TOO_LARGE : INTERPN_B200_ERR_CUDA;
}

#define CUDA_TRY(expr)                                                   \
    do {                                                                 \
        cudaError_t e_ = (expr);                                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr, __LINE__);    \
    } while (0)

// Fails loudly when there is no usable B200: the product has no CPU fallback.
static int require_device(int* sm_count = nullptr) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(t_detail, sizeof(t_detail), "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return INTERPN_B200_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) {
        snprintf(t_detail, sizeof(t_detail), "device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return INTERPN_B200_ERR_NO_DEVICE;
    }
    if (sm_count) *sm_count = sms;
    return INTERPN_B200_OK;
}

}  // namespace ib200

#include "host_exec.cuh"

using namespace ib200;

// The opaque interpolator: grid resident in HBM + the plumbing its evaluations need.
struct interpn_b200_interp {
    DeviceGrid g;
    int device = 0;
    unsigned long long* first_bad_dev = nullptr;  // latched by eval_device launches
    // Host-buffer evaluation (host_exec.cuh): slots of the home device, and — when one call may use several GPUs — one
    // replica of the grid per further device, copied device-to-device on first use and dropped by vals_updated().
    DeviceSlots home;
    struct Replica {
        DeviceGrid g;
        DeviceSlots slots;
    };
    std::vector<Replica*> replicas;
    std::mutex host_mu;  // one host-buffer evaluation per interpolator at a time
};

namespace {

unsigned long long fnv1a(unsigned long long h, const void* p, size_t bytes) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

size_t product(const size_t* d, size_t n) {
    size_t p = 1;
    for (size_t i = 0; i < n; ++i) p *= d[i];
    return p;
}

int max_dims_status(int method) { return method == INTERPN_B200_NEAREST ? INTERPN_B200_ERR_MAXDIM_6 : INTERPN_B200_ERR_MAXDIM_8; }
size_t max_dims(int method) { return method == INTERPN_B200_NEAREST ? 6 : 8; }

// `Struct::new` for regular grids (ref: multilinear/regular.rs:225-259, regular_recursive.rs:191-233,
// multicubic/regular.rs:239-288, regular_recursive.rs:236-281, nearest/regular.rs:163-197).
template <class T>
int validate_regular_new(int method, const size_t* dims, size_t ndims, size_t nstarts, const T* steps, size_t nsteps,
                         size_t nvals) {
    if (method < 0 || method > 2) return INTERPN_B200_ERR_INVALID_ARG;
    if (ndims < 1 || ndims > max_dims(method)) return max_dims_status(method);
    if (nstarts != ndims || nsteps != ndims) return INTERPN_B200_ERR_DIM_MISMATCH;
    if (nvals != product(dims, ndims)) return INTERPN_B200_ERR_DIM_MISMATCH;
    const bool cubic = method == INTERPN_B200_CUBIC;
    for (size_t i = 0; i < ndims; ++i)
        if (dims[i] < (cubic ? 4u : 2u)) return cubic ? INTERPN_B200_ERR_MIN_FOUR : INTERPN_B200_ERR_MIN_TWO;
    for (size_t i = 0; i < ndims; ++i)
        if (!(steps[i] > T(0))) return INTERPN_B200_ERR_NOT_MONOTONIC;
    for (size_t i = 0; i < ndims; ++i)
        if (dims[i] > size_t(INT_MAX) - 8) return INTERPN_B200_ERR_TOO_LARGE;
    return INTERPN_B200_OK;
}

// `Struct::new` for rectilinear grids (ref: multilinear/rectilinear.rs:175-201,
// rectilinear_recursive.rs:160-181, multicubic/rectilinear.rs:193-228, nearest/rectilinear.rs:124-150).
template <class T>
int validate_rect_new(int method, const T* const* grids, const size_t* grid_lens, size_t ngrids, size_t nvals) {
    if (method < 0 || method > 2) return INTERPN_B200_ERR_INVALID_ARG;
    if (ngrids < 1 || ngrids > max_dims(method)) return max_dims_status(method);
    if (nvals != product(grid_lens, ngrids)) return INTERPN_B200_ERR_DIM_MISMATCH;
    const bool cubic = method == INTERPN_B200_CUBIC;
    for (size_t i = 0; i < ngrids; ++i)
        if (grid_lens[i] < (cubic ? 4u : 2u)) return cubic ? INTERPN_B200_ERR_MIN_4 : INTERPN_B200_ERR_MIN_2;
    for (size_t i = 0; i < ngrids; ++i)
        if (!(grids[i][1] > grids[i][0])) return INTERPN_B200_ERR_NOT_MONOTONIC;  // only the first two entries
    size_t total = 0;
    for (size_t i = 0; i < ngrids; ++i) {
        if (grid_lens[i] > size_t(INT_MAX) - 8) return INTERPN_B200_ERR_TOO_LARGE;
        total += grid_lens[i];
    }
    if (total > size_t(INT_MAX) - 8) return INTERPN_B200_ERR_TOO_LARGE;
    return INTERPN_B200_OK;
}

// `Struct::interp` checks (ref: multilinear/regular.rs:268-274, regular_recursive.rs:242-253).
int validate_interp(size_t ndims, const size_t* obs_lens, size_t nobs, size_t nout) {
    if (nobs != ndims) return INTERPN_B200_ERR_DIM_MISMATCH;
    for (size_t j = 0; j < nobs; ++j)
        if (obs_lens[j] != nout) return INTERPN_B200_ERR_DIM_MISMATCH;
    return INTERPN_B200_OK;
}

// Which derived copy of `vals` a grid gets (interp_internal.h DeviceGrid::win). The patch / coefficient
// copies multiply the grid by ~4 and pay while that copy sits in L2 (direct kernels) or is swept slab by slab (sweep.cuh);
// the hypercube copies multiply it by 8 / 16 and serve grids beyond L2 from HBM one line per request.
// INTERPN_B200_WINDOW_MB / INTERPN_B200_HYPER_MAX_MB bound them (0 disables).
int window_width(const DeviceGrid& g) {
    int w = 0;
    if (g.method == INTERPN_B200_LINEAR && g.ndims <= 6) {
        // N >= 2: 2x2 patches (4-fold copy), except for N <= 4 grids between 48 and 100 KB, whose 2-fold row-pair copy
        // still fits L1 while the 4-fold one would not (C1's 20^3 grid: 137 vs 120 G points/s)
        const size_t b = g.nvals * static_cast<size_t>(g.elem);
        w = (g.ndims >= 2 && !(g.ndims <= 4 && b > (48u << 10) && b <= (100u << 10))) ? 4 : 2;
        // N = 3..6 beyond L2: the hypercube layout (kernels.cuh linear_hyper_kernel / linear_hyper3_kernel) — the corners of a cell
        // over the last four (N = 3: all three) dimensions as one aligned block, a 16-fold (8-fold) copy (C3: 2.1 GB, C4: 24.5 GB)
        // that HBM serves at one line per request; 180 GB of HBM is what it is for. INTERPN_B200_HYPER_MIN_KB / _MAX_MB bound it (0 MB turns it off: slab passes or the
        // bin-swept path, launch_linear.cu).
        size_t hyper_min_kb = 80u << 10, hyper_max_mb = 32768;
        if (const char* e = getenv("INTERPN_B200_HYPER_MIN_KB")) hyper_min_kb = static_cast<size_t>(strtoull(e, nullptr, 10));
        if (const char* e = getenv("INTERPN_B200_HYPER_MAX_MB")) hyper_max_mb = static_cast<size_t>(strtoull(e, nullptr, 10));
        const int hbits = g.ndims == 3 ? 3 : 4;  // N = 3: the whole 2^3 footprint (8-fold copy, read by a lane pair)
        if (g.ndims >= 3 && g.ndims <= 6 && b > (hyper_min_kb << 10) && (b << hbits) <= (hyper_max_mb << 20) &&
            g.nvals < (size_t(1) << 29)) {  // sector index f*4 + j stays below 2^31
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (b << hbits) <= free_b / 4) return 1 << hbits;
        }
    }
    // cubic N = 1: rows of four; N = 2..4: the coefficient layout (cubic_quad4.cuh), which on rectilinear axes is built from
    // the per-cell constant table and therefore exists only for strictly increasing finite axes
    if (g.method == INTERPN_B200_CUBIC && g.ndims <= 4 && !(g.rect && g.ndims >= 2 && !g.rect_cubic_table)) w = 4;
    if (!w) return 0;
    // Kept up to 8 GiB and a quarter of the free memory. Direct kernels gather from it only while it is
    // L2-resident (<= 64 MB, launch_common.cuh kWindowL2Bytes); beyond that it serves the bin-swept path
    // (sweep.cuh), where the slab being gathered from is L2-resident and a row is one L1 wavefront.
    size_t max_mb = 8192;
    if (const char* e = getenv("INTERPN_B200_WINDOW_MB")) max_mb = static_cast<size_t>(strtoull(e, nullptr, 10));
    const size_t bytes = g.nvals * static_cast<size_t>(g.elem);
    size_t min_kb = 0;  // even L1-resident grids gain: half (linear) or a quarter (cubic) as many load instructions
    if (const char* e = getenv("INTERPN_B200_WINDOW_MIN_KB")) min_kb = static_cast<size_t>(strtoull(e, nullptr, 10));
    const size_t wbytes = window_elems(g, w) * static_cast<size_t>(g.elem);
    if (bytes < (min_kb << 10) || wbytes > (max_mb << 20)) return 0;
    if (wbytes > (size_t(64) << 20)) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || wbytes > free_b / 4) return 0;
    }
    return w;
}

int upload_vals(interpn_b200_interp* h, const void* vals, int vals_location) {
    DeviceGrid& g = h->g;
    const size_t bytes = g.nvals * static_cast<size_t>(g.elem);
    CUDA_TRY(cudaMalloc(&g.vals, bytes ? bytes : 1));
    g.win_width = window_width(g);
    // cubic N = 2..4: coefficient layout (quad-cooperative kernels); linear N >= 2: patch layout
    g.win_cross = g.win_width == 4 && g.ndims >= 2;
    g.win_bytes = window_elems(g, g.win_width) * static_cast<size_t>(g.elem);
    if (g.win_width) CUDA_TRY(cudaMalloc(&g.win, g.win_bytes));
    if (vals_location == INTERPN_B200_VALS_UNINIT) return INTERPN_B200_OK;
    if (!vals) return INTERPN_B200_ERR_INVALID_ARG;
    CUDA_TRY(cudaMemcpy(g.vals, vals, bytes,
                        vals_location == INTERPN_B200_VALS_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    CUDA_TRY(launch_build_window(g, nullptr));
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    return INTERPN_B200_OK;
}

int finish_new(interpn_b200_interp* h) {
    CUDA_TRY(cudaGetDevice(&h->device));
    CUDA_TRY(cudaMalloc(&h->first_bad_dev, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(h->first_bad_dev, 0xff, sizeof(unsigned long long)));
    return INTERPN_B200_OK;
}

void set_strides(DeviceGrid& g) {
    long long acc = 1;
    for (int d = g.ndims - 1; d >= 0; --d) {
        g.stride[d] = acc;
        acc *= g.dim[d];
    }
}

template <class T>
int regular_new(int method, const size_t* dims, size_t ndims, const T* starts, size_t nstarts, const T* steps,
                size_t nsteps, const T* vals, size_t nvals, int linearize, int vals_location,
                interpn_b200_interp** out) {
    if (!out) return INTERPN_B200_ERR_INVALID_ARG;
    *out = nullptr;
    if (!dims || !starts || !steps) return INTERPN_B200_ERR_INVALID_ARG;
    int st = validate_regular_new<T>(method, dims, ndims, nstarts, steps, nsteps, nvals);
    if (st != INTERPN_B200_OK) return st;
    int sms = 0;
    st = require_device(&sms);
    if (st != INTERPN_B200_OK) return st;
    auto* h = new (std::nothrow) interpn_b200_interp();
    if (!h) return INTERPN_B200_ERR_TOO_LARGE;
    DeviceGrid& g = h->g;
    g.method = method;
    g.rect = 0;
    g.ndims = static_cast<int>(ndims);
    g.linearize = linearize != 0;
    g.elem = sizeof(T);
    g.nvals = nvals;
    g.sm_count = sms;
    for (size_t d = 0; d < ndims; ++d) {
        g.dim[d] = static_cast<int>(dims[d]);
        g.start[d] = static_cast<double>(starts[d]);  // exact (f32 -> f64 widening is lossless)
        g.step[d] = static_cast<double>(steps[d]);
    }
    set_strides(g);
    g.grid_hash = fnv1a(fnv1a(fnv1a(14695981039346656037ull, g.dim, sizeof(int) * ndims), g.start, sizeof(double) * ndims),
                        g.step, sizeof(double) * ndims);
    st = upload_vals(h, vals, vals_location);
    if (st == INTERPN_B200_OK) st = finish_new(h);
    if (st != INTERPN_B200_OK) {
        interpn_b200_interp_free(h);
        return st;
    }
    *out = h;
    return INTERPN_B200_OK;
}

// Per-cell constants of the 1-D cubic step on a rectilinear axis (cubic_quad4.cuh). Everything the reference's
// interp_inner (multicubic/rectilinear.rs:413-545) derives from the four axis nodes of a footprint depends only on
// pp = partition_point(g < x) in 0..=n: the footprint origin clamp(pp-2, 0, n-4), the saturation class, the spacing
// ratios and the centered_difference_nonuniform weights (multicubic/mod.rs:103-117). They are computed here once,
// in T, with the same IEEE operations in the same order as kernels.cuh cubic_rect_locate (this file is compiled with
// -ffp-contract=off), so a kernel that reads them gets the bits it would have computed — without eight divisions per
// dimension per query point. Only built for strictly increasing finite axes (no NaN/inf can arise).
// Row layout (12 elements, padded to cubic_cell_row_stride): wa, wc (exchanged in a low end cell, see cubic_quad4.cuh), div0, 1/div0, wa1, wc1, div1,
// 1/div1, gref, href, 1/href (t = +-(x - gref)/href), flags (int bits: CubicMode | outside<<2 | ratios in exact_div's
// range<<3 | href in range<<4).
template <class T>
void cubic_cell_table(const T* g, size_t n, std::vector<T>& packed) {
    auto in_range = [](T v) {
        const double a = std::fabs(static_cast<double>(v));
        return sizeof(T) == 8 && a >= 0x1p-300 && a < 0x1p301;
    };
    const T one = T(1);
    for (size_t pp = 0; pp <= n; ++pp) {
        const long iloc = static_cast<long>(pp) - 2;
        const long nn = static_cast<long>(n);
        const long origin = iloc < 0 ? 0 : (iloc > nn - 4 ? nn - 4 : iloc);
        int mode, outside;
        if (iloc == -2) { mode = 1; outside = 1; }
        else if (iloc == -1) { mode = 1; outside = 0; }
        else if (iloc == nn - 2) { mode = 2; outside = 1; }
        else if (iloc == nn - 3) { mode = 2; outside = 0; }
        else { mode = 0; outside = 0; }
        volatile T g0 = g[origin], g1 = g[origin + 1], g2 = g[origin + 2], g3 = g[origin + 3];
        volatile T h01 = g1 - g0, h12 = g2 - g1, h23 = g3 - g2;
        volatile T wa, wc, div0, wa1 = one, wc1 = one, div1 = one, gref, href;
        if (mode == 0) {
            volatile T r = h01 / h12;
            volatile T rp = r + one, pr = one + r;
            wa = r / rp; wc = one / pr; div0 = r;
            volatile T s = h23 / h12;
            volatile T ps = one + s, sp = s + one;
            wa1 = one / ps; wc1 = s / sp; div1 = s;
            gref = g1; href = h12;
        } else if (mode == 1) {
            volatile T q = h12 / h01;
            volatile T pq = one + q, qp = q + one;
            volatile T a = one / pq, c = q / qp;
            wa = c; wc = a;  // exchanged: see the header comment
            div0 = q;
            gref = g1; href = h01;
        } else {
            volatile T p = h12 / h23;
            volatile T pp1 = p + one, p1p = one + p;
            wa = p / pp1; wc = one / p1p; div0 = p;
            gref = g2; href = h23;
        }
        volatile T rdiv0 = one / div0, rdiv1 = one / div1, rhref = one / href;
        const int flags = mode | (outside << 2) | ((in_range(div0) && in_range(div1)) ? 8 : 0) | (in_range(href) ? 16 : 0);
        T fbits = T(0);
        memcpy(&fbits, &flags, sizeof(int));
        const T row[14] = {wa, wc, div0, rdiv0, wa1, wc1, div1, rdiv1, gref, href, rhref, fbits, T(0), T(0)};
        packed.insert(packed.end(), row, row + cubic_cell_row_stride(static_cast<int>(sizeof(T))));
    }
}

template <class T>
int rect_new(int method, const T* const* grids, const size_t* grid_lens, size_t ngrids, const T* vals, size_t nvals,
             int linearize, int vals_location, interpn_b200_interp** out) {
    if (!out) return INTERPN_B200_ERR_INVALID_ARG;
    *out = nullptr;
    if (!grids || !grid_lens) return INTERPN_B200_ERR_INVALID_ARG;
    // grids[i][1] is read by validation: guard the lengths first, like the reference's order.
    if (ngrids >= 1 && ngrids <= max_dims(method))
        for (size_t i = 0; i < ngrids; ++i)
            if (!grids[i] && grid_lens[i]) return INTERPN_B200_ERR_INVALID_ARG;
    int st = validate_rect_new<T>(method, grids, grid_lens, ngrids, nvals);
    if (st != INTERPN_B200_OK) return st;
    int sms = 0;
    st = require_device(&sms);
    if (st != INTERPN_B200_OK) return st;
    auto* h = new (std::nothrow) interpn_b200_interp();
    if (!h) return INTERPN_B200_ERR_TOO_LARGE;
    DeviceGrid& g = h->g;
    g.method = method;
    g.rect = 1;
    g.ndims = static_cast<int>(ngrids);
    g.linearize = linearize != 0;
    g.elem = sizeof(T);
    g.nvals = nvals;
    g.sm_count = sms;
    std::vector<T> packed;
    for (size_t d = 0; d < ngrids; ++d) {
        g.dim[d] = static_cast<int>(grid_lens[d]);
        g.axis_off[d] = static_cast<int>(packed.size());
        packed.insert(packed.end(), grids[d], grids[d] + grid_lens[d]);
    }
    g.grid_hash = fnv1a(fnv1a(14695981039346656037ull, g.dim, sizeof(int) * ngrids), packed.data(), packed.size() * sizeof(T));
    // Search accelerators (kernels.cuh rect_lower_bound), valid only for strictly increasing finite axes — the
    // reference itself checks just the first two nodes (multilinear/rectilinear.rs:194-198), and on anything else
    // the kernels keep the plain bisection. Per axis: reciprocal cell widths (f64: exact_div's divisor table) and
    // a bucket table lut[k] = partition_point(g < g0 + k*(g_last-g0)/nb), stored as ints in the same blob.
    bool sorted = true, widths_ok = true;  // every cell width within the guarded range of the division-free sequences
    for (size_t d = 0; d < ngrids && sorted; ++d) {
        for (size_t i = 0; i < grid_lens[d]; ++i) {
            const double v = static_cast<double>(grids[d][i]);
            if (!(v - v == 0.0)) sorted = false;  // NaN / inf
            if (i && !(grids[d][i] > grids[d][i - 1])) sorted = false;
            if (i) {
                const double w = static_cast<double>(grids[d][i]) - static_cast<double>(grids[d][i - 1]);
                if (sizeof(T) == 8 ? !(w >= 0x1p-300 && w < 0x1p301) : !(w >= 0x1p-60 && w < 0x1p60)) widths_ok = false;
            }
        }
    }
    // The bucket tables index with (x - g0) * (buckets / span) evaluated in T (kernels.cuh rect_lower_bound,
    // rect_cell_locate): the span and the scale must be finite, normal numbers IN T, or a computed bucket can land
    // arbitrarily far from the true one (f32 axes spanning more than FLT_MAX, or a subnormal span). Such axes keep the
    // plain bisection, like axes that are not strictly increasing.
    for (size_t d = 0; d < ngrids && sorted; ++d) {
        const size_t n = grid_lens[d];
        volatile T span_t = grids[d][n - 1] - grids[d][0];
        const double span = static_cast<double>(grids[d][n - 1]) - static_cast<double>(grids[d][0]);
        const double tmin = sizeof(T) == 8 ? 0x1p-1022 : 0x1p-126, tmax = sizeof(T) == 8 ? 0x1p1023 : 0x1p127;
        const double scale_hi = static_cast<double>(4 * n > 65536 ? 65536 : 4 * n) / span, scale_lo = 2.0 / span;
        const double st = static_cast<double>(span_t);
        if (!(st - st == 0.0) || !(st >= tmin) || !(scale_hi < tmax) || !(scale_lo >= tmin)) sorted = false;
    }
    g.rect_fast = sorted ? 1 : 0;
    g.rect_fast_div = sorted && widths_ok ? 1 : 0;
    if (sorted) {
        for (size_t d = 0; d < ngrids; ++d) {
            const size_t n = grid_lens[d];
            const double g0 = static_cast<double>(grids[d][0]), span = static_cast<double>(grids[d][n - 1]) - g0;
            auto append_ints = [&](const std::vector<int>& v) {
                const size_t slots = (v.size() * sizeof(int) + sizeof(T) - 1) / sizeof(T);
                const size_t at = packed.size();
                packed.resize(at + slots, T(0));
                memcpy(packed.data() + at, v.data(), v.size() * sizeof(int));
            };
            if (method == INTERPN_B200_LINEAR) {  // reciprocal cell widths: the divisor table of t = (x - g0)/(g1 - g0)
                g.rc_off[d] = static_cast<int>(packed.size());
                for (size_t i = 0; i + 1 < n; ++i) {
                    volatile T w = grids[d][i + 1] - grids[d][i];  // the kernels' divisor, rounded in T
                    packed.push_back(static_cast<T>(T(1) / w));
                }
            }
            if (method != INTERPN_B200_NEAREST) {
                // bucket table of the search (kernels.cuh rect_lower_bound): lut[k] = partition_point(g < edge_k)
                size_t nb = 2 * n;
                if (nb > 65536) nb = 65536;
                std::vector<int> lut(nb + 1);
                lut[0] = 0;
                lut[nb] = static_cast<int>(n);
                for (size_t k = 1; k < nb; ++k) {
                    const double edge = g0 + span * (static_cast<double>(k) / static_cast<double>(nb));
                    size_t lo = 0, hi = n;
                    while (lo < hi) {
                        const size_t mid = (lo + hi) / 2;
                        if (static_cast<double>(grids[d][mid]) < edge) lo = mid + 1;
                        else hi = mid;
                    }
                    lut[k] = static_cast<int>(lo);
                }
                g.lut_off[d] = static_cast<int>(packed.size());
                g.lut_nb[d] = static_cast<int>(nb);
                g.lut_scale[d] = static_cast<double>(nb) / span;
                append_ints(lut);
            }
        }
        g.axes_core = static_cast<int>(packed.size());  // what a kernel that does not use the cell tables stages
        for (size_t d = 0; d < ngrids; ++d) {
            const size_t n = grid_lens[d];
            const double g0 = static_cast<double>(grids[d][0]), span = static_cast<double>(grids[d][n - 1]) - g0;
            auto append_ints = [&](const std::vector<int>& v) {
                const size_t slots = (v.size() * sizeof(int) + sizeof(T) - 1) / sizeof(T);
                const size_t at = packed.size();
                packed.resize(at + slots, T(0));
                memcpy(packed.data() + at, v.data(), v.size() * sizeof(int));
            };
            if (method != INTERPN_B200_CUBIC) {
                // cell table of the multilinear / nearest kernels (kernels.cuh rect_cell_locate): four buckets per node,
                // clut[b] = the cell that contains the left edge of bucket b
                size_t nb = 4 * n;
                if (nb > 65536) nb = 65536;
                std::vector<int> clut(nb);
                for (size_t k = 0; k < nb; ++k) {
                    const double edge = g0 + span * (static_cast<double>(k) / static_cast<double>(nb));
                    size_t lo = 0, hi = n;  // number of nodes <= edge
                    while (lo < hi) {
                        const size_t mid = (lo + hi) / 2;
                        if (static_cast<double>(grids[d][mid]) <= edge) lo = mid + 1;
                        else hi = mid;
                    }
                    const long c = static_cast<long>(lo) - 1;
                    clut[k] = static_cast<int>(c < 0 ? 0 : (c > static_cast<long>(n) - 2 ? static_cast<long>(n) - 2 : c));
                }
                g.clut_off[d] = static_cast<int>(packed.size());
                g.clut_nb[d] = static_cast<int>(nb);
                g.clut_scale[d] = static_cast<double>(nb) / span;
                append_ints(clut);
            }
        }
        if (method != INTERPN_B200_CUBIC) g.rect_cell = 1;
        if (method == INTERPN_B200_NEAREST) g.rect_fast = 0;  // no bucket table is built for nearest
    }
    if (sorted && method == INTERPN_B200_CUBIC) {
        for (size_t d = 0; d < ngrids; ++d) {
            while (packed.size() * sizeof(T) % 16) packed.push_back(T(0));  // rows are read as 16-byte vectors
            g.ct_off[d] = static_cast<int>(packed.size());
            cubic_cell_table<T>(grids[d], grid_lens[d], packed);
        }
        g.rect_cubic_table = 1;
    }
    g.axes_total = static_cast<int>(packed.size());
    set_strides(g);
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(&g.axes, packed.size() * sizeof(T)));
        CUDA_TRY(cudaMemcpy(g.axes, packed.data(), packed.size() * sizeof(T), cudaMemcpyHostToDevice));
        int s = upload_vals(h, vals, vals_location);
        return s == INTERPN_B200_OK ? finish_new(h) : s;
    };
    st = body();
    if (st != INTERPN_B200_OK) {
        interpn_b200_interp_free(h);
        return st;
    }
    *out = h;
    return INTERPN_B200_OK;
}

// ---- several devices behind one host-buffer call (host_exec.cuh) -----------------------------------

constexpr size_t kMultiDeviceMinPoints = size_t(1) << 22;  // below this one GPU finishes before a second one has its grid
std::atomic<int> g_host_devices{-1};                       // -1: INTERPN_B200_HOST_DEVICES or every visible device

int host_device_limit() {
    int want = g_host_devices.load(std::memory_order_relaxed);
    if (want < 0) {
        const char* e = getenv("INTERPN_B200_HOST_DEVICES");
        want = e && *e ? atoi(e) : 0;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return want <= 0 || want > count ? count : want;
}

void free_replicas(interpn_b200_interp* h) {
    int home = 0;
    cudaGetDevice(&home);
    for (auto* r : h->replicas) {
        cudaSetDevice(r->slots.device);
        r->slots.destroy();
        if (r->g.vals) cudaFree(r->g.vals);
        if (r->g.win) cudaFree(r->g.win);
        if (r->g.axes) cudaFree(r->g.axes);
        delete r;
    }
    h->replicas.clear();
    cudaSetDevice(home);
}

// Replicates the resident grid (vals, its window copy, the axes blob) onto the next `want - 1` sm_100 devices after the
// interpolator's own: device-to-device copies (NVLink peer copies where the devices are peers), once per interpolator.
int ensure_replicas(interpn_b200_interp* h, int want) {
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    const DeviceGrid& g = h->g;
    const size_t bytes = g.nvals * static_cast<size_t>(g.elem);
    for (int k = 1; k < count && static_cast<int>(h->replicas.size()) + 1 < want; ++k) {
        const int dev = (h->device + k) % count;
        bool have = false;
        for (auto* r : h->replicas) have = have || r->slots.device == dev;
        if (have) continue;
        int major = 0, sms = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        if (major != 10) continue;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        auto* r = new (std::nothrow) interpn_b200_interp::Replica();
        if (!r) return INTERPN_B200_ERR_TOO_LARGE;
        r->g = g;
        r->g.vals = r->g.win = r->g.axes = nullptr;
        r->g.sm_count = sms;
        r->slots.device = dev;
        h->replicas.push_back(r);  // owned from here on (freed with the interpolator even if a copy below fails)
        CUDA_TRY(cudaSetDevice(dev));
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dev, h->device) == cudaSuccess && can) {
            cudaDeviceEnablePeerAccess(h->device, 0);  // already-enabled is fine
            cudaGetLastError();
        }
        CUDA_TRY(cudaMalloc(&r->g.vals, bytes ? bytes : 1));
        CUDA_TRY(cudaMemcpyPeer(r->g.vals, dev, g.vals, h->device, bytes));
        if (g.win) {
            CUDA_TRY(cudaMalloc(&r->g.win, g.win_bytes));
            CUDA_TRY(cudaMemcpyPeer(r->g.win, dev, g.win, h->device, g.win_bytes));
        }
        if (g.axes) {
            CUDA_TRY(cudaMalloc(&r->g.axes, static_cast<size_t>(g.axes_total) * g.elem));
            CUDA_TRY(cudaMemcpyPeer(r->g.axes, dev, g.axes, h->device, static_cast<size_t>(g.axes_total) * g.elem));
        }
        CUDA_TRY(cudaDeviceSynchronize());
    }
    CUDA_TRY(cudaSetDevice(h->device));
    return INTERPN_B200_OK;
}

template <class T>
int eval_host(interpn_b200_interp* h, const T* const* obs, const size_t* obs_lens, size_t nobs, T* out, size_t nout,
              size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    if (!h || (nobs && (!obs || !obs_lens))) return INTERPN_B200_ERR_INVALID_ARG;
    if (h->g.elem != static_cast<int>(sizeof(T))) return INTERPN_B200_ERR_INVALID_ARG;
    int st = validate_interp(static_cast<size_t>(h->g.ndims), obs_lens, nobs, nout);
    if (st != INTERPN_B200_OK) return st;
    if (nout == 0) return INTERPN_B200_OK;
    if (!out) return INTERPN_B200_ERR_INVALID_ARG;
    const void* in_host[kMaxNd];
    for (size_t j = 0; j < nobs; ++j) in_host[j] = obs[j];
    std::lock_guard<std::mutex> lk(h->host_mu);
    int home = 0;
    CUDA_TRY(cudaGetDevice(&home));
    CUDA_TRY(cudaSetDevice(h->device));
    // Devices of this call: the interpolator's own, plus replicas on the other devices the process may use
    // (interpn_b200_set_host_devices) when the batch is large enough to feed them.
    DeviceSlots* devs[64];
    const DeviceGrid* grids[64];
    int ndev = 1;
    devs[0] = &h->home;
    grids[0] = &h->g;
    h->home.device = h->device;
    const int want = host_device_limit();
    if (want > 1 && nout >= kMultiDeviceMinPoints) {
        int st2 = ensure_replicas(h, want);
        if (st2 != INTERPN_B200_OK) return st2;
        for (auto* r : h->replicas) {
            if (ndev >= want || ndev >= 64) break;
            devs[ndev] = &r->slots;
            grids[ndev] = &r->g;
            ++ndev;
        }
    }
    st = run_host_batch(
        devs, ndev, in_host, static_cast<int>(nobs), out, nout, sizeof(T),
        [&](int di, void* const* in_dev, void* out_dev, size_t cnt, unsigned long long* flag, unsigned long long base,
            cudaStream_t s) {
            const T* o[kMaxNd];
            for (size_t j = 0; j < nobs; ++j) o[j] = static_cast<const T*>(in_dev[j]);
            return launch_eval<T>(*grids[di], o, cnt, static_cast<T*>(out_dev), flag, base, s);
        },
        first_bad);
    cudaSetDevice(home);
    return st;
}

template <class T>
int eval_device(interpn_b200_interp* h, const T* const* obs, size_t nobs, size_t n, T* out, void* stream) {
    if (!h || (nobs && !obs)) return INTERPN_B200_ERR_INVALID_ARG;
    if (h->g.elem != static_cast<int>(sizeof(T))) return INTERPN_B200_ERR_INVALID_ARG;
    if (nobs != static_cast<size_t>(h->g.ndims)) return INTERPN_B200_ERR_DIM_MISMATCH;
    if (n == 0) return INTERPN_B200_OK;
    if (!out) return INTERPN_B200_ERR_INVALID_ARG;
    int cur = -1;
    CUDA_TRY(cudaGetDevice(&cur));
    if (cur != h->device) {  // the grid, the stream and the buffers all belong to the interpolator's device
        snprintf(t_detail, sizeof(t_detail), "interpolator lives on device %d but device %d is current", h->device, cur);
        return INTERPN_B200_ERR_INVALID_ARG;
    }
    CUDA_TRY(launch_eval<T>(h->g, obs, n, out, h->first_bad_dev, 0ull, static_cast<cudaStream_t>(stream)));
    return INTERPN_B200_OK;
}

// Fused multi-field evaluation (launch_fields.cu): every interpolator of `hs` must have been built
// over the same grid with the same method; the cell location runs once per point. Unrepresentable points are latched on
// hs[0] (interpn_b200_interp_status(hs[0], ...)). Combinations without a fused kernel (multicubic, N > 6, grids beyond
// L2) are evaluated field by field through the ordinary path, so the call is always valid and always bit-identical to
// `nfields` separate calls.
template <class T>
int eval_fields_device(interpn_b200_interp* const* hs, size_t nfields, const T* const* obs, size_t nobs, size_t n,
                       T* const* outs, void* stream) {
    if (!hs || !outs || nfields == 0 || (nobs && !obs)) return INTERPN_B200_ERR_INVALID_ARG;
    for (size_t k = 0; k < nfields; ++k) {
        if (!hs[k] || (n && !outs[k])) return INTERPN_B200_ERR_INVALID_ARG;
        if (hs[k]->g.elem != static_cast<int>(sizeof(T))) return INTERPN_B200_ERR_INVALID_ARG;
    }
    const DeviceGrid& g0 = hs[0]->g;
    if (nobs != static_cast<size_t>(g0.ndims)) return INTERPN_B200_ERR_DIM_MISMATCH;
    for (size_t k = 1; k < nfields; ++k) {
        const DeviceGrid& g = hs[k]->g;
        if (g.method != g0.method || g.rect != g0.rect || g.ndims != g0.ndims || g.linearize != g0.linearize ||
            g.nvals != g0.nvals || g.grid_hash != g0.grid_hash)
            return INTERPN_B200_ERR_DIM_MISMATCH;  // not the same grid
    }
    if (n == 0) return INTERPN_B200_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (size_t k0 = 0; k0 < nfields; k0 += kMaxFields) {
        const int nf = static_cast<int>(nfields - k0 < static_cast<size_t>(kMaxFields) ? nfields - k0 : kMaxFields);
        const DeviceGrid* grids[kMaxFields];
        for (int k = 0; k < nf; ++k) grids[k] = &hs[k0 + k]->g;
        cudaError_t e = launch_eval_fields<T>(grids, nf, obs, n, outs + k0, hs[0]->first_bad_dev, s);
        if (e == cudaErrorNotSupported) {
            cudaGetLastError();
            for (int k = 0; k < nf; ++k)
                CUDA_TRY(launch_eval<T>(hs[k0 + k]->g, obs, n, outs[k0 + k], hs[0]->first_bad_dev, 0ull, s));
        } else {
            CUDA_TRY(e);
        }
    }
    return INTERPN_B200_OK;
}

// ---- one-shot wrappers: dispatcher-level checks of `interpn(...)`, then new + interp ----------

template <class T>
int oneshot_regular(int method, const size_t* dims, size_t ndims, const T* starts, size_t nstarts, const T* steps,
                    size_t nsteps, const T* vals, size_t nvals, int linearize, const T* const* obs,
                    const size_t* obs_lens, size_t nobs, T* out, size_t nout, size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    // multilinear/regular.rs:60-62 and nearest/regular.rs:50-52 check lengths before matching on ndims;
    // multicubic/regular.rs:64-65 matches first.
    if (method != INTERPN_B200_CUBIC && (nstarts != ndims || nsteps != ndims || nobs != ndims))
        return INTERPN_B200_ERR_DIM_MISMATCH;
    // All of the reference's checks run on the host, in its order (new(), then interp()), before any
    // device work, so argument errors are reported identically with or without a GPU.
    if (!dims || !starts || !steps) return INTERPN_B200_ERR_INVALID_ARG;
    int st = validate_regular_new<T>(method, dims, ndims, nstarts, steps, nsteps, nvals);
    if (st != INTERPN_B200_OK) return st;
    if (nobs && (!obs || !obs_lens)) return INTERPN_B200_ERR_INVALID_ARG;
    st = validate_interp(ndims, obs_lens, nobs, nout);
    if (st != INTERPN_B200_OK) return st;
    interpn_b200_interp* h = nullptr;
    st = regular_new<T>(method, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals, linearize,
                        INTERPN_B200_VALS_HOST, &h);
    if (st != INTERPN_B200_OK) return st;
    st = eval_host<T>(h, obs, obs_lens, nobs, out, nout, first_bad);
    interpn_b200_interp_free(h);
    return st;
}

template <class T>
int oneshot_rect(int method, const T* const* grids, const size_t* grid_lens, size_t ngrids, const T* vals,
                 size_t nvals, int linearize, const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,
                 size_t nout) {
    // multilinear/rectilinear.rs:58-61, nearest/rectilinear.rs:45-48
    if (method != INTERPN_B200_CUBIC && nobs != ngrids) return INTERPN_B200_ERR_DIM_MISMATCH;
    if (!grids || !grid_lens) return INTERPN_B200_ERR_INVALID_ARG;
    if (ngrids >= 1 && ngrids <= max_dims(method))
        for (size_t i = 0; i < ngrids; ++i)
            if (!grids[i] && grid_lens[i]) return INTERPN_B200_ERR_INVALID_ARG;
    int st = validate_rect_new<T>(method, grids, grid_lens, ngrids, nvals);
    if (st != INTERPN_B200_OK) return st;
    if (nobs && (!obs || !obs_lens)) return INTERPN_B200_ERR_INVALID_ARG;
    st = validate_interp(ngrids, obs_lens, nobs, nout);
    if (st != INTERPN_B200_OK) return st;
    interpn_b200_interp* h = nullptr;
    st = rect_new<T>(method, grids, grid_lens, ngrids, vals, nvals, linearize, INTERPN_B200_VALS_HOST, &h);
    if (st != INTERPN_B200_OK) return st;
    st = eval_host<T>(h, obs, obs_lens, nobs, out, nout, nullptr);
    interpn_b200_interp_free(h);
    return st;
}

// ---- check_bounds --------------------------------------------------------------------------------

template <class T>
int check_bounds_axes(const T* lo, const T* hi, size_t ndims, const T* const* obs, const size_t* obs_lens, T atol,
                      uint8_t* out) {
    int st = require_device();
    if (st != INTERPN_B200_OK) return st;
    int* flags_dev = nullptr;
    CUDA_TRY(cudaMalloc(&flags_dev, sizeof(int) * ndims));
    HostPipeline pipe;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMemset(flags_dev, 0, sizeof(int) * ndims));
        for (size_t d = 0; d < ndims; ++d) {
            const void* in_host[1] = {obs[d]};
            int* flag = flags_dev + d;
            const T l = lo[d], h = hi[d];
            DeviceSlots* devs[1] = {&pipe.dev};
            int s = run_host_batch(
                devs, 1, in_host, 1, nullptr, obs_lens[d], sizeof(T),
                [&](int, void* const* in_dev, void*, size_t cnt, unsigned long long*, unsigned long long, cudaStream_t stream) {
                    return launch_check_bounds<T>(static_cast<const T*>(in_dev[0]), cnt, l, h, atol, flag, stream);
                },
                nullptr);
            if (s != INTERPN_B200_OK) return s;
        }
        std::vector<int> flags(ndims);
        CUDA_TRY(cudaMemcpy(flags.data(), flags_dev, sizeof(int) * ndims, cudaMemcpyDeviceToHost));
        for (size_t d = 0; d < ndims; ++d) out[d] = flags[d] ? 1 : 0;
        return INTERPN_B200_OK;
    };
    st = body();
    cudaFree(flags_dev);
    return st;
}

template <class T>
T host_min(T a, T b) {  // f64::min semantics: the non-NaN operand wins
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}
template <class T>
T host_max(T a, T b) {
    if (a != a) return b;
    if (b != b) return a;
    return a > b ? a : b;
}

template <class T>
int check_bounds_regular(const size_t* dims, size_t ndims, const T* starts, size_t nstarts, const T* steps,
                         size_t nsteps, const T* const* obs, const size_t* obs_lens, size_t nobs, T atol, uint8_t* out,
                         size_t nout) {
    if (!(nobs == ndims && nout == ndims)) return INTERPN_B200_ERR_DIM_MISMATCH;  // multilinear/regular.rs:153-156
    if (nstarts < ndims || nsteps < ndims) return INTERPN_B200_ERR_DIM_MISMATCH;  // reference would panic on indexing
    if (ndims == 0) return INTERPN_B200_OK;
    if (!dims || !starts || !steps || !obs || !obs_lens || !out) return INTERPN_B200_ERR_INVALID_ARG;
    std::vector<T> lo(ndims), hi(ndims);
    for (size_t i = 0; i < ndims; ++i) {
        // multilinear/regular.rs:159-166, in T arithmetic, product and sum rounded separately
        volatile T prod = steps[i] * static_cast<T>(dims[i] - 1);
        volatile T last = starts[i] + prod;
        lo[i] = host_min<T>(starts[i], last);
        hi[i] = host_max<T>(starts[i], last);
    }
    return check_bounds_axes<T>(lo.data(), hi.data(), ndims, obs, obs_lens, atol, out);
}

template <class T>
int check_bounds_rect(const T* const* grids, const size_t* grid_lens, size_t ngrids, const T* const* obs,
                      const size_t* obs_lens, size_t nobs, T atol, uint8_t* out, size_t nout) {
    if (!(nobs == ngrids && nout == ngrids)) return INTERPN_B200_ERR_DIM_MISMATCH;  // multilinear/rectilinear.rs:115-118
    if (ngrids == 0) return INTERPN_B200_OK;
    if (!grids || !grid_lens || !obs || !obs_lens || !out) return INTERPN_B200_ERR_INVALID_ARG;
    for (size_t i = 0; i < ngrids; ++i)
        if (grid_lens[i] == 0) return INTERPN_B200_ERR_DIM_MISMATCH;
    std::vector<T> lo(ngrids), hi(ngrids);
    for (size_t i = 0; i < ngrids; ++i) {
        lo[i] = grids[i][0];
        hi[i] = grids[i][grid_lens[i] - 1];
    }
    return check_bounds_axes<T>(lo.data(), hi.data(), ngrids, obs, obs_lens, atol, out);
}

// ---- one_dim -------------------------------------------------------------------------------------

template <class T>
int one_dim_host(int kind, bool rect, T start, T step, const T* grid, size_t ngrid, const T* vals, size_t nvals,
                 const T* locs, size_t nlocs, T* out, size_t nout, size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    if (kind < 0 || kind > 4) return INTERPN_B200_ERR_INVALID_ARG;
    if (rect) {
        if (ngrid != nvals || ngrid < 2) return INTERPN_B200_ERR_LENGTH_MISMATCH;  // one_dim/mod.rs:148-152
    } else if (nvals < 2) {
        return INTERPN_B200_ERR_LENGTH_MISMATCH;  // reference underflows `len - 2` and panics
    }
    if (nvals > size_t(INT_MAX) - 8) return INTERPN_B200_ERR_TOO_LARGE;
    if (nlocs != nout) return INTERPN_B200_ERR_LENGTH_MISMATCH;  // one_dim/mod.rs:52-54
    if (!vals || (rect && !grid) || (nlocs && (!locs || !out))) return INTERPN_B200_ERR_INVALID_ARG;
    int st = require_device();
    if (st != INTERPN_B200_OK) return st;
    T* vals_dev = nullptr;
    T* grid_dev = nullptr;
    HostPipeline pipe;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc(&vals_dev, nvals * sizeof(T)));
        CUDA_TRY(cudaMemcpy(vals_dev, vals, nvals * sizeof(T), cudaMemcpyHostToDevice));
        if (rect) {
            CUDA_TRY(cudaMalloc(&grid_dev, ngrid * sizeof(T)));
            CUDA_TRY(cudaMemcpy(grid_dev, grid, ngrid * sizeof(T), cudaMemcpyHostToDevice));
        }
        const void* in_host[1] = {locs};
        DeviceSlots* devs[1] = {&pipe.dev};
        int s = run_host_batch(
            devs, 1, in_host, 1, out, nlocs, sizeof(T),
            [&](int, void* const* in_dev, void* out_dev, size_t cnt, unsigned long long* flag, unsigned long long base,
                cudaStream_t stream) {
                return launch_one_dim<T>(kind, rect, start, step, grid_dev, vals_dev, nvals,
                                         static_cast<const T*>(in_dev[0]), cnt, static_cast<T*>(out_dev), flag, base,
                                         stream);
            },
            first_bad);
        return s == INTERPN_B200_ERR_UNREPRESENTABLE ? INTERPN_B200_ERR_UNREPRESENTABLE_NUM : s;
    };
    st = body();
    if (vals_dev) cudaFree(vals_dev);
    if (grid_dev) cudaFree(grid_dev);
    return st;
}

}  // namespace

// =================================================================================================
// extern "C"
// =================================================================================================

extern "C" {

const char* interpn_b200_strerror(int status) {
    switch (status) {
        case INTERPN_B200_OK: return "";
        case INTERPN_B200_ERR_DIM_MISMATCH: return "Dimension mismatch";
        case INTERPN_B200_ERR_MIN_TWO: return "All grids must have at least two entries";
        case INTERPN_B200_ERR_MIN_2: return "All grids must have at least 2 entries";
        case INTERPN_B200_ERR_MIN_FOUR: return "All grids must have at least four entries";
        case INTERPN_B200_ERR_MIN_4: return "All grids must have at least 4 entries";
        case INTERPN_B200_ERR_NOT_MONOTONIC: return "All grids must be monotonically increasing";
        case INTERPN_B200_ERR_UNREPRESENTABLE: return "Unrepresentable coordinate value";
        case INTERPN_B200_ERR_MAXDIM_8:
            return "Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions.";
        case INTERPN_B200_ERR_MAXDIM_6: return "Dimension exceeds maximum (6).";
        case INTERPN_B200_ERR_LENGTH_MISMATCH: return "Length mismatch";
        case INTERPN_B200_ERR_UNREPRESENTABLE_NUM: return "Unrepresentable number";
        case INTERPN_B200_ERR_CUDA: return "CUDA runtime failure";
        case INTERPN_B200_ERR_NO_DEVICE: return "No usable sm_100 CUDA device (this library has no CPU fallback)";
        case INTERPN_B200_ERR_INVALID_ARG: return "Invalid argument";
        case INTERPN_B200_ERR_TOO_LARGE: return "Grid too large for device memory or index range";
        default: return "Unknown status";
    }
}

const char* interpn_b200_last_error_detail(void) { return t_detail; }

int interpn_b200_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int interpn_b200_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return INTERPN_B200_OK;
}

int interpn_b200_set_host_devices(int n) {
    if (n < 0) return INTERPN_B200_ERR_INVALID_ARG;
    g_host_devices.store(n, std::memory_order_relaxed);
    return INTERPN_B200_OK;
}
int interpn_b200_host_devices(void) { return host_device_limit(); }
int interpn_b200_copy_threads(void) { return CopyPool::get().threads() + 1; }

int interpn_b200_arithmetic(void) { return IB200_ARITH_FMA; }
uint64_t interpn_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
uint64_t interpn_b200_swept_launch_count(void) { return g_swept.load(std::memory_order_relaxed); }

int interpn_b200_sm_count(void) {
    int sms = 0;
    return require_device(&sms) == INTERPN_B200_OK ? sms : 0;
}

#define INTERPN_B200_DEFINE(SUFFIX, T)                                                                                 \
    int interpn_b200_linear_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,        \
                                             const T* steps, size_t nsteps, const T* vals, size_t nvals,               \
                                             const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,         \
                                             size_t nout, size_t* first_bad) {                                         \
        return oneshot_regular<T>(INTERPN_B200_LINEAR, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals, 0,    \
                                  obs, obs_lens, nobs, out, nout, first_bad);                                          \
    }                                                                                                                  \
    int interpn_b200_linear_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,        \
                                                 const T* vals, size_t nvals, const T* const* obs,                     \
                                                 const size_t* obs_lens, size_t nobs, T* out, size_t nout) {           \
        return oneshot_rect<T>(INTERPN_B200_LINEAR, grids, grid_lens, ngrids, vals, nvals, 0, obs, obs_lens, nobs,     \
                               out, nout);                                                                             \
    }                                                                                                                  \
    int interpn_b200_cubic_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,         \
                                            const T* steps, size_t nsteps, const T* vals, size_t nvals,                \
                                            int linearize_extrapolation, const T* const* obs,                          \
                                            const size_t* obs_lens, size_t nobs, T* out, size_t nout,                  \
                                            size_t* first_bad) {                                                       \
        return oneshot_regular<T>(INTERPN_B200_CUBIC, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals,        \
                                  linearize_extrapolation, obs, obs_lens, nobs, out, nout, first_bad);                 \
    }                                                                                                                  \
    int interpn_b200_cubic_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,         \
                                                const T* vals, size_t nvals, int linearize_extrapolation,              \
                                                const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,      \
                                                size_t nout) {                                                         \
        return oneshot_rect<T>(INTERPN_B200_CUBIC, grids, grid_lens, ngrids, vals, nvals, linearize_extrapolation,     \
                               obs, obs_lens, nobs, out, nout);                                                        \
    }                                                                                                                  \
    int interpn_b200_nearest_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,       \
                                              const T* steps, size_t nsteps, const T* vals, size_t nvals,              \
                                              const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,        \
                                              size_t nout, size_t* first_bad) {                                        \
        return oneshot_regular<T>(INTERPN_B200_NEAREST, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals, 0,   \
                                  obs, obs_lens, nobs, out, nout, first_bad);                                          \
    }                                                                                                                  \
    int interpn_b200_nearest_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,       \
                                                  const T* vals, size_t nvals, const T* const* obs,                    \
                                                  const size_t* obs_lens, size_t nobs, T* out, size_t nout) {          \
        return oneshot_rect<T>(INTERPN_B200_NEAREST, grids, grid_lens, ngrids, vals, nvals, 0, obs, obs_lens, nobs,    \
                               out, nout);                                                                             \
    }                                                                                                                  \
    int interpn_b200_check_bounds_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts,                  \
                                                   size_t nstarts, const T* steps, size_t nsteps,                      \
                                                   const T* const* obs, const size_t* obs_lens, size_t nobs, T atol,   \
                                                   uint8_t* out, size_t nout) {                                        \
        return check_bounds_regular<T>(dims, ndims, starts, nstarts, steps, nsteps, obs, obs_lens, nobs, atol, out,    \
                                       nout);                                                                          \
    }                                                                                                                  \
    int interpn_b200_check_bounds_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens,                 \
                                                       size_t ngrids, const T* const* obs, const size_t* obs_lens,     \
                                                       size_t nobs, T atol, uint8_t* out, size_t nout) {               \
        return check_bounds_rect<T>(grids, grid_lens, ngrids, obs, obs_lens, nobs, atol, out, nout);                   \
    }                                                                                                                  \
    int interpn_b200_one_dim_regular_##SUFFIX(int kind, T start, T step, const T* vals, size_t nvals, const T* locs,   \
                                              size_t nlocs, T* out, size_t nout, size_t* first_bad) {                  \
        return one_dim_host<T>(kind, false, start, step, nullptr, 0, vals, nvals, locs, nlocs, out, nout, first_bad);  \
    }                                                                                                                  \
    int interpn_b200_one_dim_rectilinear_##SUFFIX(int kind, const T* grid, size_t ngrid, const T* vals, size_t nvals,  \
                                                  const T* locs, size_t nlocs, T* out, size_t nout) {                  \
        return one_dim_host<T>(kind, true, T(0), T(0), grid, ngrid, vals, nvals, locs, nlocs, out, nout, nullptr);     \
    }                                                                                                                  \
    int interpn_b200_regular_new_##SUFFIX(int method, const size_t* dims, size_t ndims, const T* starts,               \
                                          size_t nstarts, const T* steps, size_t nsteps, const T* vals, size_t nvals,  \
                                          int linearize_extrapolation, int vals_location,                              \
                                          interpn_b200_interp** out_interp) {                                          \
        return regular_new<T>(method, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals,                        \
                              linearize_extrapolation, vals_location, out_interp);                                     \
    }                                                                                                                  \
    int interpn_b200_rectilinear_new_##SUFFIX(int method, const T* const* grids, const size_t* grid_lens,              \
                                              size_t ngrids, const T* vals, size_t nvals,                              \
                                              int linearize_extrapolation, int vals_location,                          \
                                              interpn_b200_interp** out_interp) {                                      \
        return rect_new<T>(method, grids, grid_lens, ngrids, vals, nvals, linearize_extrapolation, vals_location,      \
                           out_interp);                                                                                \
    }                                                                                                                  \
    int interpn_b200_interp_eval_host_##SUFFIX(interpn_b200_interp* interp, const T* const* obs,                       \
                                               const size_t* obs_lens, size_t nobs, T* out, size_t nout,               \
                                               size_t* first_bad) {                                                    \
        return eval_host<T>(interp, obs, obs_lens, nobs, out, nout, first_bad);                                        \
    }                                                                                                                  \
    int interpn_b200_interp_eval_device_##SUFFIX(interpn_b200_interp* interp, const T* const* obs, size_t nobs,        \
                                                 size_t n, T* out, void* stream) {                                     \
        return eval_device<T>(interp, obs, nobs, n, out, stream);                                                      \
    }                                                                                                                  \
    int interpn_b200_interp_eval_fields_device_##SUFFIX(interpn_b200_interp* const* interps, size_t nfields,           \
                                                        const T* const* obs, size_t nobs, size_t n, T* const* outs,    \
                                                        void* stream) {                                                \
        return eval_fields_device<T>(interps, nfields, obs, nobs, n, outs, stream);                                    \
    }

INTERPN_B200_DEFINE(f64, double)
INTERPN_B200_DEFINE(f32, float)

int interpn_b200_interp_status(interpn_b200_interp* interp, void* stream, size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    if (!interp) return INTERPN_B200_ERR_INVALID_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned long long flag = ~0ull;
    CUDA_TRY(cudaMemcpyAsync(&flag, interp->first_bad_dev, sizeof(flag), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemsetAsync(interp->first_bad_dev, 0xff, sizeof(flag), s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag == ~0ull) return INTERPN_B200_OK;
    if (first_bad) *first_bad = static_cast<size_t>(flag);
    return INTERPN_B200_ERR_UNREPRESENTABLE;
}

void* interpn_b200_interp_vals_ptr(interpn_b200_interp* interp) { return interp ? interp->g.vals : nullptr; }

int interpn_b200_interp_vals_updated(interpn_b200_interp* interp, void* stream) {
    if (!interp) return INTERPN_B200_ERR_INVALID_ARG;
    CUDA_TRY(launch_build_window(interp->g, static_cast<cudaStream_t>(stream)));
    std::lock_guard<std::mutex> lk(interp->host_mu);
    free_replicas(interp);  // stale copies of the old values: rebuilt on the next multi-device call
    return INTERPN_B200_OK;
}
size_t interpn_b200_interp_vals_len(const interpn_b200_interp* interp) { return interp ? interp->g.nvals : 0; }
size_t interpn_b200_interp_elem_size(const interpn_b200_interp* interp) { return interp ? interp->g.elem : 0; }
size_t interpn_b200_interp_ndims(const interpn_b200_interp* interp) { return interp ? interp->g.ndims : 0; }

void interpn_b200_interp_free(interpn_b200_interp* interp) {
    if (!interp) return;
    free_replicas(interp);
    int home = 0;
    cudaGetDevice(&home);
    cudaSetDevice(interp->device);
    interp->home.destroy();
    if (interp->g.vals) cudaFree(interp->g.vals);
    if (interp->g.win) cudaFree(interp->g.win);
    if (interp->g.axes) cudaFree(interp->g.axes);
    if (interp->first_bad_dev) cudaFree(interp->first_bad_dev);
    cudaSetDevice(home);
    delete interp;
}

}  // extern "C"
