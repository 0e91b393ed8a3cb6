// oracle_capi.cpp — extern "C" face of the CPU ORACLE (test infrastructure, NOT product code).
//
// Batch drivers + argument validation restating the reference's `interpn(...)` dispatchers:
//   multilinear/regular.rs:51-117, multilinear/rectilinear.rs:49-83,
//   multicubic/regular.rs:52-136,  multicubic/rectilinear.rs:54-104,
//   nearest/regular.rs:41-101,     nearest/rectilinear.rs:39-66,
//   one_dim/mod.rs:41-61, multilinear/regular.rs:145-182, multilinear/rectilinear.rs:109-134.
// Loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
//
// Every slice of the Rust signatures is passed as (pointer, length) so the reference's own
// length checks can be restated. Where the reference would *panic* (`try_into().unwrap()` on a
// wrong-length slice) the oracle returns DIM_MISMATCH instead.
//
// Threading: nthreads == 1 reproduces the reference's serial loop exactly, including
// "stop at the first failing point, leave later outputs untouched"
// (multilinear/regular.rs:276-280). nthreads > 1 splits the batch into contiguous chunks
// (std::thread) for the all-cores CPU baseline; each chunk stops at its own first failure and the
// minimum failing index is reported.
#include "interpn_oracle.hpp"

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

using namespace oracle;

namespace {

template <class F>
int run_batch(size_t n, int nthreads, size_t* first_bad, F&& one) {
    // `one(i)` returns false if point i is unrepresentable.
    size_t bad = SIZE_MAX;
    if (nthreads <= 1) {
        for (size_t i = 0; i < n; ++i) {
            if (!one(i)) { bad = i; break; }
        }
    } else {
        std::mutex mu;
        std::vector<std::thread> pool;
        const size_t nt = static_cast<size_t>(nthreads);
        for (size_t tid = 0; tid < nt; ++tid) {
            pool.emplace_back([&, tid] {
                size_t lo = n / nt * tid + std::min<size_t>(tid, n % nt);
                size_t hi = lo + n / nt + (tid < n % nt ? 1 : 0);
                size_t mybad = SIZE_MAX;
                for (size_t i = lo; i < hi; ++i) {
                    if (!one(i)) { mybad = i; break; }
                }
                if (mybad != SIZE_MAX) {
                    std::lock_guard<std::mutex> g(mu);
                    bad = std::min(bad, mybad);
                }
            });
        }
        for (auto& t : pool) t.join();
    }
    if (first_bad) *first_bad = bad;
    return bad == SIZE_MAX ? OK : UNREPRESENTABLE;
}

inline bool lens_match(const size_t* obs_lens, size_t nobs, size_t nout) {
    for (size_t j = 0; j < nobs; ++j)
        if (obs_lens[j] != nout) return false;
    return true;
}

inline size_t product(const size_t* d, size_t n) {
    size_t p = 1;
    for (size_t i = 0; i < n; ++i) p *= d[i];
    return p;
}

// order: 0 = dispatch like the reference (flattened for N<=flat_max, recursive above),
//        1 = force flattened evaluation order, 2 = force recursive evaluation order.
inline bool use_recursive(int order, size_t ndims, size_t flat_max) {
    if (order == 1) return false;
    if (order == 2) return true;
    return ndims > flat_max;
}

template <class T>
int regular_impl(int method, const size_t* dims, size_t ndims, const T* starts, size_t nstarts, const T* steps,
                 size_t nsteps, const T* vals, size_t nvals, int linearize, const T* const* obs,
                 const size_t* obs_lens, size_t nobs, T* out, size_t nout, int fma, int order, int nthreads,
                 size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    const bool cubic = method == 1, nearest = method == 2;
    // Dispatcher-level checks.
    if (!cubic) {  // multilinear/regular.rs:60-62, nearest/regular.rs:50-52
        if (nstarts != ndims || nsteps != ndims || nobs != ndims) return DIM_MISMATCH;
    }
    const size_t maxdims = nearest ? 6 : 8;
    if (ndims < 1 || ndims > maxdims) return nearest ? MAXDIM_6 : MAXDIM_8;
    if (cubic && (nstarts != ndims || nsteps != ndims)) return DIM_MISMATCH;  // reference panics for N<=4
    // Struct::new
    if (nvals != product(dims, ndims)) return DIM_MISMATCH;
    const size_t mindim = cubic ? 4 : 2;
    for (size_t i = 0; i < ndims; ++i)
        if (dims[i] < mindim) return cubic ? MIN_FOUR : MIN_TWO;
    for (size_t i = 0; i < nsteps; ++i)
        if (!(steps[i] > T(0))) return NOT_MONOTONIC;
    // Struct::interp
    if (nobs != ndims) return DIM_MISMATCH;  // cubic N<=4: reference panics
    if (!lens_match(obs_lens, nobs, nout)) return DIM_MISMATCH;

    const int n = static_cast<int>(ndims);
    const bool lin = linearize != 0;
    auto gather = [&](size_t i, T* x) {
        for (int j = 0; j < n; ++j) x[j] = obs[j][i];
    };
#define ORACLE_RUN(EXPR) \
    return run_batch(nout, nthreads, first_bad, [&](size_t i) { T x[MAXDIMS]; gather(i, x); return (EXPR); })
// The reference's dispatch order with the dimensionality as a compile-time constant, like its const-generic structs
// (order == 0 only; the forced orders keep the runtime-N twins so the tests can compare the two).
#define ORACLE_RUN_N(NC, FN, ...)                                                                       \
    return run_batch(nout, nthreads, first_bad, [&](size_t i) {                                         \
        T x[MAXDIMS];                                                                                   \
        for (int j = 0; j < NC; ++j) x[j] = obs[j][i];                                                  \
        return fma ? FN<true, false, NC>(__VA_ARGS__) : FN<false, false, NC>(__VA_ARGS__);              \
    })
    if (order == 0 && method == 0) {
        switch (n) {
            case 1: ORACLE_RUN_N(1, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            case 2: ORACLE_RUN_N(2, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            case 3: ORACLE_RUN_N(3, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            case 4: ORACLE_RUN_N(4, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            case 5: ORACLE_RUN_N(5, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            case 6: ORACLE_RUN_N(6, linear_regular_one, n, dims, starts, steps, vals, x, out[i]);
            default: break;
        }
    }
    if (order == 0 && cubic) {
        switch (n) {
            case 1: ORACLE_RUN_N(1, cubic_regular_one, n, dims, starts, steps, vals, lin, x, out[i]);
            case 2: ORACLE_RUN_N(2, cubic_regular_one, n, dims, starts, steps, vals, lin, x, out[i]);
            case 3: ORACLE_RUN_N(3, cubic_regular_one, n, dims, starts, steps, vals, lin, x, out[i]);
            case 4: ORACLE_RUN_N(4, cubic_regular_one, n, dims, starts, steps, vals, lin, x, out[i]);
            default: break;
        }
    }
    if (method == 0) {
        const bool rec = use_recursive(order, ndims, 6);
        if (fma && rec) ORACLE_RUN((linear_regular_one<true, true>(n, dims, starts, steps, vals, x, out[i])));
        if (fma) ORACLE_RUN((linear_regular_one<true, false>(n, dims, starts, steps, vals, x, out[i])));
        if (rec) ORACLE_RUN((linear_regular_one<false, true>(n, dims, starts, steps, vals, x, out[i])));
        ORACLE_RUN((linear_regular_one<false, false>(n, dims, starts, steps, vals, x, out[i])));
    } else if (cubic) {
        const bool rec = use_recursive(order, ndims, 4);
        if (fma && rec) ORACLE_RUN((cubic_regular_one<true, true>(n, dims, starts, steps, vals, lin, x, out[i])));
        if (fma) ORACLE_RUN((cubic_regular_one<true, false>(n, dims, starts, steps, vals, lin, x, out[i])));
        if (rec) ORACLE_RUN((cubic_regular_one<false, true>(n, dims, starts, steps, vals, lin, x, out[i])));
        ORACLE_RUN((cubic_regular_one<false, false>(n, dims, starts, steps, vals, lin, x, out[i])));
    } else {
        if (fma) ORACLE_RUN((nearest_regular_one<true>(n, dims, starts, steps, vals, x, out[i])));
        ORACLE_RUN((nearest_regular_one<false>(n, dims, starts, steps, vals, x, out[i])));
    }
}

template <class T>
int rectilinear_impl(int method, const T* const* grids, const size_t* grid_lens, size_t ngrids, const T* vals,
                     size_t nvals, int linearize, const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,
                     size_t nout, int fma, int order, int nthreads) {
    const bool cubic = method == 1, nearest = method == 2;
    const size_t ndims = ngrids;
    if (!cubic && nobs != ndims) return DIM_MISMATCH;  // multilinear/rectilinear.rs:58-61, nearest/rectilinear.rs:45-48
    const size_t maxdims = nearest ? 6 : 8;
    if (ndims < 1 || ndims > maxdims) return nearest ? MAXDIM_6 : MAXDIM_8;
    if (nvals != product(grid_lens, ndims)) return DIM_MISMATCH;
    const size_t mindim = cubic ? 4 : 2;
    for (size_t i = 0; i < ndims; ++i)
        if (grid_lens[i] < mindim) return cubic ? MIN_4 : MIN_2;
    for (size_t i = 0; i < ndims; ++i)
        if (!(grids[i][1] > grids[i][0])) return NOT_MONOTONIC;
    if (nobs != ndims) return DIM_MISMATCH;  // cubic N<=4: reference panics
    if (!lens_match(obs_lens, nobs, nout)) return DIM_MISMATCH;

    const int n = static_cast<int>(ndims);
    const bool lin = linearize != 0;
    const size_t* dims = grid_lens;
    size_t* first_bad = nullptr;
    auto gather = [&](size_t i, T* x) {
        for (int j = 0; j < n; ++j) x[j] = obs[j][i];
    };
#define ORACLE_RECT_N(NC, FN, ...)                                                                      \
    return run_batch(nout, nthreads, first_bad, [&](size_t i) {                                         \
        T x[MAXDIMS];                                                                                   \
        for (int j = 0; j < NC; ++j) x[j] = obs[j][i];                                                  \
        if (fma) FN<true, false, NC>(__VA_ARGS__);                                                      \
        else FN<false, false, NC>(__VA_ARGS__);                                                         \
        return true;                                                                                    \
    })
    if (order == 0 && method == 0) {
        switch (n) {
            case 1: ORACLE_RECT_N(1, linear_rect_one, n, dims, grids, vals, x, out[i]);
            case 2: ORACLE_RECT_N(2, linear_rect_one, n, dims, grids, vals, x, out[i]);
            case 3: ORACLE_RECT_N(3, linear_rect_one, n, dims, grids, vals, x, out[i]);
            case 4: ORACLE_RECT_N(4, linear_rect_one, n, dims, grids, vals, x, out[i]);
            case 5: ORACLE_RECT_N(5, linear_rect_one, n, dims, grids, vals, x, out[i]);
            case 6: ORACLE_RECT_N(6, linear_rect_one, n, dims, grids, vals, x, out[i]);
            default: break;
        }
    }
    if (order == 0 && cubic) {
        switch (n) {
            case 1: ORACLE_RECT_N(1, cubic_rect_one, n, dims, grids, vals, lin, x, out[i]);
            case 2: ORACLE_RECT_N(2, cubic_rect_one, n, dims, grids, vals, lin, x, out[i]);
            case 3: ORACLE_RECT_N(3, cubic_rect_one, n, dims, grids, vals, lin, x, out[i]);
            case 4: ORACLE_RECT_N(4, cubic_rect_one, n, dims, grids, vals, lin, x, out[i]);
            default: break;
        }
    }
    if (method == 0) {
        const bool rec = use_recursive(order, ndims, 6);
        if (fma && rec) ORACLE_RUN((linear_rect_one<true, true>(n, dims, grids, vals, x, out[i]), true));
        if (fma) ORACLE_RUN((linear_rect_one<true, false>(n, dims, grids, vals, x, out[i]), true));
        if (rec) ORACLE_RUN((linear_rect_one<false, true>(n, dims, grids, vals, x, out[i]), true));
        ORACLE_RUN((linear_rect_one<false, false>(n, dims, grids, vals, x, out[i]), true));
    } else if (cubic) {
        const bool rec = use_recursive(order, ndims, 4);
        if (fma && rec) ORACLE_RUN((cubic_rect_one<true, true>(n, dims, grids, vals, lin, x, out[i]), true));
        if (fma) ORACLE_RUN((cubic_rect_one<true, false>(n, dims, grids, vals, lin, x, out[i]), true));
        if (rec) ORACLE_RUN((cubic_rect_one<false, true>(n, dims, grids, vals, lin, x, out[i]), true));
        ORACLE_RUN((cubic_rect_one<false, false>(n, dims, grids, vals, lin, x, out[i]), true));
    } else {
        ORACLE_RUN((nearest_rect_one(n, dims, grids, vals, x, out[i]), true));
    }
}
#undef ORACLE_RUN

template <class T>
int one_dim_regular_impl(int kind, T start, T step, const T* vals, size_t nvals, const T* locs, size_t nlocs, T* out,
                         size_t nout, int fma, size_t* first_bad) {
    if (first_bad) *first_bad = SIZE_MAX;
    // RegularGrid1D::new (one_dim/mod.rs:85-95). `vals.len() - 1` / `len - 2` underflow-panic in the
    // reference for fewer than two values; reported as LENGTH_MISMATCH here.
    if (nvals < 2) return LENGTH_MISMATCH;
    if (kind < 0 || kind > 4) return DIM_MISMATCH;
    T stop = start + step * static_cast<T>(nvals - 1);
    if (nlocs != nout) return LENGTH_MISMATCH;  // one_dim/mod.rs:52-54
    for (size_t i = 0; i < nlocs; ++i) {
        GridSample<T> s;
        if (!regular_1d_at(start, stop, step, vals, nvals, locs[i], s)) {
            if (first_bad) *first_bad = i;
            return UNREPRESENTABLE_NUM;
        }
        out[i] = fma ? eval_1d<true>(kind, s, locs[i]) : eval_1d<false>(kind, s, locs[i]);
    }
    return OK;
}

template <class T>
int one_dim_rect_impl(int kind, const T* grid, size_t ngrid, const T* vals, size_t nvals, const T* locs, size_t nlocs,
                      T* out, size_t nout, int fma) {
    if (ngrid != nvals || ngrid < 2) return LENGTH_MISMATCH;  // one_dim/mod.rs:148-152
    if (kind < 0 || kind > 4) return DIM_MISMATCH;
    if (nlocs != nout) return LENGTH_MISMATCH;
    for (size_t i = 0; i < nlocs; ++i) {
        GridSample<T> s;
        rect_1d_at(grid, vals, ngrid, locs[i], s);
        out[i] = fma ? eval_1d<true>(kind, s, locs[i]) : eval_1d<false>(kind, s, locs[i]);
    }
    return OK;
}

// ref: multilinear/regular.rs:145-182
template <class T>
int check_bounds_regular_impl(const size_t* dims, size_t ndims, const T* starts, const T* steps, const T* const* obs,
                              const size_t* obs_lens, size_t nobs, T atol, uint8_t* out, size_t nout) {
    if (!(nobs == ndims && nout == ndims)) return DIM_MISMATCH;
    for (size_t i = 0; i < ndims; ++i) {
        T first = starts[i];
        T last = starts[i] + steps[i] * static_cast<T>(dims[i] - 1);
        T lo = std::fmin(first, last);
        T hi = std::fmax(first, last);
        bool bad = false;
        for (size_t k = 0; k < obs_lens[i] && !bad; ++k) {
            T x = obs[i][k];
            bad = (x - lo) <= -atol || (x - hi) >= atol;
        }
        out[i] = bad ? 1 : 0;
    }
    return OK;
}

// ref: multilinear/rectilinear.rs:109-134
template <class T>
int check_bounds_rect_impl(const T* const* grids, const size_t* grid_lens, size_t ngrids, const T* const* obs,
                           const size_t* obs_lens, size_t nobs, T atol, uint8_t* out, size_t nout) {
    if (!(nobs == ngrids && nout == ngrids)) return DIM_MISMATCH;
    for (size_t i = 0; i < ngrids; ++i)
        if (grid_lens[i] == 0) return DIM_MISMATCH;
    for (size_t i = 0; i < ngrids; ++i) {
        T lo = grids[i][0];
        T hi = grids[i][grid_lens[i] - 1];
        bool bad = false;
        for (size_t k = 0; k < obs_lens[i] && !bad; ++k) {
            T x = obs[i][k];
            bad = (x - lo) <= -atol || (x - hi) >= atol;
        }
        out[i] = bad ? 1 : 0;
    }
    return OK;
}

}  // namespace

extern "C" {

const char* oracle_strerror(int status) { return oracle::strerror(status); }
int oracle_max_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

#define ORACLE_DEFINE(SUFFIX, T)                                                                                      \
    int oracle_regular_##SUFFIX(int method, const size_t* dims, size_t ndims, const T* starts, size_t nstarts,        \
                                const T* steps, size_t nsteps, const T* vals, size_t nvals, int linearize,            \
                                const T* const* obs, const size_t* obs_lens, size_t nobs, T* out, size_t nout,        \
                                int fma, int order, int nthreads, size_t* first_bad) {                                \
        return regular_impl<T>(method, dims, ndims, starts, nstarts, steps, nsteps, vals, nvals, linearize, obs,      \
                               obs_lens, nobs, out, nout, fma, order, nthreads, first_bad);                           \
    }                                                                                                                 \
    int oracle_rectilinear_##SUFFIX(int method, const T* const* grids, const size_t* grid_lens, size_t ngrids,        \
                                    const T* vals, size_t nvals, int linearize, const T* const* obs,                  \
                                    const size_t* obs_lens, size_t nobs, T* out, size_t nout, int fma, int order,     \
                                    int nthreads) {                                                                   \
        return rectilinear_impl<T>(method, grids, grid_lens, ngrids, vals, nvals, linearize, obs, obs_lens, nobs,     \
                                   out, nout, fma, order, nthreads);                                                  \
    }                                                                                                                 \
    int oracle_one_dim_regular_##SUFFIX(int kind, T start, T step, const T* vals, size_t nvals, const T* locs,        \
                                        size_t nlocs, T* out, size_t nout, int fma, size_t* first_bad) {              \
        return one_dim_regular_impl<T>(kind, start, step, vals, nvals, locs, nlocs, out, nout, fma, first_bad);       \
    }                                                                                                                 \
    int oracle_one_dim_rectilinear_##SUFFIX(int kind, const T* grid, size_t ngrid, const T* vals, size_t nvals,       \
                                            const T* locs, size_t nlocs, T* out, size_t nout, int fma) {              \
        return one_dim_rect_impl<T>(kind, grid, ngrid, vals, nvals, locs, nlocs, out, nout, fma);                     \
    }                                                                                                                 \
    int oracle_check_bounds_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, const T* steps,       \
                                             const T* const* obs, const size_t* obs_lens, size_t nobs, T atol,        \
                                             uint8_t* out, size_t nout) {                                             \
        return check_bounds_regular_impl<T>(dims, ndims, starts, steps, obs, obs_lens, nobs, atol, out, nout);        \
    }                                                                                                                 \
    int oracle_check_bounds_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,       \
                                                 const T* const* obs, const size_t* obs_lens, size_t nobs, T atol,    \
                                                 uint8_t* out, size_t nout) {                                         \
        return check_bounds_rect_impl<T>(grids, grid_lens, ngrids, obs, obs_lens, nobs, atol, out, nout);             \
    }

ORACLE_DEFINE(f64, double)
ORACLE_DEFINE(f32, float)

}  // extern "C"
