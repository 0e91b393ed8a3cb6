/* interpn_b200.h — C ABI of the B200-native (sm_100a) InterpN interpolation hot path.
 *
 * Drop-in boundary for jlogan03/interpn 0.8.2: each entry point replaces one reference
 * interface (cited as file:line relative to the reference's src/). The library is built by nvcc
 * for sm_100a only; there is no CPU fallback — every compute entry point fails with
 * INTERPN_B200_ERR_CUDA / INTERPN_B200_ERR_NO_DEVICE when no B200 is usable.
 *
 * Conventions
 *  - Every Rust slice `&[T]` is passed as (pointer, length); `&[&[T]]` as (array of pointers,
 *    array of lengths, count). Lengths are checked exactly where the reference checks them and
 *    yield the same messages (interpn_b200_strerror). Where the reference would panic on a
 *    wrong-length slice (`try_into().unwrap()`, e.g. multicubic/regular.rs:65-73) the library
 *    returns INTERPN_B200_ERR_DIM_MISMATCH instead.
 *  - `vals` is flat C-order (last dimension contiguous); `obs` is one contiguous array per
 *    dimension (SoA); `out` has one entry per query point.
 *  - Arithmetic follows the crate's default feature set (no `fma`): every a*b+c is two rounded
 *    operations, in the reference's operation order, so results are bit-identical to the Rust
 *    crate built with default features.
 *  - Caller owns all buffers. The `*_host` flavour borrows HOST pointers for the call, copies
 *    to/from the device internally and returns after the result is in `out`. On a per-point
 *    failure ("Unrepresentable coordinate value") it reproduces the reference's serial
 *    semantics: out[0 .. first_bad) is written, out[first_bad ..] is left untouched
 *    (multilinear/regular.rs:276-280).
 *  - The interpolator ("interp") API keeps the grid resident in HBM across calls — the analogue
 *    of the reference's structs (`MultilinearRegular::new(..)?.interp(obs, out)`) — and offers
 *    a stream-ordered device-pointer evaluation for callers whose data already lives on the GPU.
 *  - Thread safety: distinct interps may be used from distinct host threads. One interp may be
 *    used from several threads only through interpn_b200_interp_eval_device_* on distinct streams.
 */
#ifndef INTERPN_B200_H
#define INTERPN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes. 1..11 carry the reference's `&'static str` error literals verbatim. */
typedef enum interpn_b200_status {
    INTERPN_B200_OK = 0,
    INTERPN_B200_ERR_DIM_MISMATCH = 1,        /* "Dimension mismatch"  multilinear/regular.rs:61,240,273 */
    INTERPN_B200_ERR_MIN_TWO = 2,             /* "All grids must have at least two entries"  multilinear/regular.rs:245 */
    INTERPN_B200_ERR_MIN_2 = 3,               /* "All grids must have at least 2 entries"  multilinear/rectilinear.rs:192 */
    INTERPN_B200_ERR_MIN_FOUR = 4,            /* "All grids must have at least four entries"  multicubic/regular.rs:261 */
    INTERPN_B200_ERR_MIN_4 = 5,               /* "All grids must have at least 4 entries"  multicubic/rectilinear.rs:214 */
    INTERPN_B200_ERR_NOT_MONOTONIC = 6,       /* "All grids must be monotonically increasing"  multilinear/regular.rs:250 */
    INTERPN_B200_ERR_UNREPRESENTABLE = 7,     /* "Unrepresentable coordinate value"  multilinear/regular.rs:330,418 */
    INTERPN_B200_ERR_MAXDIM_8 = 8,            /* "Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions."  multilinear/regular.rs:111-113 */
    INTERPN_B200_ERR_MAXDIM_6 = 9,            /* "Dimension exceeds maximum (6)."  nearest/regular.rs:97 */
    INTERPN_B200_ERR_LENGTH_MISMATCH = 10,    /* "Length mismatch"  one_dim/mod.rs:53,150 */
    INTERPN_B200_ERR_UNREPRESENTABLE_NUM = 11,/* "Unrepresentable number"  one_dim/mod.rs:88,111 */
    /* Library-level failures with no reference analogue: */
    INTERPN_B200_ERR_CUDA = 100,              /* a CUDA runtime call failed; see interpn_b200_last_error_detail() */
    INTERPN_B200_ERR_NO_DEVICE = 101,         /* no CUDA device / not an sm_100 device */
    INTERPN_B200_ERR_INVALID_ARG = 102,       /* NULL pointer, unknown method/kind, dtype mismatch with the interp */
    INTERPN_B200_ERR_TOO_LARGE = 103          /* grid does not fit device memory / index range */
} interpn_b200_status;

/* Interpolation methods (module names multilinear / multicubic / nearest). */
enum { INTERPN_B200_LINEAR = 0, INTERPN_B200_CUBIC = 1, INTERPN_B200_NEAREST = 2 };

/* one_dim interpolator kinds: Linear1D, LinearHoldLast1D (one_dim/linear.rs:9,43),
 * Left1D, Right1D, Nearest1D (one_dim/hold.rs:8,43,79). */
enum {
    INTERPN_B200_1D_LINEAR = 0,
    INTERPN_B200_1D_LINEAR_HOLD_LAST = 1,
    INTERPN_B200_1D_LEFT = 2,
    INTERPN_B200_1D_RIGHT = 3,
    INTERPN_B200_1D_NEAREST = 4
};

/* Where the `vals` passed to an interp constructor live. */
enum {
    INTERPN_B200_VALS_HOST = 0,    /* host pointer: copied to the device once */
    INTERPN_B200_VALS_DEVICE = 1,  /* device pointer: copied device-to-device once */
    INTERPN_B200_VALS_UNINIT = 2   /* `vals` ignored: device storage is allocated but not filled; the caller
                                      fills interpn_b200_interp_vals_ptr() itself (e.g. the target of the one
                                      NCCL broadcast that replicates the grid to every rank) */
};

/* The reference's literal message for a status (empty string for OK). Never NULL. */
const char* interpn_b200_strerror(int status);
/* Detail of the last INTERPN_B200_ERR_CUDA on this host thread (cudaGetErrorString + call site). */
const char* interpn_b200_last_error_detail(void);

int interpn_b200_device_count(void);
/* Select the CUDA device used by subsequent calls on this host thread (cudaSetDevice). */
int interpn_b200_set_device(int device);
/* How many GPUs ONE host-buffer call may use (the one-shot functions and interpn_b200_interp_eval_host_*): the query
 * batch is cut into chunks that the devices pull from a shared counter, each evaluating on its own replica of the grid
 * (copied device-to-device on first use) — BASELINE north_star: "multi-GPU runs shard the query batch ... with the
 * grid replicated once". n = 0 (the default, or INTERPN_B200_HOST_DEVICES unset/0): every visible sm_100 device;
 * n = 1: only the interpolator's own device (what a one-process-per-GPU launcher such as torchrun wants). Batches
 * below 2^22 points always stay on one device. Results, the failure index and the "earlier outputs written, later
 * untouched" rule are the same for any n. */
int interpn_b200_set_host_devices(int n);
int interpn_b200_host_devices(void);
/* Host threads that copy PAGEABLE caller memory to and from pinned staging buffers (INTERPN_B200_COPY_THREADS; pinned
 * caller memory is DMA'd in place). */
int interpn_b200_copy_threads(void);
/* Number of kernels this library has launched in this process (all threads); bench.py reports the delta. */
uint64_t interpn_b200_launch_count(void);
/* How many of those launches were bin-swept evaluations (grids beyond L2, interpn_b200/csrc/sweep.cuh). */
uint64_t interpn_b200_swept_launch_count(void);
/* Arithmetic flavour this build of the library reproduces: 0 = the crate's default features (no fused
 * multiply-add anywhere, e.g. multilinear/regular.rs:382-385), 1 = the crate's `fma` feature (mul_add at the
 * reference's sites: multilinear/regular.rs:334-337,377-388, multicubic/mod.rs:84-90,111-116,
 * multicubic/regular.rs:528-564, nearest/regular.rs:272-275, one_dim/linear.rs:30-35; the Python wheel's build,
 * pyproject.toml:72). libinterpn_b200.so is flavour 0, libinterpn_b200_fma.so flavour 1; same symbols. */
int interpn_b200_arithmetic(void);
/* Streaming multiprocessors of the current device (148 on B200); 0 when there is no device. */
int interpn_b200_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * One-shot, host-buffer entry points: exactly what the 12 PyO3 functions bind
 * (python.rs:55-85, 119-147, 149-178, 180-201, 228-260, 262-292), i.e. the Rust functions
 *   multilinear::regular::interpn      multilinear/regular.rs:51-58
 *   multilinear::rectilinear::interpn  multilinear/rectilinear.rs:49-54
 *   multicubic::regular::interpn       multicubic/regular.rs:52-60
 *   multicubic::rectilinear::interpn   multicubic/rectilinear.rs:54-60
 *   nearest::regular::interpn          nearest/regular.rs:41-48
 *   nearest::rectilinear::interpn      nearest/rectilinear.rs:39-44
 * `first_bad` (may be NULL) receives the index of the first unrepresentable query point when
 * the status is INTERPN_B200_ERR_UNREPRESENTABLE, else SIZE_MAX.
 * ------------------------------------------------------------------------------------------------ */
#define INTERPN_B200_DECLARE_ONESHOT(SUFFIX, T)                                                                   \
    int interpn_b200_linear_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,   \
                                             const T* steps, size_t nsteps, const T* vals, size_t nvals,          \
                                             const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,    \
                                             size_t nout, size_t* first_bad);                                     \
    int interpn_b200_linear_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,   \
                                                 const T* vals, size_t nvals, const T* const* obs,                \
                                                 const size_t* obs_lens, size_t nobs, T* out, size_t nout);       \
    int interpn_b200_cubic_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,    \
                                            const T* steps, size_t nsteps, const T* vals, size_t nvals,           \
                                            int linearize_extrapolation, const T* const* obs,                     \
                                            const size_t* obs_lens, size_t nobs, T* out, size_t nout,             \
                                            size_t* first_bad);                                                   \
    int interpn_b200_cubic_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,    \
                                                const T* vals, size_t nvals, int linearize_extrapolation,         \
                                                const T* const* obs, const size_t* obs_lens, size_t nobs, T* out, \
                                                size_t nout);                                                     \
    int interpn_b200_nearest_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts, size_t nstarts,  \
                                              const T* steps, size_t nsteps, const T* vals, size_t nvals,         \
                                              const T* const* obs, const size_t* obs_lens, size_t nobs, T* out,   \
                                              size_t nout, size_t* first_bad);                                    \
    int interpn_b200_nearest_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens, size_t ngrids,  \
                                                  const T* vals, size_t nvals, const T* const* obs,               \
                                                  const size_t* obs_lens, size_t nobs, T* out, size_t nout);      \
    /* check_bounds (multilinear/regular.rs:145-182, multilinear/rectilinear.rs:109-134;                          \
     * python.rs:87-117, 203-226). out[i] = 1 if any obs[i][k] violates axis i's bounds by atol. */                \
    int interpn_b200_check_bounds_regular_##SUFFIX(const size_t* dims, size_t ndims, const T* starts,             \
                                                   size_t nstarts, const T* steps, size_t nsteps,                 \
                                                   const T* const* obs, const size_t* obs_lens, size_t nobs,      \
                                                   T atol, uint8_t* out, size_t nout);                            \
    int interpn_b200_check_bounds_rectilinear_##SUFFIX(const T* const* grids, const size_t* grid_lens,            \
                                                       size_t ngrids, const T* const* obs,                        \
                                                       const size_t* obs_lens, size_t nobs, T atol, uint8_t* out, \
                                                       size_t nout);                                              \
    /* one_dim: `Kind::new(RegularGrid1D::new(start, step, vals)?).eval(locs, out)` and the                       \
     * RectilinearGrid1D twin (one_dim/mod.rs:41-61, 85-138, 147-187; linear.rs; hold.rs). */                     \
    int interpn_b200_one_dim_regular_##SUFFIX(int kind, T start, T step, const T* vals, size_t nvals,             \
                                              const T* locs, size_t nlocs, T* out, size_t nout,                   \
                                              size_t* first_bad);                                                 \
    int interpn_b200_one_dim_rectilinear_##SUFFIX(int kind, const T* grid, size_t ngrid, const T* vals,           \
                                                  size_t nvals, const T* locs, size_t nlocs, T* out,              \
                                                  size_t nout);

INTERPN_B200_DECLARE_ONESHOT(f64, double)
INTERPN_B200_DECLARE_ONESHOT(f32, float)

/* ------------------------------------------------------------------------------------------------
 * Resident interpolators: the struct API `X::new(..)?` / `.interp(obs, out)`
 * (e.g. multilinear/regular.rs:225-283, multicubic/rectilinear.rs:193-253) with the grid kept in
 * HBM. `method` selects linear / cubic / nearest. Constructors perform the reference's `new()`
 * validation and return its errors.
 * ------------------------------------------------------------------------------------------------ */
typedef struct interpn_b200_interp interpn_b200_interp;

#define INTERPN_B200_DECLARE_INTERP(SUFFIX, T)                                                                    \
    int interpn_b200_regular_new_##SUFFIX(int method, const size_t* dims, size_t ndims, const T* starts,          \
                                          size_t nstarts, const T* steps, size_t nsteps, const T* vals,           \
                                          size_t nvals, int linearize_extrapolation, int vals_location,           \
                                          interpn_b200_interp** out_interp);                                      \
    int interpn_b200_rectilinear_new_##SUFFIX(int method, const T* const* grids, const size_t* grid_lens,         \
                                              size_t ngrids, const T* vals, size_t nvals,                         \
                                              int linearize_extrapolation, int vals_location,                     \
                                              interpn_b200_interp** out_interp);                                  \
    /* `.interp(obs, out)` on HOST buffers (copies in/out, synchronous, reference error semantics); pageable or   \
     * pinned memory, one or several GPUs (interpn_b200_set_host_devices). */                                     \
    int interpn_b200_interp_eval_host_##SUFFIX(interpn_b200_interp* interp, const T* const* obs,                  \
                                               const size_t* obs_lens, size_t nobs, T* out, size_t nout,          \
                                               size_t* first_bad);                                                \
    /* `.interp(obs, out)` on DEVICE buffers: `obs` is a host array of `nobs` device pointers, each with `n`      \
     * entries; `out` is a device pointer with `n` entries. Enqueued on `stream` (a cudaStream_t; NULL = legacy   \
     * default stream) and returns without synchronising. Unrepresentable points are not written; their smallest  \
     * index is latched in the interp and read by interpn_b200_interp_status(). */                                \
    int interpn_b200_interp_eval_device_##SUFFIX(interpn_b200_interp* interp, const T* const* obs, size_t nobs,   \
                                                 size_t n, T* out, void* stream);                                 \
    /* Several fields over ONE grid and ONE query batch (the pattern of the reference's benchmark, six            \
     * interpolators on one grid, bench_cpu.py:501-510): `interps` are `nfields` interpolators built over the      \
     * same grid with the same method (else "Dimension mismatch"), `outs` a host array of `nfields` device         \
     * pointers. Multilinear and nearest fields on grids within L2 share one cell location per point; every       \
     * field's result is bit-identical to its own interpn_b200_interp_eval_device_* call. Failures are latched   \
     * on interps[0]. */                                                                                          \
    int interpn_b200_interp_eval_fields_device_##SUFFIX(interpn_b200_interp* const* interps, size_t nfields,      \
                                                        const T* const* obs, size_t nobs, size_t n,               \
                                                        T* const* outs, void* stream);

INTERPN_B200_DECLARE_INTERP(f64, double)
INTERPN_B200_DECLARE_INTERP(f32, float)

/* Synchronise `stream`, then report and clear the latched per-point failure of all device evaluations
 * enqueued so far: INTERPN_B200_OK, or INTERPN_B200_ERR_UNREPRESENTABLE with *first_bad = smallest index. */
int interpn_b200_interp_status(interpn_b200_interp* interp, void* stream, size_t* first_bad);
/* Device pointer / element count / element size of the resident `vals` copy (for INTERPN_B200_VALS_UNINIT). */
void* interpn_b200_interp_vals_ptr(interpn_b200_interp* interp);
/* Must be called (stream-ordered on `stream`) after the caller has written the resident `vals` through
 * interpn_b200_interp_vals_ptr(): refreshes the gather-optimised copies the library derives from them. */
int interpn_b200_interp_vals_updated(interpn_b200_interp* interp, void* stream);
size_t interpn_b200_interp_vals_len(const interpn_b200_interp* interp);
size_t interpn_b200_interp_elem_size(const interpn_b200_interp* interp);
size_t interpn_b200_interp_ndims(const interpn_b200_interp* interp);
void interpn_b200_interp_free(interpn_b200_interp* interp);

#ifdef __cplusplus
}
#endif
#endif /* INTERPN_B200_H */
