// interp_internal.h — host-side description of a resident grid and the launcher interface
// between the C ABI (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// 1 in the build that reproduces the crate's `fma` feature (device_math.cuh kArithFma), else 0.
#ifndef IB200_ARITH_FMA
#define IB200_ARITH_FMA 0
#endif

namespace ib200 {

constexpr int kMaxNd = 8;

// Elements per row of the rectilinear cubic cell table (capi.cu cubic_cell_table, cubic_quad4.cuh): 12 used; f64 rows
// are padded to 14 (112 bytes) so that the 16-byte chunks of random rows spread over all shared-memory bank groups.
constexpr int cubic_cell_row_stride(int elem_bytes) { return elem_bytes == 8 ? 14 : 12; }

// Grid resident in HBM. Host struct; pointers are device pointers.
struct DeviceGrid {
    int method = 0;     // INTERPN_B200_LINEAR / CUBIC / NEAREST
    int rect = 0;       // 0 regular, 1 rectilinear
    int ndims = 0;
    int linearize = 0;
    int elem = 8;       // sizeof(T)
    int dim[kMaxNd] = {};
    long long stride[kMaxNd] = {};  // C-order element strides (ref: multilinear/regular.rs:315-326)
    double start[kMaxNd] = {};      // regular grids; exact copies of the caller's T values
    double step[kMaxNd] = {};
    void* vals = nullptr;           // device, nvals elements
    size_t nvals = 0;
    // Derived copy of `vals` the kernels gather from (policy: capi.cu window_width, builders: launch_misc.cu,
    // launch_cubic_build.cu). win_width = elements per flat index: 2 row pairs / 4 rows of four (plain), 4 = 2x2 patches
    // (linear, win_cross) or first-level coefficient sectors (cubic N = 2..4, win_cross), 8 / 16 = hypercube blocks (linear
    // beyond L2, N = 3 / N = 4..6). nullptr when the grid has none.
    void* win = nullptr;
    int win_width = 0;
    int win_cross = 0;         // not plain rows: linear N >= 2 the 2x2 patch layout, cubic N = 2..4 the coefficient layout
    size_t win_bytes = 0;      // bytes of `win` (the coefficient layout holds (dim0 + 1) slots per node of the other dimensions)
    void* axes = nullptr;           // device, all rectilinear axes packed back to back
    int axis_off[kMaxNd] = {};      // element offset of axis d inside `axes`
    int axes_core = 0;              // elements of the blob before the cell tables (kernels that do not read them stage only this)
    int axes_total = 0;             // elements of the whole blob: axes, then reciprocal cell widths, then bucket tables
    int rect_fast = 0;              // axes strictly increasing and finite: bucket-table search is valid
    int rect_fast_div = 0;          // ... and (f64) every cell width within exact_div's range
    int rc_off[kMaxNd] = {};        // element offset of axis d's reciprocal cell widths (dim-1 entries)
    int lut_off[kMaxNd] = {};       // element offset of axis d's bucket table (lut_nb+1 ints)
    int lut_nb[kMaxNd] = {};
    double lut_scale[kMaxNd] = {};  // buckets per unit length
    int rect_cell = 0;              // linear / nearest, strictly increasing finite axes: cell tables present (search replaced)
    int clut_off[kMaxNd] = {};      // element offset of axis d's cell table (clut_nb ints)
    int clut_nb[kMaxNd] = {};
    double clut_scale[kMaxNd] = {}; // buckets per unit length
    int rect_cubic_table = 0;       // cubic, strictly increasing finite axes: per-cell constant table present
    int ct_off[kMaxNd] = {};        // element offset of axis d's cubic cell table (dim+1 rows of cubic_cell_row_stride elements, 16-byte aligned)
    int sm_count = 148;
    unsigned long long grid_hash = 0;  // FNV-1a of dims and starts/steps or axes: two interpolators over the same grid agree
};

// Enqueue one evaluation of `n` query points. `obs` is a HOST array of ndims device pointers.
// `first_bad` is a device counter (atomicMin of failing indices, offset by `index_base`).
template <class T>
cudaError_t launch_eval(const DeviceGrid& g, const T* const* obs, size_t n, T* out,
                        unsigned long long* first_bad, unsigned long long index_base, cudaStream_t stream);

// Fused evaluation of up to kMaxFields fields over one grid and one query batch (launch_fields.cu). Returns
// cudaErrorNotSupported when there is no fused kernel for the combination.
constexpr int kMaxFields = 8;
template <class T>
cudaError_t launch_eval_fields(const DeviceGrid* const* grids, int nf, const T* const* obs, size_t n, T* const* outs,
                               unsigned long long* first_bad, cudaStream_t stream);

// one_dim kernels (device pointers). kind: INTERPN_B200_1D_*; rect: grid != nullptr.
template <class T>
cudaError_t launch_one_dim(int kind, bool rect, T start, T step, const T* grid, const T* vals, size_t nvals,
                           const T* locs, size_t n, T* out, unsigned long long* first_bad,
                           unsigned long long index_base, cudaStream_t stream);

// check_bounds kernel for one axis: ORs a violation flag into *flag (device int).
template <class T>
cudaError_t launch_check_bounds(const T* x, size_t n, T lo, T hi, T atol, int* flag, cudaStream_t stream);

// Fills g.win from g.vals (stream-ordered).
cudaError_t launch_build_window(const DeviceGrid& g, cudaStream_t stream);
// The cubic coefficient layout (cubic_quad4.cuh; launch_cubic_build.cu), called by launch_build_window.
cudaError_t launch_build_coef_window(const DeviceGrid& g, cudaStream_t stream);
// Elements of the window copy of `g` for window width `w` (0 when there is none).
inline size_t window_elems(const DeviceGrid& g, int w) {
    if (!w) return 0;
    if (g.method == 1 && g.ndims >= 2 && w == 4)  // INTERPN_B200_CUBIC: coefficient layout
        return static_cast<size_t>(g.dim[0] + 1) * static_cast<size_t>(g.stride[0]) * 4;
    return g.nvals * static_cast<size_t>(w);
}

// True when the kernels must index `vals` with 64-bit arithmetic (ref: lib.rs:119-144 indexes with usize): grids of
// 2^31 values or more, or any grid while INTERPN_B200_INDEX64=1 (test hook: runs the `long long` instantiations on
// small grids; read at every call).
bool force_index64();
inline bool index64(const DeviceGrid& g) { return g.nvals >= (size_t(1) << 31) || force_index64(); }

void count_launch();
void count_swept_launch();

}  // namespace ib200
