/* A plain C consumer of include/interpn_b200.h — what a host crate's FFI layer does, without Python in between.
 *
 * Known answers, all from the reference's own tests (restated): a field that is linear in every coordinate is reproduced
 * exactly by multilinear interpolation inside the grid (multilinear/regular.rs tests) and by multicubic interpolation
 * with linearized extrapolation inside and outside (multicubic/regular.rs tests, 1e-12 there); nearest returns a grid
 * value; argument errors come back as the reference's literal messages; an unrepresentable coordinate stops the batch at
 * its index with the earlier outputs written.
 *
 * Exit code 0 = every check passed; 77 = no sm_100 device (the library refuses to compute: there is no CPU path);
 * anything else = failure. Built and run by tests/test_cabi_c_program.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "interpn_b200.h"

#define CHECK(cond, ...)                          \
    do {                                          \
        if (!(cond)) {                            \
            fprintf(stderr, "FAILED: " __VA_ARGS__); \
            fprintf(stderr, "\n");                \
            return 1;                             \
        }                                         \
    } while (0)

static double field(double x, double y, double z) { return 1.5 * x - 0.25 * y + 2.0 * z + 3.0; }

int main(void) {
    enum { NX = 7, NY = 5, NZ = 6, N = 1000 };
    const size_t dims[3] = {NX, NY, NZ};
    const double starts[3] = {-1.0, 0.0, 2.0}, steps[3] = {0.5, 1.0, 0.25};
    static double vals[NX * NY * NZ];
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NY; ++j)
            for (int k = 0; k < NZ; ++k)
                vals[(i * NY + j) * NZ + k] = field(starts[0] + steps[0] * i, starts[1] + steps[1] * j, starts[2] + steps[2] * k);
    static double x[N], y[N], z[N], out[N];
    unsigned s = 12345u;
    for (int p = 0; p < N; ++p) { /* inside the grid */
        s = s * 1664525u + 1013904223u; x[p] = starts[0] + steps[0] * (NX - 1) * (s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; y[p] = starts[1] + steps[1] * (NY - 1) * (s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; z[p] = starts[2] + steps[2] * (NZ - 1) * (s >> 8) / 16777216.0;
    }
    const double* obs[3] = {x, y, z};
    const size_t lens[3] = {N, N, N};
    size_t first_bad = 0;

    /* argument errors are reported without a device, with the reference's messages */
    int st = interpn_b200_linear_regular_f64(dims, 3, starts, 2, steps, 3, vals, NX * NY * NZ, obs, lens, 3, out, N, &first_bad);
    CHECK(st != INTERPN_B200_OK && strcmp(interpn_b200_strerror(st), "Dimension mismatch") == 0, "starts of the wrong length: %s", interpn_b200_strerror(st));
    const size_t small[3] = {3, NY, NZ};
    st = interpn_b200_cubic_regular_f64(small, 3, starts, 3, steps, 3, vals, 3 * NY * NZ, 1, obs, lens, 3, out, N, &first_bad);
    CHECK(st == INTERPN_B200_ERR_MIN_FOUR && strcmp(interpn_b200_strerror(st), "All grids must have at least four entries") == 0,
          "cubic on a 3-node axis: %s", interpn_b200_strerror(st));

    st = interpn_b200_linear_regular_f64(dims, 3, starts, 3, steps, 3, vals, NX * NY * NZ, obs, lens, 3, out, N, &first_bad);
    if (st == INTERPN_B200_ERR_NO_DEVICE) {
        printf("no sm_100 device: %s\n", interpn_b200_strerror(st));
        return 77;
    }
    CHECK(st == INTERPN_B200_OK, "linear: %s (%s)", interpn_b200_strerror(st), interpn_b200_last_error_detail());
    for (int p = 0; p < N; ++p) CHECK(fabs(out[p] - field(x[p], y[p], z[p])) < 1e-12, "linear point %d: %.17g", p, out[p]);

    /* multicubic with linearized extrapolation reproduces a linear field outside the grid as well */
    for (int p = 0; p < N; p += 3) x[p] += (p % 2 ? 5.0 : -5.0);
    st = interpn_b200_cubic_regular_f64(dims, 3, starts, 3, steps, 3, vals, NX * NY * NZ, 1, obs, lens, 3, out, N, &first_bad);
    CHECK(st == INTERPN_B200_OK, "cubic: %s", interpn_b200_strerror(st));
    for (int p = 0; p < N; ++p) CHECK(fabs(out[p] - field(x[p], y[p], z[p])) < 1e-11, "cubic point %d: %.17g vs %.17g", p, out[p], field(x[p], y[p], z[p]));

    /* the struct API: new once, evaluate twice, free; rectilinear axes that happen to be regular give the same field */
    static double gx[NX], gy[NY], gz[NZ];
    for (int i = 0; i < NX; ++i) gx[i] = starts[0] + steps[0] * i;
    for (int j = 0; j < NY; ++j) gy[j] = starts[1] + steps[1] * j;
    for (int k = 0; k < NZ; ++k) gz[k] = starts[2] + steps[2] * k;
    const double* grids[3] = {gx, gy, gz};
    const size_t glens[3] = {NX, NY, NZ};
    interpn_b200_interp* it = NULL;
    st = interpn_b200_rectilinear_new_f64(INTERPN_B200_CUBIC, grids, glens, 3, vals, NX * NY * NZ, 1, INTERPN_B200_VALS_HOST, &it);
    CHECK(st == INTERPN_B200_OK && it != NULL, "rectilinear_new: %s", interpn_b200_strerror(st));
    CHECK(interpn_b200_interp_ndims(it) == 3 && interpn_b200_interp_vals_len(it) == (size_t)(NX * NY * NZ), "resident metadata");
    for (int rep = 0; rep < 2; ++rep) {
        memset(out, 0, sizeof out);
        st = interpn_b200_interp_eval_host_f64(it, obs, lens, 3, out, N, &first_bad);
        CHECK(st == INTERPN_B200_OK, "eval_host: %s", interpn_b200_strerror(st));
        for (int p = 0; p < N; ++p) CHECK(fabs(out[p] - field(x[p], y[p], z[p])) < 1e-11, "rectilinear cubic point %d", p);
    }
    interpn_b200_interp_free(it);

    /* nearest returns grid values; an unrepresentable coordinate stops the batch at its index */
    st = interpn_b200_nearest_regular_f64(dims, 3, starts, 3, steps, 3, vals, NX * NY * NZ, obs, lens, 3, out, N, &first_bad);
    CHECK(st == INTERPN_B200_OK, "nearest: %s", interpn_b200_strerror(st));
    for (int p = 0; p < N; ++p) {
        int hit = 0;
        for (int q = 0; q < NX * NY * NZ && !hit; ++q) hit = out[p] == vals[q];
        CHECK(hit, "nearest point %d is not a grid value", p);
    }
    for (int p = 0; p < N; ++p) out[p] = -7.0;
    x[400] = NAN;
    st = interpn_b200_linear_regular_f64(dims, 3, starts, 3, steps, 3, vals, NX * NY * NZ, obs, lens, 3, out, N, &first_bad);
    CHECK(st == INTERPN_B200_ERR_UNREPRESENTABLE && first_bad == 400, "NaN coordinate: status %d first_bad %zu", st, first_bad);
    CHECK(strcmp(interpn_b200_strerror(st), "Unrepresentable coordinate value") == 0, "message: %s", interpn_b200_strerror(st));
    CHECK(out[399] != -7.0 && out[400] == -7.0 && out[N - 1] == -7.0, "outputs before the failing point written, later ones untouched");
    printf("cabi_demo: all checks passed (arithmetic flavour %d, %d SMs, %llu kernel launches)\n", interpn_b200_arithmetic(),
           interpn_b200_sm_count(), (unsigned long long)interpn_b200_launch_count());
    return 0;
}
