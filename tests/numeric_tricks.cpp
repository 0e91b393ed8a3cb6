// numeric_tricks.cpp — CPU re-check of the division-avoiding sequences of interpn_b200/csrc/device_math.cuh
// (markstein_div / exact_div, fast_cell, nearest_upper) and of the exact fusions, permuted end-cell formulas and
// merged operand guard of cubic_quad4.cuh against the IEEE operations the reference performs. Test infrastructure: compiled and run by
// tests/test_numeric_tricks.py with g++ -O2 -ffp-contract=off (std::fma is a single rounding).
//
// Usage: numeric_tricks <millions of random trials>; prints "OK <trials>" or the first counterexample.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static uint64_t s_state = 0x243F6A8885A308D3ull;
static inline uint64_t rnd() {
    uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t bits(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }
static inline double from_bits(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }
static inline double nudge(double x, int k) { return from_bits(bits(x) + (int64_t)k); }  // k ulps (same sign/exponent region)

// ---- device_math.cuh restated with host operations -------------------------------------------
static inline bool operand_ok(double a) {
    const uint64_t b = bits(a);
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned e = (hi >> 20) & 0x7ffu;
    return (e - 723u <= 600u) || ((hi << 1 | lo) == 0u);
}
static inline double markstein_div(double a, double b, double rb) {
    const double q0 = a * rb;
    const double e0 = std::fma(-q0, b, a);
    const double q1 = std::fma(e0, rb, q0);
    const double e1 = std::fma(-q1, b, a);
    return std::copysign(std::fma(e1, rb, q1), a);
}
static inline int floor_sat(double q) {  // F2I.S32.F64.FLOOR: saturating, NaN -> 0
    if (q != q) return 0;
    double f = std::floor(q);
    if (f >= 2147483647.0) return 2147483647;
    if (f <= -2147483648.0) return (int)-2147483648ll;
    return (int)f;
}
static inline bool fast_cell(double x, double start, double step, double rstep, double lim, int dim, int& origin,
                             double& od, double& d) {
    d = x - start;
    const double q = d * rstep;
    const int f = floor_sat(q);
    int o = f < 0 ? 0 : f;
    origin = o > dim - 2 ? dim - 2 : o;
    od = (double)origin;
    const double r = std::fma(-od, step, d);
    const bool proven = r >= 0.0 && r <= lim;
    const bool sane = (unsigned)f + (1u << 30) <= (1u << 31);
    return origin == f ? proven : sane;
}
static inline bool nearest_upper(double e, double hstep, double tau) { return !((e - hstep) <= tau); }

// ---- the reference's operations (multilinear/regular.rs:414-425, nearest/regular.rs:259-293) ---
static inline bool ref_cell(double x, double start, double step, int dim, int& origin) {
    const double q = std::floor((x - start) / step);
    if (!(q >= -9223372036854775808.0 && q < 9223372036854775808.0)) return false;
    long long i = (long long)q;
    long long o = i < 0 ? 0 : i;
    origin = (int)(o > dim - 2 ? dim - 2 : o);
    return true;
}

#define FAIL(...) do { printf(__VA_ARGS__); printf("\n"); return 1; } while (0)

int main(int argc, char** argv) {
    const long trials = (argc > 1 ? atol(argv[1]) : 20) * 1000000L;
    long checked = 0;
    for (long it = 0; it < trials; ++it) {
        // a grid: step over ~60 binades, start over a wide range, a few hundred nodes
        const double step = std::ldexp(0.5 + u01(), (int)(rnd() % 61) - 30);
        const double rstep = 1.0 / step;
        const double start = (u01() - 0.5) * std::ldexp(1.0, (int)(rnd() % 40) - 10);
        const int dim = 2 + (int)(rnd() % 300);
        const double hstep = step * 0.5, tau = step * 0x1p-54, lim = step * (1.0 - 0x1p-20);
        // a query: uniform over the grid +-30 %, or planted on / next to a node, or next to a mid-point
        const int k = (int)(rnd() % (unsigned)dim);
        double x;
        switch (rnd() % 6) {
            case 0: x = start + step * (double)k; break;                                   // the node as the kernels compute it
            case 1: x = nudge(start + step * (double)k, (int)(rnd() % 9) - 4); break;      // within 4 ulps of a node
            case 2: x = nudge(start + step * (double)k + hstep, (int)(rnd() % 9) - 4); break;  // around the nearest tie
            case 3: x = nudge((start + step * (double)k) + hstep, (int)(rnd() % 5) - 2); break;
            default: x = start + step * (double)(dim - 1) * (1.6 * u01() - 0.3); break;
        }
        int o_ref, o_fast;
        double od, d;
        const bool ref_ok = ref_cell(x, start, step, dim, o_ref);
        const bool sure = fast_cell(x, start, step, rstep, lim, dim, o_fast, od, d);
        if (sure) {
            if (!ref_ok || o_ref != o_fast)
                FAIL("fast_cell: x=%a start=%a step=%a dim=%d ref=%d fast=%d", x, start, step, dim, o_ref, o_fast);
            const double x0 = start + step * od;
            const double e = x - x0;
            const bool up_ref = !((e / step) <= 0.5);
            if (nearest_upper(e, hstep, tau) != up_ref) FAIL("nearest_upper: e=%a step=%a", e, step);
            if (operand_ok(e)) {
                const double t_ref = e / step, t = markstein_div(e, step, rstep);
                if (bits(t) != bits(t_ref)) FAIL("markstein t: e=%a step=%a got=%a want=%a", e, step, t, t_ref);
            }
            ++checked;
        }
        // general operands for the division itself (cubic rectilinear spacing ratios reuse it)
        const double b = std::ldexp(0.5 + u01(), (int)(rnd() % 400) - 200);
        double a = std::ldexp(u01() - 0.5, (int)(rnd() % 400) - 200);
        if (rnd() % 8 == 0) a = nudge(b * (double)(rnd() % 1000), (int)(rnd() % 5) - 2);  // near-exact quotients
        if (rnd() % 64 == 0) a = (rnd() & 1) ? 0.0 : -0.0;
        if (operand_ok(a)) {
            const double w = a / b, g = markstein_div(a, b, 1.0 / b);
            if (bits(w) != bits(g)) FAIL("markstein_div: a=%a b=%a got=%a want=%a", a, b, g, w);
        }
    }
    // ---- cubic_quad4.cuh: unclamped cell proof, exact fusions, permuted end-cell formulas, merged operand guard ----
    for (long it = 0; it < trials / 4; ++it) {
        const double step = std::ldexp(0.5 + u01(), (int)(rnd() % 61) - 30);
        const double rstep = 1.0 / step, lim = step * (1.0 - 0x1p-20);
        const double start = (u01() - 0.5) * std::ldexp(1.0, (int)(rnd() % 40) - 10);
        const int dim = 4 + (int)(rnd() % 300);
        const int k = (int)(rnd() % (unsigned)dim);
        double x;
        switch (rnd() % 4) {
            case 0: x = start + step * (double)k; break;
            case 1: x = nudge(start + step * (double)k, (int)(rnd() % 9) - 4); break;
            default: x = start + step * (double)(dim - 1) * (2.0 * u01() - 0.5); break;
        }
        // quad4_locate: f~ = floor(d*rstep) proven by 0 <= fma(-f~, step, d) <= lim, |f~| <= 2^30
        const double d = x - start;
        const int f = floor_sat(d * rstep);
        const double r = std::fma(-(double)f, step, d);
        if (r >= 0.0 && r <= lim && (unsigned)f + (1u << 30) <= (1u << 31)) {
            const double q = std::floor(d / step);
            if (q != (double)f) FAIL("cubic cell: x=%a start=%a step=%a f=%d ref=%a", x, start, step, f, q);
        }
        // grid values over a wide (normal) range, sometimes equal neighbours
        const int ex = (int)(rnd() % 600) - 300;
        double v[4];
        for (double& w : v) w = std::ldexp(u01() - 0.5, ex + (int)(rnd() % 8));
        if (rnd() % 16 == 0) v[1] = v[0];
        if (rnd() % 16 == 0) v[2] = v[1];
        const double dy = v[2] - v[1], d20 = v[2] - v[0], d31 = v[3] - v[1];
        const double k0 = d20 * 0.5, k1 = d31 * 0.5;
        const double a_ref = k0 - dy, b_ref = -k1 + dy;
        const double a_f = std::fma(0.5, d20, -dy), b_f = std::fma(-0.5, d31, dy);
        if (bits(a_ref) != bits(a_f) || bits(b_ref) != bits(b_f)) FAIL("fused a/b: v=%a %a %a %a", v[0], v[1], v[2], v[3]);
        if (bits(b_ref - (a_ref + a_ref)) != bits(std::fma(-2.0, a_ref, b_ref))) FAIL("fused c2: a=%a b=%a", a_ref, b_ref);
        const double two = 2.0;
        if (bits(two * dy - k0) != bits(std::fma(2.0, dy, -k0))) FAIL("fused k1: dy=%a k0=%a", dy, k0);
        // regular low end (multicubic/regular.rs:519-530) on permuted inputs (v2,v1,v0): k0 = -(v2-v0)/2, dy = v0-v1
        // (a zero comes out as +0 where the reference has -0: neg_zero_if)
        auto neg_zero = [](double w) { return w == 0.0 ? -0.0 : w; };
        if (bits(-(v[2] - v[0]) / two) != bits(neg_zero(v[0] - v[2]) * 0.5)) FAIL("low-end k0");
        // rectilinear low end (multicubic/rectilinear.rs:463-477): -(a*((v2-v1)/q) + c*(v1-v0)) with the weights
        // exchanged on the permuted inputs u = (v2,v1,v0): c*(u2-u1) + a*((u1-u0)/q)
        const double qq = 0.25 + 4.0 * u01(), wa = 1.0 / (1.0 + qq), wc = qq / (qq + 1.0);
        const double low_ref = -(wa * ((v[2] - v[1]) / qq) + wc * (v[1] - v[0]));
        const double u0 = v[2], u1 = v[1], u2 = v[0];
        const double low_perm = neg_zero(wc * (u2 - u1) + wa * ((u1 - u0) / qq));
        if (bits(low_ref) != bits(low_perm)) FAIL("rect low end: v=%a %a %a q=%a", v[0], v[1], v[2], qq);
        // merged guard (exact_div_operand_ok on the high word) + bare sequence == IEEE division
        const double num = (rnd() % 8 == 0) ? nudge(qq * (double)(rnd() % 1000), (int)(rnd() % 5) - 2) : v[1] - v[0];
        const unsigned hi = (unsigned)(bits(num) >> 32);
        if ((hi & 0x7fffffffu) - (723u << 20) < (601u << 20)) {
            const double rq = 1.0 / qq;
            const double q0 = num * rq, e0 = std::fma(-q0, qq, num), q1 = std::fma(e0, rq, q0), e1 = std::fma(-q1, qq, num);
            if (bits(std::fma(e1, rq, q1)) != bits(num / qq)) FAIL("bare sequence: a=%a b=%a", num, qq);
        }
    }
    // ---- f32 twins (device_math.cuh): fast_cell / nearest_upper / markstein_div in float against float IEEE ops ----
    {
        auto fbits = [](float x) { uint32_t b; memcpy(&b, &x, 4); return b; };
        auto ffrom = [](uint32_t b) { float x; memcpy(&x, &b, 4); return x; };
        auto fnudge = [&](float x, int k) { return ffrom(fbits(x) + (int32_t)k); };
        long checked32 = 0;
        for (long it = 0; it < trials / 3; ++it) {
            const float step = (float)std::ldexp(0.5 + u01(), (int)(rnd() % 41) - 20);
            const float rstep = 1.0f / step;
            const float start = (float)((u01() - 0.5) * std::ldexp(1.0, (int)(rnd() % 24) - 8));
            const int dim = 2 + (int)(rnd() % 4095);  // <= 4096: the host's precondition
            const float hstep = step * 0.5f, tau = (float)((double)step * 0x1p-25), lim = (float)((double)step * (1.0 - 0x1p-11));
            const int k = (int)(rnd() % (unsigned)dim);
            float x;
            switch (rnd() % 6) {
                case 0: x = start + step * (float)k; break;
                case 1: x = fnudge(start + step * (float)k, (int)(rnd() % 9) - 4); break;
                case 2: x = fnudge(start + step * (float)k + hstep, (int)(rnd() % 9) - 4); break;
                case 3: x = fnudge((start + step * (float)k) + hstep, (int)(rnd() % 5) - 2); break;
                default: x = start + step * (float)(dim - 1) * (float)(1.6 * u01() - 0.3); break;
            }
            // reference (multilinear/regular.rs:414-425 in f32)
            const float qref = std::floor((x - start) / step);
            long long iref = (long long)qref;
            int o_ref = (int)(iref < 0 ? 0 : (iref > dim - 2 ? dim - 2 : iref));
            // fast_cell (float)
            const float d = x - start;
            const float q = d * rstep;
            int f = (q != q) ? 0 : (int)std::floor(q);
            int origin = f < 0 ? 0 : (f > dim - 2 ? dim - 2 : f);
            const float od = (float)origin;
            const float r = std::fmaf(-od, step, d);
            const bool proven = r >= 0.0f && r <= lim;
            const bool sane = (unsigned)f + (1u << 22) <= (1u << 23);
            if (origin == f ? proven : sane) {
                if (o_ref != origin) FAIL("fast_cell f32: x=%a start=%a step=%a dim=%d ref=%d fast=%d", x, start, step, dim, o_ref, origin);
                const float x0 = start + step * od;
                const float e = x - x0;
                const bool up_ref = !((e / step) <= 0.5f);
                if ((!((e - hstep) <= tau)) != up_ref) FAIL("nearest_upper f32: e=%a step=%a", e, step);
                const uint32_t eb = fbits(e);
                if ((((eb >> 23) & 0xffu) - 67u <= 119u) || ((eb << 1) == 0u)) {
                    const float q0 = e * rstep, e0 = std::fmaf(-q0, step, e), q1 = std::fmaf(e0, rstep, q0), e1 = std::fmaf(-q1, step, e);
                    const float t = std::copysign(std::fmaf(e1, rstep, q1), e), t_ref = e / step;
                    if (fbits(t) != fbits(t_ref)) FAIL("markstein f32: e=%a step=%a got=%a want=%a", e, step, t, t_ref);
                }
                ++checked32;
            }
        }
        if (checked32 < trials / 4) FAIL("f32 fast path proved only %ld of %ld", checked32, trials / 3);
        const float fsteps[] = {1.0f, 0.1f, 100.0f / 99.0f, 0.125f, 3.0f, 1e-3f, 12345.678f};
        for (float step : fsteps)
            for (int k = -64; k <= 64; ++k) {
                const float e = fnudge(step * 0.5f, k);
                if ((!((e - step * 0.5f) <= (float)((double)step * 0x1p-25))) != !((e / step) <= 0.5f)) FAIL("f32 tie sweep: e=%a step=%a", e, step);
            }
    }
    // exhaustive neighbourhood of the tie for a few steps: e = step/2 + k ulps, k = -64..64
    const double steps[] = {1.0, 0.1, 100.0 / 99.0, 0.125, 3.0, 1e-7, 12345.678, 0x1.fffffffffffffp+3, 0x1.0000000000001p-5};
    for (double step : steps)
        for (int k = -64; k <= 64; ++k) {
            const double e = nudge(step * 0.5, k);
            if (nearest_upper(e, step * 0.5, step * 0x1p-54) != !((e / step) <= 0.5)) FAIL("tie sweep: e=%a step=%a", e, step);
        }
    // special values never pass the fast path
    const double specials[] = {NAN, INFINITY, -INFINITY, 1e300, -1e300, 0x1p62, -0x1p62};
    for (double x : specials) {
        int o; double od, d;
        if (fast_cell(x, 0.0, 1e-3, 1e3, 1e-3 * (1.0 - 0x1p-20), 10, o, od, d)) FAIL("special value %a passed fast_cell", x);
    }
    printf("OK %ld %ld\n", trials, checked);
    return 0;
}
