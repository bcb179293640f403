/*
 * groundwork_ref.c — CPU checker for libb9_groundwork.so.  TEST INFRASTRUCTURE.
 *
 * PARITY UNPINNED — and deliberately so.  This is NOT an oracle for BASE-9.
 * /root/reference holds only README.md:1-4 (a relocation notice); base-cpp is
 * not staged, and BASELINE.json's north_star forbids reconstructing it from
 * memory.  There is therefore no restatement of the reference's likelihood in
 * this directory, and no function below follows any reference file:line.
 *
 * What is here: closed-form host arithmetic that mirrors, operation for
 * operation, the reference-independent groundwork kernels declared in
 * include/b9_groundwork.h, so those kernels can be checked bit for bit where
 * only IEEE add/mul/div/fma are involved (the DFMA chain, the synthetic term
 * generator, the warp-order shard partials, the virtual-shard totals), and to
 * a stated ULP bound where libm is (exp/log/exp10/log10, the row log-sum-exp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this file's library; the product path never does.  oracle/b9_dump.h is the
 * writer of the golden-vector format (tests/golden_io.py) for unblock day.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared groundwork_ref.c -lm
 * (-ffp-contract=off: the only fused operations are the explicit fma() calls.)
 */
#define _GNU_SOURCE            /* exp10() */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ILP_DFMA 8
#define ILP_TRANS 4

/* Result every thread with (threadIdx & 31) == lane stores in b9gw_dfma_peak. */
double b9ref_dfma_lane(int lane, int ilp, double a, double b, int iters) {
    double x[ILP_DFMA];
    if (ilp < 1 || ilp > ILP_DFMA) return NAN;
    for (int j = 0; j < ilp; ++j) x[j] = 1.0 + 0.125 * j + lane * 0x1p-10;
    for (int i = 0; i < iters; ++i)
        for (int j = 0; j < ilp; ++j) x[j] = fma(x[j], a, b);
    double s = x[0];
    for (int j = 1; j < ilp; ++j) s += x[j];
    return s;
}

/* The argument generator of the spread variants (groundwork.cu:spread_arg): sign, one of
 * 16 consecutive binary exponents from 2^e_lo, 20 mantissa bits from a 32-bit LCG. */
static double spread_arg(uint32_t *state, int e_lo, uint32_t sign) {
    *state = *state * 1664525u + 1013904223u;
    uint32_t e = (uint32_t)(1023 + e_lo) + (*state >> 28);
    uint64_t bits = (uint64_t)(sign | (e << 20) | ((*state >> 8) & 0xFFFFFu)) << 32;
    double d;
    memcpy(&d, &bits, sizeof d);
    return d;
}

/* The arguments chain j of lane `lane` feeds to exp (which 4) or log (which 5), in order. */
void b9ref_spread_args(int lane, int j, int which, int iters, double *args) {
    uint32_t st = (uint32_t)(lane * ILP_TRANS + j) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i)
        args[i] = which == 4 ? spread_arg(&st, -6, 0x80000000u) : spread_arg(&st, -8, 0u);
}

/* Same for b9gw_transcendental_rate (0 exp(-x), 1 log(x+3), 2 exp10(-0.4x), 3 log10(x+3),
 * 4 sum of exp over the spread, 5 sum of log over the spread). */
double b9ref_trans_lane(int lane, int which, int iters) {
    if (which >= 4) {
        uint32_t st[ILP_TRANS];
        double acc[ILP_TRANS];
        for (int j = 0; j < ILP_TRANS; ++j) {
            st[j] = (uint32_t)(lane * ILP_TRANS + j) * 2654435761u + 12345u;
            acc[j] = 0.0;
        }
        for (int i = 0; i < iters; ++i)
            for (int j = 0; j < ILP_TRANS; ++j)
                acc[j] += which == 4 ? exp(spread_arg(&st[j], -6, 0x80000000u))
                                     : log(spread_arg(&st[j], -8, 0u));
        double t = acc[0];
        for (int j = 1; j < ILP_TRANS; ++j) t += acc[j];
        return t;
    }
    double x[ILP_TRANS];
    for (int j = 0; j < ILP_TRANS; ++j) x[j] = 0.5 + 0.25 * j + lane * 0x1p-8;
    for (int i = 0; i < iters; ++i)
        for (int j = 0; j < ILP_TRANS; ++j)
            x[j] = which == 0   ? exp(-x[j])
                   : which == 1 ? log(x[j] + 3.0)
                   : which == 2 ? pow(10.0, -0.4 * x[j])
                                : log10(x[j] + 3.0);
    double s = x[0];
    for (int j = 1; j < ILP_TRANS; ++j) s += x[j];
    return s;
}

/* which: 0 exp, 1 log, 2 exp10 (glibc's exp10), 3 log10, 12 pow(10, x) — the two host
 * spellings of 10^x are different functions with different error bounds. */
void b9ref_map(int which, const double *x, double *y, long long n) {
    for (long long i = 0; i < n; ++i)
        y[i] = which == 0 ? exp(x[i]) : which == 1 ? log(x[i]) : which == 2 ? exp10(x[i])
               : which == 3 ? log10(x[i]) : pow(10.0, x[i]);
}

/* Row log-sum-exp the way a plain CPU loop writes it: max, then a left-to-right
 * sum of exp(x - max). */
double b9ref_lse_serial(const double *x, long long cols) {
    double m = -INFINITY;
    for (long long c = 0; c < cols; ++c) m = fmax(m, x[c]);
    if (m == -INFINITY) return -INFINITY;
    double s = 0.0;
    for (long long c = 0; c < cols; ++c) s += exp(x[c] - m);
    return m + log(s);
}

/* Row log-sum-exp in the kernel's order: 32 lane-strided partials, then the
 * xor-butterfly (16,8,4,2,1).  After the butterfly every lane holds a sum; the
 * kernel stores lane 0's, so this returns lane 0's. */
double b9ref_lse_warp_order(const double *x, long long cols) {
    double m = -INFINITY;
    for (long long c = 0; c < cols; ++c) m = fmax(m, x[c]);
    if (m == -INFINITY) return -INFINITY;
    double s[32], t[32];
    for (int l = 0; l < 32; ++l) {
        s[l] = 0.0;
        for (long long c = l; c < cols; c += 32) s[l] += exp(x[c] - m);
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ o];
        for (int l = 0; l < 32; ++l) s[l] = t[l];
    }
    return m + log(s[0]);
}

void b9ref_lse_rows(const double *x, long long rows, long long cols, int warp_order,
                    double *row_lse) {
    for (long long r = 0; r < rows; ++r)
        row_lse[r] = warp_order ? b9ref_lse_warp_order(x + r * cols, cols)
                                : b9ref_lse_serial(x + r * cols, cols);
}

double b9ref_serial_sum(const double *v, long long n) {
    double s = 0.0;
    for (long long i = 0; i < n; ++i) s += v[i];
    return s;
}

/* ---- the synthetic term generator of b9gw_generate_terms / b9gw_lse_generated
 * (include/b9_groundwork.h states it; lse.cu:GeneratedRow is the device statement).
 * Every operation is rounded once, as written; -ffp-contract=off keeps it so. */
double b9ref_gen_term(long long row, long long col, long long cols) {
    const double colsd = (double)cols;
    const double inv_cols = 1.0 / colsd;                 /* rounded once, then multiplied */
    const double u = (double)row * 0.6180339887498949;
    const double c0 = (u - floor(u)) * colsd;
    const double w = (double)(34 + (int)(row % 7)) * inv_cols;
    const double b = -(c0 * w);
    const double t = fma((double)col, w, b);
    return -(t * t);
}

void b9ref_generate_terms(long long rows, long long cols, double *x) {
    for (long long r = 0; r < rows; ++r)
        for (long long c = 0; c < cols; ++c) x[r * cols + c] = b9ref_gen_term(r, c, cols);
}

/* ---- world-size-independent sum over virtual shards (include/b9_groundwork.h) */
long long b9ref_shard_lo(long long n, int V, int v) {
    return (long long)(((__int128)v * n) / V);
}

/* Warp-order sum of v[lo..hi): 32 lane-strided serial partials, xor-butterfly, lane 0. */
double b9ref_shard_partial(const double *v, long long lo, long long hi) {
    double s[32], t[32];
    for (int l = 0; l < 32; ++l) {
        s[l] = 0.0;
        for (long long i = lo + l; i < hi; i += 32) s[l] += v[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ o];
        for (int l = 0; l < 32; ++l) s[l] = t[l];
    }
    return s[0];
}

/* values is [chains][n]; partials (may be NULL) receives [V][chains]; total[chains] is
 * (((0 + P[0]) + P[1]) + ...) + P[V-1]. */
void b9ref_vshard_total(const double *values, long long chains, long long n, int V,
                        double *partials, double *total) {
    for (long long c = 0; c < chains; ++c) {
        double acc = 0.0;
        for (int v = 0; v < V; ++v) {
            double p = b9ref_shard_partial(values + c * n, b9ref_shard_lo(n, V, v),
                                           b9ref_shard_lo(n, V, v + 1));
            if (partials) partials[(long long)v * chains + c] = p;
            acc += p;
        }
        total[c] = acc;
    }
}
