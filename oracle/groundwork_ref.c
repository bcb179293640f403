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
 * only IEEE add/fma are involved, and to a stated ULP bound where libm is.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this file's library; the product path never does.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared groundwork_ref.c -lm
 * (-ffp-contract=off: the only fused operations are the explicit fma() calls.)
 */
#include <math.h>
#include <stddef.h>

#define ILP_DFMA 8
#define ILP_TRANS 4

/* Result every thread with (threadIdx & 31) == lane stores in b9gw_dfma_peak. */
double b9ref_dfma_lane(int lane, double a, double b, int iters) {
    double x[ILP_DFMA];
    for (int j = 0; j < ILP_DFMA; ++j) x[j] = 1.0 + 0.125 * j + lane * 0x1p-10;
    for (int i = 0; i < iters; ++i)
        for (int j = 0; j < ILP_DFMA; ++j) x[j] = fma(x[j], a, b);
    double s = x[0];
    for (int j = 1; j < ILP_DFMA; ++j) s += x[j];
    return s;
}

/* Same for b9gw_transcendental_rate (0 exp(-x), 1 log(x+3), 2 exp10(-0.4x), 3 log10(x+3)). */
double b9ref_trans_lane(int lane, int which, int iters) {
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

void b9ref_map(int which, const double *x, double *y, long long n) {
    for (long long i = 0; i < n; ++i) y[i] = which == 0 ? exp(x[i]) : log(x[i]);
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

/* Sum in the order of ordered_sum_kernel: 1024 strided partials + pairwise tree. */
double b9ref_ordered_sum(const double *v, long long n) {
    double p[1024];
    for (int t = 0; t < 1024; ++t) {
        double s = 0.0;
        for (long long i = t; i < n; i += 1024) s += v[i];
        p[t] = s;
    }
    for (int w = 512; w > 0; w >>= 1)
        for (int t = 0; t < w; ++t) p[t] += p[t + w];
    return p[0];
}

double b9ref_serial_sum(const double *v, long long n) {
    double s = 0.0;
    for (long long i = 0; i < n; ++i) s += v[i];
    return s;
}
