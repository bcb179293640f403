// groundwork.cu — reference-independent FP64 rate and latency probes for sm_100a.
//
// The BASE-9 hot path is BLOCKED: /root/reference/README.md:1-4 is a
// relocation notice and no base-cpp source is staged.  Nothing here restates or
// imitates reference code.  These kernels measure what north_star requires
// before a roofline fraction can be quoted for an FP64 likelihood on B200:
// the DFMA peak, exp/log/exp10/log10 rates (on one argument and on a spread),
// CUDA-libm vs host-libm distance, and the host round trip of a dependent step.
// The fixed-order log-sum-exp lives in lse.cu, the cross-rank sum in vshard.cu.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
// (FMA is explicit via fma(); value paths use __dadd_rn/__dmul_rn, which -fmad
// never contracts.)

#include <chrono>
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b9gw {
char *err_buf() {
    thread_local char buf[kErrBytes] = "";
    return buf;
}
}  // namespace b9gw

namespace {

using b9gw::fail;
using b9gw::sm_count_of;
using b9gw::Timer;

// ---------------------------------------------------------------- DFMA peak
// Each thread: ILP independent chains, fully unrolled; a and b arrive as kernel
// arguments so nothing folds.  x0 depends on (chain, lane) only, so a CPU
// checker needs 32*ILP chains, not one per thread.  ILP 8 is the peak measurement;
// ILP 1/2/4 at one CTA per SM (two warps per scheduler) expose the dependent-issue
// latency: with c chains in flight per scheduler the pipe retires c DFMAs per latency.
// FILL 1/2 adds that many independent integer multiply-adds per DFMA: it measures whether a
// non-FP64 instruction can issue in the shadow of a DFMA (it cannot: profiles/r02_groundwork.md).
// FILL > 0: that many independent integer instructions per DFMA, KIND 0 = multiply-adds
// (n <- n*n + c, one IMAD, a half-rate instruction like DFMA but on another pipe), KIND 1 =
// add/xor pairs (n <- n + m; m <- m ^ n, full-rate ALU instructions; FILL is even).
template <int ILP, int FILL, int KIND>
__global__ void __launch_bounds__(B9GW_DFMA_THREADS)
dfma_peak_kernel(double *__restrict__ out, double a, double b, int iters) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double x[ILP];
    unsigned n[ILP], q[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        x[j] = 1.0 + 0.125 * j + lane * 0x1p-10;
        n[j] = (unsigned)t + j;
        q[j] = (unsigned)t * 2654435761u + j;
    }
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            x[j] = fma(x[j], a, b);
            if (KIND == 0) {
#pragma unroll
                for (int f = 0; f < FILL; ++f) n[j] = n[j] * n[j] + 1013904223u;   // squares do not fold
            } else {
#pragma unroll
                for (int f = 0; f < FILL / 2; ++f) {
                    n[j] += q[j];
                    q[j] ^= n[j];
                }
            }
        }
    }
    double s = x[0];
    unsigned m = n[0] ^ q[0];
#pragma unroll
    for (int j = 1; j < ILP; ++j) {
        s = __dadd_rn(s, x[j]);
        m ^= n[j] ^ q[j];
    }
    // the integer work must stay live without touching the checked value: one output in 2^32
    // would be replaced, and only when FILL > 0 and iters is large enough to reach the pattern
    out[t] = (FILL > 0 && m == 0x9e3779b9u && iters < 0) ? -1.0 : s;
    if (FILL > 0 && m == 0x9e3779b9u) out[t + (long long)gridDim.x * blockDim.x] = 0.0;
}

// ------------------------------------------------------- exp / log chain rate
// WHICH 0..3 are contractions: after a few dozen iterations each thread sits at
// the map's fixed point, so the rate is a single-argument, mid-range rate.
template <int WHICH>
__global__ void __launch_bounds__(B9GW_DFMA_THREADS)
trans_rate_kernel(double *__restrict__ out, int iters) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double x[B9GW_TRANS_ILP];
#pragma unroll
    for (int j = 0; j < B9GW_TRANS_ILP; ++j)
        x[j] = 0.5 + 0.25 * j + lane * 0x1p-8;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < B9GW_TRANS_ILP; ++j)
            x[j] = (WHICH == 0)   ? exp(-x[j])
                   : (WHICH == 1) ? log(__dadd_rn(x[j], 3.0))
                   : (WHICH == 2) ? exp10(__dmul_rn(-0.4, x[j]))
                                  : log10(__dadd_rn(x[j], 3.0));
    }
    double s = x[0];
#pragma unroll
    for (int j = 1; j < B9GW_TRANS_ILP; ++j) s = __dadd_rn(s, x[j]);
    out[t] = s;
}

// A fresh argument per evaluation, assembled on the integer pipe: sign, one of
// 16 consecutive binary exponents starting at 2^e_lo, and 20 mantissa bits from
// a 32-bit LCG.  The only FP64 work besides the function is one DADD.
__device__ __forceinline__ double spread_arg(unsigned &state, int e_lo, unsigned sign) {
    state = state * 1664525u + 1013904223u;
    const unsigned e = (unsigned)(1023 + e_lo) + (state >> 28);
    return __hiloint2double((int)(sign | (e << 20) | ((state >> 8) & 0xFFFFFu)), 0);
}

template <int WHICH>   // 4: exp(-(2^[-6,9] * 1.f)),  5: log(2^[-8,7] * 1.f)
__global__ void __launch_bounds__(B9GW_DFMA_THREADS)
spread_rate_kernel(double *__restrict__ out, int iters) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned st[B9GW_TRANS_ILP];
    double acc[B9GW_TRANS_ILP];
#pragma unroll
    for (int j = 0; j < B9GW_TRANS_ILP; ++j) {
        st[j] = (unsigned)(lane * B9GW_TRANS_ILP + j) * 2654435761u + 12345u;
        acc[j] = 0.0;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < B9GW_TRANS_ILP; ++j) {
            const double v = (WHICH == 4) ? exp(spread_arg(st[j], -6, 0x80000000u))
                                          : log(spread_arg(st[j], -8, 0u));
            acc[j] = __dadd_rn(acc[j], v);
        }
    }
    double s = acc[0];
#pragma unroll
    for (int j = 1; j < B9GW_TRANS_ILP; ++j) s = __dadd_rn(s, acc[j]);
    out[t] = s;
}

template <int WHICH>
__global__ void map_kernel(const double *__restrict__ x, double *__restrict__ y,
                           long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        y[i] = (WHICH == 0) ? exp(x[i]) : (WHICH == 1) ? log(x[i])
               : (WHICH == 2) ? exp10(x[i]) : log10(x[i]);
}

// The LSE kernel's branch-free copy of exp's fast path (common.cuh), on its own, so a
// test can hold it against exp() bit for bit.  Only meaningful for -708 < x <= 0.
__global__ void map_exp_fast_path_kernel(const __grid_constant__ b9gw::ExpConstants K,
                                         const double *__restrict__ x, double *__restrict__ y,
                                         long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double in[1] = {-x[i]};
        double out[1];
        b9gw::exp_fast_path(K, in, out);
        y[i] = out[0];
    }
}

__global__ void tick_kernel(double *out, double v) { *out = v; }

// Shared body of the chain benchmarks: launch(grid, stream, out) runs one launch.
template <class Launch>
int run_chain_bench(int device, int ctas_per_sm, int iters, int warmup, int reps,
                    double *out_host, long long *n_threads, float *ms_per_launch,
                    Launch launch, int out_factor = 1) {
    int rc = B9GW_OK, sms = 0;
    double *d_out = nullptr;
    cudaStream_t st = nullptr;
    Timer tm;
    float ms = 0.f;
    long long nthr = 0;
    if (ctas_per_sm < 1 || ctas_per_sm > 32 || iters < 1 || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need 1<=ctas_per_sm<=32, iters>=1, warmup>=0, reps>=1");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    if ((rc = sm_count_of(device, &sms)) != B9GW_OK) return rc;
    {
        const int grid = sms * ctas_per_sm;
        nthr = (long long)grid * B9GW_DFMA_THREADS;
        CK(cudaMalloc(&d_out, nthr * out_factor * sizeof(double)));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        for (int i = 0; i < warmup; ++i) launch(grid, st, d_out);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(tm.a, st));
        for (int i = 0; i < reps; ++i) launch(grid, st, d_out);
        CK(cudaEventRecord(tm.b, st));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (out_host)
            CK(cudaMemcpy(out_host, d_out, nthr * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (n_threads) *n_threads = nthr;
    if (ms_per_launch) *ms_per_launch = ms / reps;
done:
    if (st) cudaStreamDestroy(st);
    if (d_out) cudaFree(d_out);
    return rc;
}

}  // namespace

extern "C" {

int b9gw_abi_version(void) { return 5; }

const char *b9gw_last_error(void) { return b9gw::err_buf(); }

int b9gw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b9gw_device_info(int device, int *sm_count, int *sm_clock_mhz, long long *l2_bytes) {
    int rc = B9GW_OK, v = 0;
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = v;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device));
    if (sm_clock_mhz) *sm_clock_mhz = v / 1000;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    if (l2_bytes) *l2_bytes = v;
done:
    return rc;
}

int b9gw_dev_malloc(int device, long long bytes, void **ptr_dev) {
    if (!ptr_dev || bytes < 0) return fail(B9GW_E_ARG, "need ptr_dev and bytes>=0");
    *ptr_dev = nullptr;
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    const size_t n = bytes > 0 ? (size_t)bytes : 1;
    cudaError_t e = cudaMalloc(ptr_dev, n);
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "cudaMalloc", e);
    e = cudaMemset(*ptr_dev, 0, n);          // zero-filled, and finished before we return
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(*ptr_dev);
        *ptr_dev = nullptr;
        return fail(B9GW_E_CUDA, "cudaMemset", e);
    }
    return B9GW_OK;
}

int b9gw_dev_free(int device, void *ptr_dev) {
    if (!ptr_dev) return B9GW_OK;
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    cudaError_t e = cudaFree(ptr_dev);
    return e == cudaSuccess ? B9GW_OK : fail(B9GW_E_CUDA, "cudaFree", e);
}

int b9gw_memcpy_h2d(int device, void *dst_dev, const void *src_host, long long bytes) {
    if (bytes < 0 || (bytes > 0 && (!dst_dev || !src_host))) return fail(B9GW_E_ARG, "null buffer or bytes<0");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    cudaError_t e = cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? B9GW_OK : fail(B9GW_E_CUDA, "cudaMemcpy H2D", e);
}

int b9gw_memcpy_d2h(int device, void *dst_host, const void *src_dev, long long bytes) {
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) return fail(B9GW_E_ARG, "null buffer or bytes<0");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    cudaError_t e = cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? B9GW_OK : fail(B9GW_E_CUDA, "cudaMemcpy D2H", e);
}

int b9gw_dfma_peak(int device, int ctas_per_sm, int ilp, int int_per_fma, int iters, double a,
                   double b, int warmup, int reps, double *out_host, long long *n_threads,
                   float *ms_per_launch, double *tflops) {
    long long nthr = 0;
    float ms = 0.f;
    if (ilp != 1 && ilp != 2 && ilp != 4 && ilp != 8) return fail(B9GW_E_ARG, "ilp must be 1, 2, 4 or 8");
    if ((int_per_fma != 0 && int_per_fma != 1 && int_per_fma != 2 && int_per_fma != -2 && int_per_fma != -4) ||
        (int_per_fma != 0 && ilp != 8))
        return fail(B9GW_E_ARG, "int_per_fma must be 0, 1, 2 (multiply-adds) or -2, -4 (ALU ops), and needs ilp = 8");
    int rc = run_chain_bench(
        device, ctas_per_sm, iters, warmup, reps, out_host, &nthr, &ms,
        [=](int grid, cudaStream_t st, double *out) {
            constexpr int T = B9GW_DFMA_THREADS;
            switch (ilp * 10 + int_per_fma) {
                case 10: dfma_peak_kernel<1, 0, 0><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 20: dfma_peak_kernel<2, 0, 0><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 40: dfma_peak_kernel<4, 0, 0><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 81: dfma_peak_kernel<8, 1, 0><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 82: dfma_peak_kernel<8, 2, 0><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 78: dfma_peak_kernel<8, 2, 1><<<grid, T, 0, st>>>(out, a, b, iters); break;
                case 76: dfma_peak_kernel<8, 4, 1><<<grid, T, 0, st>>>(out, a, b, iters); break;
                default: dfma_peak_kernel<8, 0, 0><<<grid, T, 0, st>>>(out, a, b, iters);
            }
        },
        int_per_fma != 0 ? 2 : 1);
    if (rc != B9GW_OK) return rc;
    if (n_threads) *n_threads = nthr;
    if (ms_per_launch) *ms_per_launch = ms;
    if (tflops) *tflops = 2.0 * ilp * (double)iters * (double)nthr / (ms * 1e-3) * 1e-12;
    return rc;
}

int b9gw_transcendental_rate(int device, int which, int ctas_per_sm, int iters,
                             int warmup, int reps, double *out_host,
                             long long *n_threads, float *ms_per_launch,
                             double *gevals_per_s) {
    if (which < 0 || which > 5)
        return fail(B9GW_E_ARG, "which must be 0 exp, 1 log, 2 exp10, 3 log10, 4 exp-spread, 5 log-spread");
    long long nthr = 0;
    float ms = 0.f;
    int rc = run_chain_bench(
        device, ctas_per_sm, iters, warmup, reps, out_host, &nthr, &ms,
        [=](int grid, cudaStream_t st, double *out) {
            constexpr int T = B9GW_DFMA_THREADS;
            switch (which) {
                case 0: trans_rate_kernel<0><<<grid, T, 0, st>>>(out, iters); break;
                case 1: trans_rate_kernel<1><<<grid, T, 0, st>>>(out, iters); break;
                case 2: trans_rate_kernel<2><<<grid, T, 0, st>>>(out, iters); break;
                case 3: trans_rate_kernel<3><<<grid, T, 0, st>>>(out, iters); break;
                case 4: spread_rate_kernel<4><<<grid, T, 0, st>>>(out, iters); break;
                default: spread_rate_kernel<5><<<grid, T, 0, st>>>(out, iters);
            }
        });
    if (rc != B9GW_OK) return rc;
    if (n_threads) *n_threads = nthr;
    if (ms_per_launch) *ms_per_launch = ms;
    if (gevals_per_s)
        *gevals_per_s = (double)B9GW_TRANS_ILP * (double)iters * (double)nthr / (ms * 1e-3) * 1e-9;
    return rc;
}

int b9gw_step_latency(int device, int warmup, int reps, float *us_launch_sync,
                      float *us_launch_d2h_sync, float *us_graph_d2h_sync) {
    using clk = std::chrono::steady_clock;
    int rc = B9GW_OK;
    double *d = nullptr, *h = nullptr;
    cudaStream_t st = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool capturing = false;
    if (warmup < 0 || reps < 1) return fail(B9GW_E_ARG, "need warmup>=0, reps>=1");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    {
        CK(cudaMalloc(&d, sizeof(double)));
        CK(cudaMallocHost(&h, sizeof(double)));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        auto us = [&](clk::time_point a) {
            return (float)(std::chrono::duration<double, std::micro>(clk::now() - a).count() / reps);
        };
        clk::time_point t0;
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) t0 = clk::now();
            tick_kernel<<<1, 1, 0, st>>>(d, (double)i);
            CK(cudaStreamSynchronize(st));
        }
        if (us_launch_sync) *us_launch_sync = us(t0);
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) t0 = clk::now();
            tick_kernel<<<1, 1, 0, st>>>(d, (double)i);
            CK(cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (*h != (double)i) { rc = fail(B9GW_E_CUDA, "step_latency: stale read-back"); goto done; }
        }
        if (us_launch_d2h_sync) *us_launch_d2h_sync = us(t0);
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        tick_kernel<<<1, 1, 0, st>>>(d, -1.0);
        CK(cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, st));
        capturing = false;                 // EndCapture ends it whether or not it succeeds
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) t0 = clk::now();
            CK(cudaGraphLaunch(exec, st));
            CK(cudaStreamSynchronize(st));
        }
        if (us_graph_d2h_sync) *us_graph_d2h_sync = us(t0);
        if (*h != -1.0) { rc = fail(B9GW_E_CUDA, "step_latency: graph read-back wrong"); goto done; }
    }
done:
    if (capturing) {                       // a call failed mid-capture: do not destroy a capturing stream
        cudaGraph_t dead = nullptr;
        cudaStreamEndCapture(st, &dead);
        if (dead) cudaGraphDestroy(dead);
        cudaGetLastError();
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (st) cudaStreamDestroy(st);
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    return rc;
}

int b9gw_map(int device, int which, const double *x_host, double *y_host, long long n) {
    int rc = B9GW_OK, sms = 0;
    double *dx = nullptr, *dy = nullptr;
    if (which < 0 || which > 4)
        return fail(B9GW_E_ARG, "which must be 0 exp, 1 log, 2 exp10, 3 log10, 4 exp fast path");
    if (!b9gw::count_ok(n) || (n > 0 && (!x_host || !y_host)))
        return fail(B9GW_E_ARG, "null buffer, n<0 or n*8 overflows");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    if (n == 0) return B9GW_OK;
    if ((rc = sm_count_of(device, &sms)) != B9GW_OK) return rc;
    CK(cudaMalloc(&dx, n * sizeof(double)));
    CK(cudaMalloc(&dy, n * sizeof(double)));
    CK(cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
    switch (which) {
        case 0: map_kernel<0><<<sms * 8, 256>>>(dx, dy, n); break;
        case 1: map_kernel<1><<<sms * 8, 256>>>(dx, dy, n); break;
        case 2: map_kernel<2><<<sms * 8, 256>>>(dx, dy, n); break;
        case 3: map_kernel<3><<<sms * 8, 256>>>(dx, dy, n); break;
        default: map_exp_fast_path_kernel<<<sms * 8, 256>>>(b9gw::exp_constants(), dx, dy, n);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(y_host, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
done:
    if (dx) cudaFree(dx);
    if (dy) cudaFree(dy);
    return rc;
}

}  // extern "C"
