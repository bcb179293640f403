// groundwork.cu — reference-independent FP64 groundwork kernels for sm_100a.
//
// The BASE-9 hot path is BLOCKED: /root/reference/README.md:1-4 is a
// relocation notice and no base-cpp source is staged.  Nothing here restates or
// imitates reference code.  These kernels measure what north_star requires
// before a roofline fraction can be quoted for an FP64 likelihood on B200:
// the DFMA peak, exp/log rates, CUDA-libm vs host-libm distance, and how a
// fixed-order warp log-sum-exp compares with a serial CPU one.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=true
// (FMA is explicit via fma(); --fmad only affects a*b+c the compiler finds, and
// the lse kernels contain none on the value path.)

#include <cuda_runtime.h>
#include <chrono>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "b9_groundwork.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}

#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) { rc = fail(B9GW_E_CUDA, #call, e_); goto done; } \
    } while (0)

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B9GW_E_NODEVICE, "no CUDA device visible (no CPU fallback exists)");
    }
    if (device < 0 || device >= n) return fail(B9GW_E_ARG, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "cudaSetDevice", e);
    return B9GW_OK;
}

// ---------------------------------------------------------------- DFMA peak
// Each thread: ILP independent chains, fully unrolled; a and b arrive as kernel
// arguments so nothing folds.  x0 depends on (chain, lane) only, so a CPU
// checker needs 32*ILP chains, not one per thread.
__global__ void __launch_bounds__(B9GW_DFMA_THREADS)
dfma_peak_kernel(double *__restrict__ out, double a, double b, int iters) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double x[B9GW_DFMA_ILP];
#pragma unroll
    for (int j = 0; j < B9GW_DFMA_ILP; ++j)
        x[j] = 1.0 + 0.125 * j + lane * 0x1p-10;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < B9GW_DFMA_ILP; ++j) x[j] = fma(x[j], a, b);
    }
    double s = x[0];
#pragma unroll
    for (int j = 1; j < B9GW_DFMA_ILP; ++j) s = __dadd_rn(s, x[j]);
    out[t] = s;
}

// ------------------------------------------------------- exp / log chain rate
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

template <int WHICH>
__global__ void map_kernel(const double *__restrict__ x, double *__restrict__ y,
                           long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = (WHICH == 0) ? exp(x[i]) : log(x[i]);
}

// ------------------------------------------------------ fixed-order row LSE
// One warp per row, 8 rows per CTA.  Lane-strided reads are 256-byte coalesced
// segments.  The second pass re-reads the row; for the row lengths of interest
// (<= a few thousand columns) that is an L1/L2 hit, not HBM traffic.
constexpr int LSE_WARPS = 8;

__global__ void __launch_bounds__(LSE_WARPS * 32)
lse_rows_kernel(const double *__restrict__ x, long long rows, long long cols,
                double *__restrict__ row_lse) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LSE_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;  // whole warp exits together
    const double *xr = x + row * cols;

    double m = -INFINITY;
    for (long long c = lane; c < cols; c += 32) m = fmax(m, xr[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));

    double r;
    if (m == -INFINITY) {
        r = -INFINITY;  // every term is exp(-inf) = 0 (also covers cols == 0)
    } else {
        double s = 0.0;
        for (long long c = lane; c < cols; c += 32)
            s = __dadd_rn(s, exp(__dsub_rn(xr[c], m)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        r = __dadd_rn(m, log(s));
    }
    if (lane == 0) row_lse[row] = r;
}

// Fixed-order sum of n doubles: 1024 strided serial partials, then a pairwise
// tree in shared memory.  One CTA; the order is a function of n alone.
__global__ void __launch_bounds__(1024)
ordered_sum_kernel(const double *__restrict__ v, long long n, double *__restrict__ out) {
    __shared__ double p[1024];
    const int t = threadIdx.x;
    double s = 0.0;
    for (long long i = t; i < n; i += 1024) s = __dadd_rn(s, v[i]);
    p[t] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (t < w) p[t] = __dadd_rn(p[t], p[t + w]);
        __syncthreads();
    }
    if (t == 0) *out = p[0];
}

__global__ void tick_kernel(double *out, double v) { *out = v; }

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t init() {
        cudaError_t e = cudaEventCreate(&a);
        return e != cudaSuccess ? e : cudaEventCreate(&b);
    }
    ~Timer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

int sm_count_of(int device, int *sms) {
    cudaError_t e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "cudaDeviceGetAttribute(SM count)", e);
    return B9GW_OK;
}

// Shared body of the two chain benchmarks: launch(grid, stream) runs one launch.
template <class Launch>
int run_chain_bench(int device, int ctas_per_sm, int iters, int warmup, int reps,
                    double *out_host, long long *n_threads, float *ms_per_launch,
                    Launch launch) {
    int rc = B9GW_OK, sms = 0;
    double *d_out = nullptr;
    cudaStream_t st = nullptr;
    Timer tm;
    float ms = 0.f;
    long long nthr = 0;
    if (ctas_per_sm < 1 || ctas_per_sm > 32 || iters < 1 || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need 1<=ctas_per_sm<=32, iters>=1, warmup>=0, reps>=1");
    if ((rc = select_device(device)) != B9GW_OK) return rc;
    if ((rc = sm_count_of(device, &sms)) != B9GW_OK) return rc;
    {
        const int grid = sms * ctas_per_sm;
        nthr = (long long)grid * B9GW_DFMA_THREADS;
        CK(cudaMalloc(&d_out, nthr * sizeof(double)));
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

int b9gw_abi_version(void) { return 2; }

const char *b9gw_last_error(void) { return g_err; }

int b9gw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b9gw_device_info(int device, int *sm_count, int *sm_clock_mhz, long long *l2_bytes) {
    int rc = select_device(device);
    if (rc != B9GW_OK) return rc;
    int v = 0;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = v;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device));
    if (sm_clock_mhz) *sm_clock_mhz = v / 1000;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    if (l2_bytes) *l2_bytes = v;
done:
    return rc;
}

int b9gw_dfma_peak(int device, int ctas_per_sm, int iters, double a, double b,
                   int warmup, int reps, double *out_host, long long *n_threads,
                   float *ms_per_launch, double *tflops) {
    long long nthr = 0;
    float ms = 0.f;
    int rc = run_chain_bench(
        device, ctas_per_sm, iters, warmup, reps, out_host, &nthr, &ms,
        [=](int grid, cudaStream_t st, double *out) {
            dfma_peak_kernel<<<grid, B9GW_DFMA_THREADS, 0, st>>>(out, a, b, iters);
        });
    if (rc != B9GW_OK) return rc;
    if (n_threads) *n_threads = nthr;
    if (ms_per_launch) *ms_per_launch = ms;
    if (tflops) *tflops = 2.0 * B9GW_DFMA_ILP * (double)iters * (double)nthr / (ms * 1e-3) * 1e-12;
    return rc;
}

int b9gw_transcendental_rate(int device, int which, int ctas_per_sm, int iters,
                             int warmup, int reps, double *out_host,
                             long long *n_threads, float *ms_per_launch,
                             double *gevals_per_s) {
    if (which < 0 || which > 3) return fail(B9GW_E_ARG, "which must be 0 exp, 1 log, 2 exp10, 3 log10");
    long long nthr = 0;
    float ms = 0.f;
    int rc = run_chain_bench(
        device, ctas_per_sm, iters, warmup, reps, out_host, &nthr, &ms,
        [=](int grid, cudaStream_t st, double *out) {
            switch (which) {
                case 0: trans_rate_kernel<0><<<grid, B9GW_DFMA_THREADS, 0, st>>>(out, iters); break;
                case 1: trans_rate_kernel<1><<<grid, B9GW_DFMA_THREADS, 0, st>>>(out, iters); break;
                case 2: trans_rate_kernel<2><<<grid, B9GW_DFMA_THREADS, 0, st>>>(out, iters); break;
                default: trans_rate_kernel<3><<<grid, B9GW_DFMA_THREADS, 0, st>>>(out, iters);
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
    if (warmup < 0 || reps < 1) return fail(B9GW_E_ARG, "need warmup>=0, reps>=1");
    if ((rc = select_device(device)) != B9GW_OK) return rc;
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
        tick_kernel<<<1, 1, 0, st>>>(d, -1.0);
        CK(cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, st));
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
    if (which != 0 && which != 1) return fail(B9GW_E_ARG, "which must be 0 (exp) or 1 (log)");
    if (n < 0 || (n > 0 && (!x_host || !y_host))) return fail(B9GW_E_ARG, "null buffer or n<0");
    if ((rc = select_device(device)) != B9GW_OK) return rc;
    if (n == 0) return B9GW_OK;
    if ((rc = sm_count_of(device, &sms)) != B9GW_OK) return rc;
    CK(cudaMalloc(&dx, n * sizeof(double)));
    CK(cudaMalloc(&dy, n * sizeof(double)));
    CK(cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
    if (which == 0)
        map_kernel<0><<<sms * 8, 256>>>(dx, dy, n);
    else
        map_kernel<1><<<sms * 8, 256>>>(dx, dy, n);
    CK(cudaGetLastError());
    CK(cudaMemcpy(y_host, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
done:
    if (dx) cudaFree(dx);
    if (dy) cudaFree(dy);
    return rc;
}

int b9gw_lse_rows(int device, const double *x_host, long long rows, long long cols,
                  int warmup, int reps, double *row_lse_host, double *total_host,
                  float *ms_per_launch) {
    int rc = B9GW_OK;
    double *dx = nullptr, *dr = nullptr, *dt = nullptr;
    cudaStream_t st = nullptr;
    Timer tm;
    float ms = 0.f;
    if (rows < 0 || cols < 0 || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need rows>=0, cols>=0, warmup>=0, reps>=1");
    if (rows * cols > 0 && !x_host) return fail(B9GW_E_ARG, "x_host is null");
    if (!total_host) return fail(B9GW_E_ARG, "total_host is null");
    if ((rows + LSE_WARPS - 1) / LSE_WARPS > 0x7fffffffLL) return fail(B9GW_E_ARG, "too many rows");
    if ((rc = select_device(device)) != B9GW_OK) return rc;
    {
        const long long n = rows * cols;
        const unsigned grid = (unsigned)((rows + LSE_WARPS - 1) / LSE_WARPS);
        CK(cudaMalloc(&dx, (n > 0 ? n : 1) * sizeof(double)));
        CK(cudaMalloc(&dr, (rows > 0 ? rows : 1) * sizeof(double)));
        CK(cudaMalloc(&dt, sizeof(double)));
        if (n > 0) CK(cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            if (grid > 0)
                lse_rows_kernel<<<grid, LSE_WARPS * 32, 0, st>>>(dx, rows, cols, dr);
            ordered_sum_kernel<<<1, 1024, 0, st>>>(dr, rows, dt);
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (row_lse_host && rows > 0)
            CK(cudaMemcpy(row_lse_host, dr, rows * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(total_host, dt, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (ms_per_launch) *ms_per_launch = ms / reps;
done:
    if (st) cudaStreamDestroy(st);
    if (dx) cudaFree(dx);
    if (dr) cudaFree(dr);
    if (dt) cudaFree(dt);
    return rc;
}

}  // extern "C"
