// lse.cu — fixed-order FP64 log-sum-exp over rows, for sm_100a.
//
// Reference-independent groundwork (the BASE-9 hot path is BLOCKED;
// nothing here restates or imitates reference code).  One warp owns a row.  The
// ORDER is the contract, pinned bit for bit by oracle/groundwork_ref.c:
//   max   : exact, any order;
//   sum   : lane l adds exp(x[c] - max) for c = l, l+32, l+64, ... in increasing
//           c, starting from +0; the 32 lane sums are combined by an
//           xor-butterfly with offsets 16, 8, 4, 2, 1;
//   value : max + log(sum), or -inf when max == -inf;
//   total : sum of the row values in the order of b9ref_ordered_sum.
// Two term sources share that code: a matrix in memory (lse_rows) and a
// closed-form generator evaluated in registers (lse_generated), so the second
// shows what the fixed order costs when terms never travel through memory.
//
// Every value-path operation is an explicit __dadd_rn/__dsub_rn/__dmul_rn/fma,
// so -fmad cannot contract anything the host checker does not.

#include <math.h>

#include "common.cuh"

namespace {

using b9gw::fail;

constexpr int LSE_WARPS = 8;              // rows per CTA
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------- term sources
struct MatrixRow {                        // terms live in memory
    const double *p;
    __device__ double operator()(long long c) const { return __ldg(p + c); }
};

struct GeneratedRow {                     // terms are arithmetic on (row, col)
    double w, b;                          // t = fma(c, w, b), term = -(t*t)
    __device__ GeneratedRow(long long row, long long cols) {
        const double colsd = (double)cols;
        const double u = __dmul_rn((double)row, 0.6180339887498949);
        const double c0 = __dmul_rn(__dsub_rn(u, floor(u)), colsd);
        w = __ddiv_rn((double)(34 + (int)(row % 7)), colsd);
        b = -__dmul_rn(c0, w);
    }
    __device__ double operator()(long long c) const {
        const double t = fma((double)c, w, b);
        return __dmul_rn(-t, t);
    }
};

__device__ __forceinline__ double warp_max(double m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    return m;
}

__device__ __forceinline__ double warp_add(double s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(FULL, s, o));
    return s;
}

// Row of at most 32*TPL columns: every term is fetched/generated once, all of a
// lane's TPL fetches are independent (loads in flight together), and the terms
// stay in registers between the max pass and the exp pass.
template <int TPL, class Src>
__device__ __forceinline__ double warp_lse_regs(const Src &src, long long cols, int lane) {
    double v[TPL];
#pragma unroll
    for (int k = 0; k < TPL; ++k) {
        const long long c = lane + 32 * k;
        v[k] = c < cols ? src(c) : -INFINITY;
    }
    double m = -INFINITY;
#pragma unroll
    for (int k = 0; k < TPL; ++k) m = fmax(m, v[k]);
    m = warp_max(m);
    if (m == -INFINITY) return -INFINITY;   // every term is exp(-inf) = 0 (also cols == 0)
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < TPL; ++k)
        if (lane + 32 * k < cols) s = __dadd_rn(s, exp(__dsub_rn(v[k], m)));
    return __dadd_rn(m, log(warp_add(s)));
}

// Longer rows: two passes over the source, 8 independent fetches per lane at a time.
template <class Src>
__device__ __forceinline__ double warp_lse_stream(const Src &src, long long cols, int lane) {
    constexpr int B = 8;
    double m = -INFINITY;
    for (long long c0 = lane; c0 < cols; c0 += 32 * B) {
        double v[B];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const long long c = c0 + 32 * k;
            v[k] = c < cols ? src(c) : -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < B; ++k) m = fmax(m, v[k]);
    }
    m = warp_max(m);
    if (m == -INFINITY) return -INFINITY;
    double s = 0.0;
    for (long long c0 = lane; c0 < cols; c0 += 32 * B) {
        double v[B];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const long long c = c0 + 32 * k;
            v[k] = c < cols ? src(c) : -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < B; ++k)
            if (c0 + 32 * k < cols) s = __dadd_rn(s, exp(__dsub_rn(v[k], m)));
    }
    return __dadd_rn(m, log(warp_add(s)));
}

// Fixed-order sum of n doubles by one CTA of 256 threads, in the order of
// b9ref_ordered_sum: 1024 strided serial partials (thread t owns partials t,
// t+256, t+512, t+768), then a pairwise tree over the 1024.
__device__ double cta_ordered_sum(const double *v, long long n, double *p /* shared[1024] */) {
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = t + 256 * q;
        double s = 0.0;
        for (long long i = j; i < n; i += 1024) s = __dadd_rn(s, __ldcg(v + i));
        p[j] = s;
    }
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        for (int j = t; j < w; j += 256) p[j] = __dadd_rn(p[j], p[j + w]);
        __syncthreads();
    }
    return p[0];
}

// SRC 0: matrix, 1: generator.  TPL 0 selects the streaming (two-pass) path.
// The last CTA to retire adds the row values in the fixed order, so the total
// needs no second launch; its order does not depend on which CTA that is.
template <int SRC, int TPL>
__global__ void __launch_bounds__(LSE_WARPS * 32)
lse_kernel(const double *__restrict__ x, long long rows, long long cols,
           double *__restrict__ row_lse, double *__restrict__ total,
           unsigned *__restrict__ ticket) {
    __shared__ double p[1024];
    __shared__ bool last;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LSE_WARPS + (threadIdx.x >> 5);
    if (row < rows) {                      // whole warp takes the same branch
        double r;
        if constexpr (SRC == 0) {
            const MatrixRow src{x + row * cols};
            if constexpr (TPL > 0) r = warp_lse_regs<TPL>(src, cols, lane);
            else r = warp_lse_stream(src, cols, lane);
        } else {
            const GeneratedRow src(row, cols);
            if constexpr (TPL > 0) r = warp_lse_regs<TPL>(src, cols, lane);
            else r = warp_lse_stream(src, cols, lane);
        }
        if (lane == 0) row_lse[row] = r;
    }
    __threadfence();                       // row values visible before the ticket
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const double s = cta_ordered_sum(row_lse, rows, p);
    if (threadIdx.x == 0) {
        *total = s;
        *ticket = 0;                       // ready for the next launch on this stream
    }
}

__global__ void __launch_bounds__(256)
generate_terms_kernel(double *__restrict__ x, long long rows, long long cols) {
    const long long n = rows * cols;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long r = i / cols, c = i - r * cols;
        x[i] = GeneratedRow(r, cols)(c);
    }
}

__global__ void zero_total_kernel(double *total) { *total = 0.0; }

template <int SRC>
void launch_lse(unsigned grid, cudaStream_t st, const double *x, long long rows, long long cols,
                double *row_lse, double *total, unsigned *ticket) {
    constexpr int T = LSE_WARPS * 32;
    if (cols <= 128)
        lse_kernel<SRC, 4><<<grid, T, 0, st>>>(x, rows, cols, row_lse, total, ticket);
    else if (cols <= 512)
        lse_kernel<SRC, 16><<<grid, T, 0, st>>>(x, rows, cols, row_lse, total, ticket);
    else if (cols <= B9GW_LSE_REG_COLS)
        lse_kernel<SRC, 32><<<grid, T, 0, st>>>(x, rows, cols, row_lse, total, ticket);
    else
        lse_kernel<SRC, 0><<<grid, T, 0, st>>>(x, rows, cols, row_lse, total, ticket);
}

// Shared host body: SRC 0 uploads x_host, SRC 1 has no input at all.
template <int SRC>
int run_lse(int device, const double *x_host, long long rows, long long cols, int warmup, int reps,
            double *row_lse_host, double *total_host, float *ms_per_launch) {
    int rc = B9GW_OK;
    double *dx = nullptr, *dr = nullptr, *dt = nullptr;
    unsigned *dticket = nullptr;
    cudaStream_t st = nullptr;
    b9gw::Timer tm;
    float ms = 0.f;
    if (rows < 0 || cols < 0 || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need rows>=0, cols>=0, warmup>=0, reps>=1");
    if (!b9gw::product_ok(rows, cols) || !b9gw::count_ok(rows))
        return fail(B9GW_E_ARG, "rows*cols overflows");
    if (SRC == 0 && rows * cols > 0 && !x_host) return fail(B9GW_E_ARG, "x_host is null");
    if (!total_host) return fail(B9GW_E_ARG, "total_host is null");
    if ((rows + LSE_WARPS - 1) / LSE_WARPS > 0x7fffffffLL) return fail(B9GW_E_ARG, "too many rows");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    {
        const long long n = rows * cols;
        const unsigned grid = (unsigned)((rows + LSE_WARPS - 1) / LSE_WARPS);
        if (SRC == 0) {
            CK(cudaMalloc(&dx, (n > 0 ? n : 1) * sizeof(double)));
            if (n > 0) CK(cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
        }
        CK(cudaMalloc(&dr, (rows > 0 ? rows : 1) * sizeof(double)));
        CK(cudaMalloc(&dt, sizeof(double)));
        CK(cudaMalloc(&dticket, sizeof(unsigned)));
        CK(cudaMemset(dticket, 0, sizeof(unsigned)));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            if (grid > 0)
                launch_lse<SRC>(grid, st, dx, rows, cols, dr, dt, dticket);
            else
                zero_total_kernel<<<1, 1, 0, st>>>(dt);   // no rows: the empty sum
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
    if (dticket) cudaFree(dticket);
    return rc;
}

}  // namespace

extern "C" {

int b9gw_lse_rows(int device, const double *x_host, long long rows, long long cols,
                  int warmup, int reps, double *row_lse_host, double *total_host,
                  float *ms_per_launch) {
    return run_lse<0>(device, x_host, rows, cols, warmup, reps, row_lse_host, total_host,
                      ms_per_launch);
}

int b9gw_lse_generated(int device, long long rows, long long cols, int warmup, int reps,
                       double *row_lse_host, double *total_host, float *ms_per_launch) {
    return run_lse<1>(device, nullptr, rows, cols, warmup, reps, row_lse_host, total_host,
                      ms_per_launch);
}

int b9gw_generate_terms(int device, long long rows, long long cols, double *x_host) {
    int rc = B9GW_OK, sms = 0;
    double *dx = nullptr;
    if (rows < 0 || cols < 0 || !b9gw::product_ok(rows, cols))
        return fail(B9GW_E_ARG, "need rows>=0, cols>=0 and rows*cols representable");
    if (rows * cols > 0 && !x_host) return fail(B9GW_E_ARG, "x_host is null");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    if (rows * cols == 0) return B9GW_OK;
    if ((rc = b9gw::sm_count_of(device, &sms)) != B9GW_OK) return rc;
    CK(cudaMalloc(&dx, rows * cols * sizeof(double)));
    generate_terms_kernel<<<sms * 8, 256>>>(dx, rows, cols);
    CK(cudaGetLastError());
    CK(cudaMemcpy(x_host, dx, rows * cols * sizeof(double), cudaMemcpyDeviceToHost));
done:
    if (dx) cudaFree(dx);
    return rc;
}

}  // extern "C"
