/*
 * b9_groundwork.h — C-ABI of libb9_groundwork.so
 *
 * STATUS: the BASE-9 hot path is BLOCKED.  /root/reference is a
 * 4-line relocation notice (/root/reference/README.md:1-4); the base-cpp source
 * it points to is not staged offline.  BASELINE.json's north_star forbids
 * reconstructing the reference from memory, so NOTHING in this header is, or
 * claims to be, the drop-in boundary for BASE-9's cluster log-likelihood.  That
 * boundary can only be declared after the chain driver's call into the
 * likelihood has been read from source.
 *
 * What this header does declare: the reference-independent groundwork that
 * north_star asks for before any roofline fraction can be quoted —
 *   - the FP64 (DFMA) vector peak of the B200, absent from MEASURED_PEAKS.json;
 *   - FP64 exp()/log() issue rates, the other pipe a likelihood will sit on;
 *   - CUDA-libm exp/log versus host libm, element by element (ULP distance);
 *   - a fixed-order FP64 log-sum-exp over rows and a fixed-order sum over rows,
 *     to measure how far reduction order + libm differences move a result that
 *     must later agree with a serial CPU loop to 1e-10 relative.
 * Each is plain mathematics with a closed-form CPU checker
 * (oracle/groundwork_ref.c); none cites or imitates reference code.
 *
 * Conventions: every entry point returns 0 on success or a negative B9GW_E_*
 * code; b9gw_last_error() gives the message for the calling thread.  All
 * pointers are HOST pointers unless the name ends in _dev.  There is no CPU
 * fallback: with no usable device every compute entry point fails with
 * B9GW_E_NODEVICE.
 */
#ifndef B9_GROUNDWORK_H
#define B9_GROUNDWORK_H

#ifdef __cplusplus
extern "C" {
#endif

#define B9GW_OK            0
#define B9GW_E_NODEVICE  (-1)   /* no CUDA device / driver */
#define B9GW_E_CUDA      (-2)   /* a CUDA runtime call failed */
#define B9GW_E_ARG       (-3)   /* bad argument */

#define B9GW_DFMA_ILP      8    /* independent FMA chains per thread */
#define B9GW_DFMA_THREADS  256  /* threads per CTA */
#define B9GW_TRANS_ILP     4    /* independent exp/log chains per thread */

/* ABI version of this header; bumped on any signature change. */
int b9gw_abi_version(void);

/* Message for the last failing call on this thread ("" if none). */
const char *b9gw_last_error(void);

/* Number of visible CUDA devices (0 if none); never fails. */
int b9gw_device_count(void);

/* SM count and current SM clock ceiling (MHz) of `device`. */
int b9gw_device_info(int device, int *sm_count, int *sm_clock_mhz,
                     long long *l2_bytes);

/*
 * FP64 FMA peak.  Launches sm_count*ctas_per_sm CTAs of B9GW_DFMA_THREADS
 * threads; every thread advances B9GW_DFMA_ILP chains x <- fma(x, a, b) for
 * `iters` steps and stores the in-order sum of its chains.  Does `warmup`
 * untimed launches, then `reps` timed ones bracketed by CUDA events on the
 * launching stream.
 *   out_host   : n_threads doubles (may be NULL) — thread t's result depends
 *                only on (t & 31); see oracle/groundwork_ref.c:b9ref_dfma_lane
 *   n_threads  : total threads launched
 *   ms_per_launch, tflops : average over `reps`; flops = 2*ILP*iters*n_threads
 */
int b9gw_dfma_peak(int device, int ctas_per_sm, int iters, double a, double b,
                   int warmup, int reps, double *out_host,
                   long long *n_threads, float *ms_per_launch, double *tflops);

/*
 * FP64 transcendental issue rate.  which = 0: x <- exp(-x); 1: x <- log(x + 3);
 * 2: x <- exp10(-0.4 x) (magnitude -> flux); 3: x <- log10(x + 3).  All four
 * are contractions, so the stored value is insensitive to last-bit libm
 * differences.  Same launch shape and timing as above with
 * B9GW_TRANS_ILP chains per thread; gevals = 1e-9 * ILP*iters*n_threads / s.
 */
int b9gw_transcendental_rate(int device, int which, int ctas_per_sm, int iters,
                             int warmup, int reps, double *out_host,
                             long long *n_threads, float *ms_per_launch,
                             double *gevals_per_s);

/*
 * Host round trip of one dependent step, the floor under a sequential MCMC
 * chain: wall-clock microseconds per iteration, averaged over `reps`, of
 *   us_launch_sync     : launch a 1-thread kernel, cudaStreamSynchronize
 *   us_launch_d2h_sync : launch, cudaMemcpyAsync 8 B to pinned host, synchronize
 *   us_graph_d2h_sync  : the same launch + copy replayed as one CUDA graph
 * Timed on the host (std::chrono) because the host wait is what is measured.
 */
int b9gw_step_latency(int device, int warmup, int reps, float *us_launch_sync,
                      float *us_launch_d2h_sync, float *us_graph_d2h_sync);

/* Elementwise y[i] = exp(x[i]) (which=0) or log(x[i]) (which=1) on the device,
 * for comparing CUDA libm with the host's bit by bit. */
int b9gw_map(int device, int which, const double *x_host, double *y_host,
             long long n);

/*
 * Fixed-order log-sum-exp.  x is rows x cols, row-major.  One warp per row:
 * exact row max, then each lane sums exp(x - max) over its columns
 * (col = lane, lane+32, ...) in increasing order, then an xor-butterfly
 * (offsets 16,8,4,2,1) adds the 32 partials; row_lse = max + log(sum), and
 * -inf for a row whose max is -inf.  `total` is the sum of row_lse in the fixed
 * order of b9ref_ordered_sum (1024 strided partials, then a pairwise tree).
 * Timing covers `reps` launches of both kernels with x already on the device.
 *   row_lse_host : rows doubles (may be NULL);  total_host : 1 double
 */
int b9gw_lse_rows(int device, const double *x_host, long long rows,
                  long long cols, int warmup, int reps, double *row_lse_host,
                  double *total_host, float *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* B9_GROUNDWORK_H */
