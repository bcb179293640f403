/*
 * b9_groundwork.h — C-ABI of libb9_groundwork.so  (ABI version 5)
 *
 * STATUS: the BASE-9 hot path is BLOCKED.  /root/reference is a
 * 4-line relocation notice (/root/reference/README.md:1-4); the base-cpp source
 * it points to is not staged offline.  BASELINE.json's north_star forbids
 * reconstructing the reference from memory, so NOTHING in this header is, or
 * claims to be, the drop-in boundary for BASE-9's cluster log-likelihood.  That
 * boundary can only be declared after the chain driver's call into the
 * likelihood has been read from source.
 *
 * What this header does declare is reference-independent:
 *   1. FP64 denominators for the roofline (DFMA peak; exp/log/exp10/log10
 *      rates, each labelled with the argument distribution it was taken on);
 *   2. CUDA-libm versus host libm, element by element (ULP distance);
 *   3. a fixed-order FP64 log-sum-exp over rows, fed either from a matrix in
 *      memory or from terms produced in registers by a closed-form synthetic
 *      generator (pure arithmetic, NOT a photometric model), to show what the
 *      fixed order costs when the terms never travel through memory;
 *   4. the one collective north_star names — a sum of per-chain FP64 scalars
 *      over ranks — made bit-identical for every world size by reducing over
 *      V fixed VIRTUAL SHARDS instead of over ranks, and implemented as one
 *      kernel over NVLink peer memory behind this C-ABI (no NCCL call, no
 *      Python on the step path).
 * Each has a closed-form CPU checker (oracle/groundwork_ref.c); none cites or
 * imitates reference code.
 *
 * Conventions: every entry point returns 0 on success or a negative B9GW_E_*
 * code; b9gw_last_error() gives the message for the calling thread.  Pointers
 * are HOST pointers unless the name ends in _dev.  No entry point changes the
 * caller's current CUDA device.  There is no CPU fallback: with no usable
 * device every compute entry point fails with B9GW_E_NODEVICE.
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
#define B9GW_E_TIMEOUT   (-4)   /* a peer did not arrive within the comm's timeout */
#define B9GW_E_STATE     (-5)   /* call made in the wrong state (e.g. not connected) */

#define B9GW_DFMA_ILP      8    /* independent FMA chains per thread at the peak (ilp = 8) */
#define B9GW_DFMA_THREADS  256  /* threads per CTA */
#define B9GW_TRANS_ILP     4    /* independent exp/log chains per thread */

#define B9GW_LSE_STAGED_COLS 1024 /* rows up to this length are staged on chip */
#define B9GW_LSE_MAX_CHAINS 65535 /* chains one b9gw_lse_generated_shards launch takes */

#define B9GW_MAX_WORLD        16  /* ranks in one comm (one NVSwitch domain) */
#define B9GW_MAX_VSHARDS      128 /* virtual shards; must be one of 4,8,16,32,64,128 */
#define B9GW_IPC_HANDLE_BYTES 64  /* opaque per-rank handle exchanged at connect */

/* ABI version of this header; bumped on any signature change. */
int b9gw_abi_version(void);

/* Message for the last failing call on this thread ("" if none). */
const char *b9gw_last_error(void);

/* Number of visible CUDA devices (0 if none); never fails. */
int b9gw_device_count(void);

/* SM count and current SM clock ceiling (MHz) of `device`. */
int b9gw_device_info(int device, int *sm_count, int *sm_clock_mhz,
                     long long *l2_bytes);

/* ------------------------------------------------------------------ rates */

/*
 * FP64 FMA peak and dependent-issue latency.  Launches sm_count*ctas_per_sm CTAs
 * of B9GW_DFMA_THREADS threads; every thread advances `ilp` (1, 2, 4 or 8)
 * independent chains x <- fma(x, a, b) for `iters` steps and stores the in-order
 * sum of its chains.  ilp = 8 measures the peak.  ilp = 1 with ctas_per_sm = 1
 * leaves two chains in flight per scheduler, so the time per step is half the
 * dependent-issue latency: the number a design needs to know how many
 * independent FP64 chains it must keep in flight.  int_per_fma (ilp 8 only)
 * gives every DFMA independent integer instructions for company: 1 or 2 = that
 * many multiply-adds (IMAD: half-rate, on another pipe), -2 or -4 = that many
 * add/xor instructions (full-rate ALU).  How far the DFMA rate drops says what
 * a non-FP64 instruction costs next to FP64 work on this part.  Does `warmup` untimed launches, then `reps` timed ones bracketed
 * by CUDA events on the launching stream.
 *   out_host   : n_threads doubles (may be NULL) — thread t's result depends
 *                only on (t & 31); see oracle/groundwork_ref.c:b9ref_dfma_lane
 *   n_threads  : total threads launched
 *   ms_per_launch, tflops : average over `reps`; flops = 2*ilp*iters*n_threads
 */
int b9gw_dfma_peak(int device, int ctas_per_sm, int ilp, int int_per_fma, int iters,
                   double a, double b, int warmup, int reps, double *out_host,
                   long long *n_threads, float *ms_per_launch, double *tflops);

/*
 * FP64 transcendental issue rate.
 *   which 0..3 — CONTRACTIONS: x <- exp(-x); x <- log(x + 3); x <- exp10(-0.4 x);
 *     x <- log10(x + 3).  After a few dozen iterations every thread evaluates
 *     the function at ONE mid-range argument (exp near 0.567, log near 3.5), so
 *     these are single-argument, mid-range rates; the stored value is
 *     insensitive to last-bit libm differences.
 *   which 4 — exp over a SPREAD of arguments: every evaluation takes a fresh
 *     argument -(2^e * 1.f) with e uniform in [-6, 9] and f 20 random bits from
 *     a per-chain 32-bit LCG (integer pipe), i.e. log-uniform magnitudes from
 *     1/64 to 1024, ~3 % of them below exp's underflow threshold — the range a
 *     max-shifted log-sum-exp feeds to exp.  The thread stores the sum of its
 *     evaluations (one extra DADD per exp).
 *   which 5 — log over a spread: arguments 2^e * 1.f, e uniform in [-8, 7].
 * Same launch shape and timing as above with B9GW_TRANS_ILP chains per thread;
 * gevals = 1e-9 * ILP*iters*n_threads / s.
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

/* Elementwise y[i] = f(x[i]) on the device, f = exp (which 0), log (1),
 * exp10 (2), log10 (3), for comparing CUDA libm with the host's bit by bit;
 * which 4 is the LSE kernels' branch-free copy of exp's fast path, defined only
 * for -708 < x <= 0 and required to equal exp there bit for bit. */
int b9gw_map(int device, int which, const double *x_host, double *y_host,
             long long n);

/* ------------------------------------------------- fixed-order log-sum-exp */

/*
 * Fixed-order log-sum-exp of every row, and the rows' sum over virtual shards.
 * x is rows x cols, row-major.
 *   row value : exact row max m; lane l (of 32) adds exp(x[c] - m) over its
 *               columns c = l, l+32, ... in increasing order starting from +0;
 *               an xor-butterfly (offsets 16,8,4,2,1) adds the 32 lane sums;
 *               value = m + log(sum), and -inf for a row whose max is -inf.
 *   P[v]      : the rows are cut into n_vshards virtual shards exactly as stars
 *               are below ("world-size-independent sum"); P[v] is the warp-order
 *               sum of shard v's row values (+0 for an empty shard).
 *   total     : (((0 + P[0]) + P[1]) + ...) + P[V-1].
 * The order fixes which lane ADDS which term, not which thread evaluates exp:
 * rows of up to B9GW_LSE_STAGED_COLS columns are read from memory ONCE, parked
 * in shared memory, exponentiated in place by two warps and then added in the
 * pinned order by one; longer rows are read twice by one warp.  Every bit of the
 * result is the same on both paths.  P[] and total are produced in the same
 * launch, by whichever warp finishes a shard / the last shard.
 * Timing covers `reps` launches with x already on the device.
 *   row_lse_host : rows doubles (may be NULL);  partials_host : n_vshards
 *   doubles (may be NULL);  total_host : 1 double
 */
int b9gw_lse_rows(int device, const double *x_host, long long rows,
                  long long cols, int n_vshards, int warmup, int reps,
                  double *row_lse_host, double *partials_host,
                  double *total_host, float *ms_per_launch);

/*
 * The synthetic term generator, materialised: x[r*cols + c] = g(r, c, cols),
 *   u  = r * 0.6180339887498949,  c0 = (u - floor(u)) * cols,
 *   w  = (34 + (r mod 7)) * (1 / cols),
 *   t  = fma(c, w, -(c0 * w)),    g = -(t * t)
 * (each operation rounded once, as written; oracle/groundwork_ref.c:b9ref_gen_term
 * is the host statement).  A row is a parabola with its maximum (0 >= g > -1)
 * near column c0 and a minimum between about -290 and -1600: some rows reach
 * past exp's underflow threshold, all rows have a few hundred terms near the
 * maximum that carry the sum.  It is arithmetic chosen to give exp a realistic
 * spread of arguments; it models nothing.  cols <= 2^40.
 */
int b9gw_generate_terms(int device, long long rows, long long cols, double *x_host);

/*
 * The same computation, but every term is produced on chip by g(r, c, cols)
 * above: no matrix exists in memory, one exp per term.  Row values, P[] and
 * total are bit-identical to b9gw_lse_rows run on b9gw_generate_terms' output.
 */
int b9gw_lse_generated(int device, long long rows, long long cols, int n_vshards,
                       int warmup, int reps, double *row_lse_host,
                       double *partials_host, double *total_host,
                       float *ms_per_launch);

/*
 * One rank's share of a star-sharded job, on the device: the stars of virtual
 * shards [first_shard, first_shard + n_shards) of an n_stars_total-star job (cut
 * as under "world-size-independent sum" below), for `chains` independent chains
 * in ONE launch (grid.y = chains).  The terms of (chain, star) are row
 * chain * n_stars_total + star of the generator above, so the work a rank does
 * depends on which stars it owns and on nothing else.  Writes
 *   row_lse_dev  [chains][n_local]  n_local = stars in the local shards
 *   partial_dev  [n_shards][chains] P[v][chain] — the input of b9gw_ordered_allreduce
 *   total_dev    [chains]           only when every shard is local (n_shards ==
 *                                   n_vshards); may be NULL, is untouched otherwise
 * workspace_dev is b9gw_lse_workspace_bytes(chains, n_shards) bytes, zero before
 * the first launch (b9gw_dev_malloc zero-fills); every launch leaves it zero
 * again, and launches that share a workspace must be ordered on one stream.
 * Launched on `cuda_stream` (NULL = the legacy default stream), not
 * synchronised, capturable.  chains <= B9GW_LSE_MAX_CHAINS and
 * chains * n_stars_total < 2^31.
 */
long long b9gw_lse_workspace_bytes(long long chains, int n_shards);
int b9gw_lse_generated_shards(int device, long long n_stars_total, long long cols,
                              long long chains, int n_vshards, int first_shard,
                              int n_shards, double *row_lse_dev, double *partial_dev,
                              double *total_dev, void *workspace_dev, void *cuda_stream);

/* ------------------------------------------------------- device memory */

/*
 * Plain device-memory helpers, so that a host program (the C++ chain driver)
 * can hold the *_dev buffers of the calls below without linking the CUDA
 * runtime itself.  Copies are synchronous and ordered after work already
 * launched on the device's legacy default stream (cuda_stream = NULL).
 * b9gw_dev_malloc returns zero-filled memory.
 */
int b9gw_dev_malloc(int device, long long bytes, void **ptr_dev);
int b9gw_dev_free(int device, void *ptr_dev);
int b9gw_memcpy_h2d(int device, void *dst_dev, const void *src_host, long long bytes);
int b9gw_memcpy_d2h(int device, void *dst_host, const void *src_dev, long long bytes);

/* -------------------------- world-size-independent sum over virtual shards */

/*
 * A job's N stars are cut into V VIRTUAL SHARDS, a function of (N, V) only:
 * shard v holds stars [floor(v*N/V), floor((v+1)*N/V)).  A world of W ranks
 * (V % W == 0) gives rank r the V/W consecutive shards starting at r*V/W.
 * A per-chain sum over stars is then defined as
 *     total[c] = (((0 + P[0][c]) + P[1][c]) + ... ) + P[V-1][c],
 *     P[v][c]  = warp-order sum of the shard's per-star values
 *                (32 lane-strided serial partials, xor-butterfly 16,8,4,2,1),
 * which does not mention W: the same N, V and per-star values give the same
 * bits at W = 1, 2, 4, 8.  (Summing per-rank partials in rank order does not
 * have this property — the grouping changes with W.)
 */

/* Stars [lo, hi) of virtual shard `shard`.  Pure host arithmetic.  n_stars <= 2^48. */
int b9gw_vshard_bounds(long long n_stars, int n_vshards, int shard,
                       long long *lo, long long *hi);

/*
 * P for the local shards.  values_dev is [chains][ld] row-major and holds this
 * rank's stars only: column 0 is star lo(first_shard).  partial_dev receives
 * [n_shards][chains].  One warp per (shard, chain); launched on `cuda_stream`
 * (a cudaStream_t of `device`; NULL = its legacy default stream) and not
 * synchronised.
 */
int b9gw_shard_partials(int device, const double *values_dev, long long chains, long long ld,
                        long long n_stars_total, int n_vshards, int first_shard,
                        int n_shards, double *partial_dev, void *cuda_stream);

typedef struct b9gw_comm b9gw_comm;

/*
 * Phase 1, on every rank: allocate this rank's mailbox on `device` (2 parities
 * x V x max_chains 16-byte slots, zeroed) and write an opaque handle to
 * handle_out[B9GW_IPC_HANDLE_BYTES].  The caller publishes the handles by any
 * means it has (MPI, a file, torch.distributed.all_gather ...).
 * n_vshards must be a power of two in [4, B9GW_MAX_VSHARDS] and a multiple of world.
 */
int b9gw_comm_create(int device, int rank, int world, int n_vshards,
                     long long max_chains, b9gw_comm **comm, void *handle_out);

/*
 * Phase 2: all_handles is world x B9GW_IPC_HANDLE_BYTES in rank order.  Maps
 * every peer's mailbox into this process (CUDA IPC; peers must be P2P-capable
 * GPUs of one node).  With world == 1 this does nothing and may be skipped.
 */
int b9gw_comm_connect(b9gw_comm *comm, const void *all_handles);

/*
 * One step.  partial_dev is this rank's [V/world][chains] (as written by
 * b9gw_shard_partials); out_dev receives total[chains] on EVERY rank.  A single
 * kernel, launched on `cuda_stream` and not synchronised: each thread stores
 * its partials straight into every rank's mailbox over NVLink as 16-byte
 * {lo32, step, hi32, step} packets (8-byte-atomic, so a packet carries its own
 * arrival flag: no fence, no separate flag, no barrier), then polls its own
 * mailbox slots and adds shards 0..V-1 left to right.  The step counter lives
 * in device memory, so the launch can be captured in a CUDA graph and replayed.
 * All calls on one comm must be ordered on one stream, every rank must make
 * the same sequence of calls, and each rank must own its GPU (ranks that wait
 * on one another cannot share a device).  If a peer does not arrive within the
 * comm's timeout the kernel stores NaN, raises the comm's sticky status and
 * returns; it never spins forever, and once the status is raised every later
 * step on that comm fails at once (NaN) instead of waiting again.
 */
int b9gw_ordered_allreduce(b9gw_comm *comm, const double *partial_dev,
                           double *out_dev, long long chains, void *cuda_stream);

/* Spin budget of one step in milliseconds (default 2000). */
int b9gw_comm_set_timeout_ms(b9gw_comm *comm, int ms);

/*
 * Synchronises the device and reports the sticky status: *timed_out != 0 if any
 * step gave up waiting; *steps = steps completed.  Returns B9GW_E_TIMEOUT in
 * that case so a caller that ignores the out-parameters still sees it.
 */
int b9gw_comm_status(b9gw_comm *comm, int *timed_out, unsigned long long *steps);

/*
 * Latency of the step kernel alone, measured with CUDA events on a private
 * stream with `reps` back-to-back launches after `warmup`: us_stream for plain
 * launches, us_graph for a CUDA graph of 8 steps per launch (reps rounded up to
 * a multiple of 8).  Every rank must call it with the same arguments.
 */
int b9gw_allreduce_latency(b9gw_comm *comm, long long chains, int warmup,
                           int reps, float *us_stream, float *us_graph);

/*
 * The star-sharded step as ONE kernel: b9gw_lse_generated_shards over this
 * rank's V/world shards of the job, with the cross-rank sum fused into its tail.
 * Only the warp that completes a chain's last local shard talks to the peers:
 * it stores the chain's local P[] straight into every OTHER rank's mailbox (the
 * same 16-byte self-flagging packets as above), polls its own mailbox for the
 * remote shards and adds all V left to right into total_dev[chain] — while
 * other chains' rows are still being evaluated on the rest of the GPU.  (A
 * polling warp keeps its CTA resident; progress relies on the hardware issuing
 * a grid's CTAs in block-index order on every rank, as every stream-K style
 * kernel does.)  A rank whose shards hold no star at all still takes part (it
 * only pulls).  Same
 * buffers, limits and workspace as b9gw_lse_generated_shards (the local shard
 * range is the comm's); same ordering rules, step counters, timeout behaviour
 * and bits as b9gw_ordered_allreduce, with which it may be freely mixed on one
 * comm.  total_dev [chains] is written on every rank.
 */
int b9gw_lse_generated_step(b9gw_comm *comm, long long n_stars_total, long long cols,
                            long long chains, double *row_lse_dev, double *partial_dev,
                            double *total_dev, void *workspace_dev, void *cuda_stream);

/*
 * A star-sharded step, end to end and timed, two ways on one private stream
 * (CUDA events, `reps` after `warmup`, no host work inside a step):
 *   us_lse_alone  : b9gw_lse_generated_shards over this rank's V/world shards of
 *                   an (n_stars_total x cols)-term, `chains`-chain job;
 *   us_step       : that launch, then b9gw_ordered_allreduce of the partials it
 *                   wrote — two launches per step;
 *   us_fused_step : b9gw_lse_generated_step — one launch per step.
 * total_host / total_fused_host (may be NULL) receive total[chains] of the last
 * two-launch / fused step: the same bits as each other, on every rank AND at
 * every world size, because no W enters the definition of any kernel's result.
 * Every rank must call with the same arguments.
 */
int b9gw_sharded_step(b9gw_comm *comm, long long n_stars_total, long long cols,
                      long long chains, int warmup, int reps, double *total_host,
                      double *total_fused_host, float *us_step, float *us_lse_alone,
                      float *us_fused_step);

/* Unmaps the peers and frees the mailbox.  The caller must make sure (barrier)
 * that no peer is still inside a step. */
int b9gw_comm_destroy(b9gw_comm *comm);

/*
 * Host-buffer convenience, world = 1: values_host is [chains][n_stars];
 * partials_host receives P as [V][chains] (may be NULL), total_host receives
 * total[chains].  Runs b9gw_shard_partials + b9gw_ordered_allreduce on one GPU.
 */
int b9gw_vshard_total(int device, const double *values_host, long long chains,
                      long long n_stars, int n_vshards, double *partials_host,
                      double *total_host);

#ifdef __cplusplus
}
#endif
#endif /* B9_GROUNDWORK_H */
