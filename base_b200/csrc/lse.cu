// lse.cu — fixed-order FP64 log-sum-exp over rows, and their sum over virtual
// shards, for sm_100a.
//
// Reference-independent groundwork (the BASE-9 hot path is BLOCKED;
// nothing here restates or imitates reference code).  The ORDER is the
// contract, pinned bit for bit by oracle/groundwork_ref.c:
//   max     : exact, any order;
//   sum     : lane l adds exp(x[c] - max) for c = l, l+32, l+64, ... in
//             increasing c, starting from +0; the 32 lane sums are combined by
//             an xor-butterfly with offsets 16, 8, 4, 2, 1;
//   value   : max + log(sum), or -inf when max == -inf;
//   P[v]    : warp-order sum (same lane-strided + butterfly shape) of the row
//             values of virtual shard v = rows [floor(v*R/V), floor((v+1)*R/V));
//   total   : (((0 + P[0]) + P[1]) + ...) + P[V-1]  — the same definition the
//             cross-rank sum uses (vshard.cu), so one kernel's P[] can feed it.
//
// The order says which lane ADDS which term; it does not say which thread
// evaluates exp.  Staged rows (<= B9GW_LSE_STAGED_COLS columns) use that:
//   pass 1  : the row's two warps fetch/generate every term once and park it
//             in the row's slice of shared memory, tracking the max;
//   pass 2a : the same 64 threads overwrite each parked term with
//             exp(term - max) — independent work, any thread may do any term;
//   pass 2b : one warp adds the parked exponentials in the pinned order.
// Two warps per row double the exp chains in flight per staged row (shared
// memory, 8 KB a row, is what limits rows in flight), which is what an
// exp-latency-bound loop needs; profiles/r02_groundwork.md has the numbers for
// the one-warp-per-row variants this replaced.
// Two term sources share the code: a matrix in memory (lse_rows) and a
// closed-form generator evaluated on chip (lse_generated).
//
// Every value-path operation is an explicit __dadd_rn/__dsub_rn/__dmul_rn/fma,
// so -fmad cannot contract anything the host checker does not.

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

using b9gw::fail;
using b9gw::ExpConstants;
using b9gw::exp_constants;
using b9gw::exp_fast_path;
using b9gw::PeerArgs;

constexpr unsigned FULL = 0xffffffffu;
constexpr int CTA_THREADS = 256;
#ifndef B9GW_STAGED_WARPS
#define B9GW_STAGED_WARPS 4
#endif
#ifndef B9GW_PASS2_TERMS
#define B9GW_PASS2_TERMS 4
#endif
// Without a minimum-CTAs launch bound ptxas schedules for full occupancy (32 registers a
// thread) and un-interleaves the four exps of pass 2 into 1 + 1 + 2 dependent chains; shared
// memory allows 6 CTAs per SM anyway.  With the bound it keeps all four chains interleaved
// (62 registers): 304 instead of 330 us at 160 000 x 1024 (gpurun A/B, profiles/r02b_*).
// The fused (PEER) instantiation carries the mailbox tail as well and spills at 64 registers;
// it is bounded by the 6 CTAs shared memory allows at 1024 columns.
#ifndef B9GW_STAGED_MIN_CTAS
#define B9GW_STAGED_MIN_CTAS 8
#endif
#ifndef B9GW_STAGED_MIN_CTAS_PEER
#define B9GW_STAGED_MIN_CTAS_PEER 6
#endif
constexpr int STAGED_WARPS = B9GW_STAGED_WARPS;   // staged rows in flight per CTA, one warp each
constexpr int PASS2_TERMS = B9GW_PASS2_TERMS;     // terms a lane exponentiates per iteration (divides 128 / 32 * k)
constexpr int STREAM_ROWS = 8;            // rows per CTA, one warp each

// ------------------------------------------------------------- term sources
struct MatrixRow {                        // terms live in memory
    const double *p;
    __device__ double operator()(long long c) const { return __ldg(p + c); }
};

struct GeneratedRow {                     // terms are arithmetic on (row, col)
    double w, b;                          // t = fma(c, w, b), term = -(t*t)
    __device__ GeneratedRow(long long row, long long cols, double inv_cols) {
        const unsigned r32 = (unsigned)row;                       // rows < 2^31 (checked by the host)
        const double u = __dmul_rn((double)r32, 0.6180339887498949);
        const double c0 = __dmul_rn(__dsub_rn(u, floor(u)), (double)cols);
        w = __dmul_rn((double)(34u + r32 % 7u), inv_cols);        // inv_cols = 1.0 / cols, rounded once
        b = -__dmul_rn(c0, w);
    }
    __device__ double operator()(long long c) const {
        // (double)c for 0 <= c < 2^52 as one exact DADD instead of a 64-bit I2F conversion
        const double cd = __dsub_rn(__longlong_as_double(0x4330000000000000LL | c), 4503599627370496.0);
        const double t = fma(cd, w, b);
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

__host__ __device__ __forceinline__ long long shard_lo(long long n, int shift, long long v) {
    return (long long)(((unsigned long long)v * (unsigned long long)n) >> shift);
}

// Lane-strided serial partials over v[lo..hi), 4 loads in flight, then the butterfly.
__device__ __forceinline__ double warp_ordered_sum(const double *v, long long lo, long long hi, int lane) {
    double s = 0.0;
    for (long long i0 = lo + lane; i0 < hi; i0 += 32 * 4) {
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = i0 + 32 * u < hi ? __ldcg(v + i0 + 32 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < hi) s = __dadd_rn(s, x[u]);
    }
    return warp_add(s);
}

__device__ __forceinline__ unsigned ticket_add(unsigned *p, unsigned v) {
    unsigned old;                          // release our row values, acquire everyone else's
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// One launch works on `chains` independent chains (blockIdx.y) over the stars of the local
// virtual shards [first_shard, first_shard + n_shards) of an n_total-star job; local star s is
// global star star0 + s.  With all V shards local the launch also produces total[chain].
struct LseArgs {
    const double *x;                       // SRC 0: [chains][n_local][cols]
    double *row_lse;                       // [chains][n_local]
    double *partials;                      // [n_shards][chains]
    double *total;                         // [chains] or null; written only when n_shards == V
    unsigned *tickets;                     // [chains][n_shards + 1], zero between launches
    long long n_local, n_total, star0, chains, cols_ll;
    double inv_cols;
    int cols, cap, rpw, vshift, first_shard, n_shards;
#ifdef B9GW_DIAG
    int diag;                              // tools/fused_diag.py: 1 = do not wait for remote shards, 2 = do not push to peers
#endif
};

#ifdef B9GW_DIAG
__device__ unsigned long long g_diag[8 * 16 * 4];   // [step % 8][chain < 16]: enter pull, lane 0 has its packets, total stored, chain's first CTA start
#define DIAG_SLOT(step, chain) (((step) & 7u) * 64 + (chain) * 4)
#endif

// Shards 32q + lane, q = 0.., held one per lane in pk[q] (empty shards: +0), added strictly left
// to right starting from +0.  The shuffles pipeline; only the adds are serial.
__device__ __forceinline__ double add_shards_in_order(const double (&pk)[B9GW_MAX_VSHARDS / 32], int V) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < B9GW_MAX_VSHARDS / 32; ++q) {
        if (32 * q >= V) break;                            // warp-uniform
#pragma unroll 8
        for (int u = 0; u < 32; ++u) {
            const double pv = __shfl_sync(FULL, pk[q], u);
            if (32 * q + u < V) acc = __dadd_rn(acc, pv);
        }
    }
    return acc;
}

// PEER: the cross-rank half of a fused step, run by the one warp per chain that took the
// chain's last local ticket (so every local P[k] is in memory and acquired).  Lane l owns
// shards l, l+32, ...: a local one it reads from partials[] and stores into every OTHER rank's
// mailbox as a 16-byte self-flagging packet (vshard.cu); a remote one it polls for in this
// rank's own mailbox; an empty one is known to be +0 and neither sent nor awaited.  Then the V
// values are added left to right, total[chain] is stored and the chain's step counter moves.
// The counter cannot move before this launch's pushes have been issued (same warp), nor can
// a peer run ahead by more than a step (it needs our packets), so two mailbox parities are
// enough — as in vshard.cu.  A peer that never arrives costs the comm's timeout: NaN and the
// sticky status.
__device__ void pull_and_total(const LseArgs &a, const PeerArgs &pa, long long chain, unsigned *tk, int lane) {
    const int V = 1 << a.vshift;
    const unsigned step = *(volatile unsigned *)(pa.seq + chain) + 1u;
    const size_t parity_base = (size_t)(step & 1u) * V;
    double pk[B9GW_MAX_VSHARDS / 32];
    unsigned need = 0;
#pragma unroll
    for (int q = 0; q < B9GW_MAX_VSHARDS / 32; ++q) {
        const int u = lane + 32 * q;
        pk[q] = 0.0;
        if (u >= V || shard_lo(a.n_total, a.vshift, u + 1) <= shard_lo(a.n_total, a.vshift, u)) continue;
        if (u >= a.first_shard && u < a.first_shard + a.n_shards) {
            pk[q] = __ldcg(a.partials + (long long)(u - a.first_shard) * a.chains + chain);
            const unsigned lo = (unsigned)__double2loint(pk[q]), hi = (unsigned)__double2hiint(pk[q]);
            const size_t slot = (parity_base + (size_t)u) * (size_t)pa.max_chains + (size_t)chain;
#ifdef B9GW_DIAG
            if (!(a.diag & 2))
#endif
            for (int d = 1; d < pa.world; ++d) {
                int peer = pa.rank + d;
                if (peer >= pa.world) peer -= pa.world;
                b9gw::st_packet(pa.mail[peer] + slot, lo, hi, step);
            }
        } else {
            need |= 1u << q;
#ifdef B9GW_DIAG
            if (a.diag & 1) need &= ~(1u << q);
#endif
        }
    }
#ifdef B9GW_DIAG
    if (lane == 0 && chain < 16) g_diag[DIAG_SLOT(step, chain)] = b9gw::globaltimer_ns();
#endif
    // The warp stays converged: it leaves the loop as a whole, on a vote.  (Lanes leaving one by
    // one made everything after the loop — 128 shuffles — run ~12x slower on the rank that had
    // to wait: profiles/r02b_groundwork.md.)
    const uint4 *mine = pa.mail[pa.rank] + parity_base * (size_t)pa.max_chains + (size_t)chain;
    bool ok = true;
    unsigned long long t0 = *(volatile int *)pa.status ? ~0ULL : 0;   // a comm that timed out stays out of step
    for (;;) {
        uint4 r[B9GW_MAX_VSHARDS / 32];
#pragma unroll
        for (int q = 0; q < B9GW_MAX_VSHARDS / 32; ++q)
            if (need >> q & 1u) r[q] = b9gw::ld_packet(mine + (size_t)(lane + 32 * q) * (size_t)pa.max_chains);
#pragma unroll
        for (int q = 0; q < B9GW_MAX_VSHARDS / 32; ++q)
            if ((need >> q & 1u) && r[q].y == step && r[q].w == step) {
                pk[q] = __hiloint2double((int)r[q].z, (int)r[q].x);
                need &= ~(1u << q);
            }
        if (__all_sync(FULL, need == 0)) break;
        bool expired = t0 == ~0ULL;
        if (!expired) {
            const unsigned long long now = b9gw::globaltimer_ns();
            if (t0 == 0) t0 = now;
            else expired = now - t0 > pa.timeout_ns;
        }
        if (__any_sync(FULL, expired)) { ok = need == 0; break; }
    }
    const bool bad = __any_sync(FULL, !ok);
#ifdef B9GW_DIAG
    if (lane == 0 && chain < 16) g_diag[DIAG_SLOT(step, chain) + 1] = b9gw::globaltimer_ns();
#endif
    const double acc = add_shards_in_order(pk, V);
    if (lane == 0) {
#ifdef B9GW_DIAG
        if (chain < 16) g_diag[DIAG_SLOT(step, chain) + 2] = b9gw::globaltimer_ns();
#endif
        a.total[chain] = bad ? __longlong_as_double(0x7ff8000000000000LL) : acc;
        if (bad) atomicExch(pa.status, 1);
        pa.seq[chain] = step;
        tk[a.n_shards] = 0;
    }
}

// Called by one whole warp once the row values of local stars [s0, s1) of `chain` are
// stored — all of them by this warp's lane 0, which also takes the tickets, so a ticket
// releases them.  The warp that completes a virtual shard adds that shard's row values; the
// warp that completes the chain's last local shard finishes the chain: with all shards local
// (and no peers) it adds the shards left to right; PEER, it runs pull_and_total.  Which warp
// that is does not change any bit.  tickets[chain][k] counts finished rows of local shard k,
// tickets[chain][n_shards] finished shards; each is reset by its last user, ready for the
// next launch on the stream.  (PEER: pushing each P from the warp that completed it — three
// more round trips on 64 warps' tails per chain — cost more than it gained; measured.)
template <bool PEER>
__device__ void finish_rows(const LseArgs &a, const PeerArgs &pa, long long chain, long long s0,
                            long long s1, int lane) {
    const int V = 1 << a.vshift;
    const long long g0 = a.star0 + s0, g1 = a.star0 + s1;
    unsigned *tk = a.tickets + chain * (a.n_shards + 1);
    const double *rows = a.row_lse + chain * a.n_local;
    // local shards that hold a star: with fewer stars than shards every shard holds 0 or 1
    const unsigned nonempty = a.n_total < V ? (unsigned)a.n_local : (unsigned)a.n_shards;
    const bool totals = PEER || (a.n_shards == V && a.total);
    bool all_done = false;
    for (long long v = (((g0 + 1) << a.vshift) + a.n_total - 1) / a.n_total - 1;    // shard holding g0
         v < a.first_shard + a.n_shards; ++v) {
        const long long lo = shard_lo(a.n_total, a.vshift, v), hi = shard_lo(a.n_total, a.vshift, v + 1);
        if (lo >= g1) break;
        if (hi <= lo) continue;                            // empty shard (n_total < V)
        const unsigned mine = (unsigned)((hi < g1 ? hi : g1) - (lo > g0 ? lo : g0));
        const int k = (int)(v - a.first_shard);
        unsigned done = 0;
        if (lane == 0) done = ticket_add(&tk[k], mine) + mine == (unsigned)(hi - lo);
        if (!__shfl_sync(FULL, done, 0)) continue;
        // lane 0's acquire ordered the other warps' row values before the shuffle above; the
        // loads below bypass L1 (ld.cg), so no further fence is needed (each fence here is a
        // round trip on the kernel's serial tail, ~1.5 us apiece measured)
        const double p = warp_ordered_sum(rows, lo - a.star0, hi - a.star0, lane);
        done = 0;
        if (lane == 0) {
            a.partials[(long long)k * a.chains + chain] = p;
            tk[k] = 0;
            if (totals) done = ticket_add(&tk[a.n_shards], 1u) + 1u == nonempty;
        }
        all_done |= __shfl_sync(FULL, done, 0) != 0;
    }
    if (!all_done) return;
    if constexpr (PEER) {
        pull_and_total(a, pa, chain, tk, lane);
    } else {
        double pk[B9GW_MAX_VSHARDS / 32];
#pragma unroll
        for (int q = 0; q < B9GW_MAX_VSHARDS / 32; ++q) {
            const int u = lane + 32 * q;
            pk[q] = u < V ? __ldcg(a.partials + (long long)u * a.chains + chain) : 0.0;   // empty shards hold +0
        }
        const double acc = add_shards_in_order(pk, V);
        if (lane == 0) {
            a.total[chain] = acc;
            tk[V] = 0;
        }
    }
}

// Empty local shards (fewer stars than shards) complete nothing: their P is +0, written here
// by the first warp of each chain's first CTA before any ticket can be taken.
__device__ __forceinline__ void zero_empty_shards(const LseArgs &a, long long chain, int lane) {
    for (int k = lane; k < a.n_shards; k += 32) {
        const long long v = a.first_shard + k;
        if (shard_lo(a.n_total, a.vshift, v + 1) <= shard_lo(a.n_total, a.vshift, v))
            a.partials[(long long)k * a.chains + chain] = 0.0;
    }
    __syncwarp();                          // ordered before lane 0's first ticket, which releases them
}

// Exact max without fmax's NaN fix-up: m is never NaN, and a NaN term compares false
// either way, so this returns what fmax(m, v) would.
__device__ __forceinline__ double max_keep(double m, double v) { return v > m ? v : m; }

// Staged rows: one warp per row, STAGED_WARPS rows per CTA, `cap` doubles of shared memory
// per row (cols rounded up to 128).  Pass 1 fetches/generates every term once and parks it;
// pass 2 takes four parked terms per lane at a time: when all 128 of the warp's arguments
// are inside libm's fast-path range the four exps run interleaved and branch-free,
// otherwise the warp calls exp() itself.  Trip counts are warp-uniform; the padding columns
// hold -inf, whose exp is +0 and changes no bit of the sum.
template <int SRC, bool PEER>
__global__ void __launch_bounds__(STAGED_WARPS * 32, PEER ? B9GW_STAGED_MIN_CTAS_PEER : B9GW_STAGED_MIN_CTAS)
lse_staged_kernel(const __grid_constant__ ExpConstants K, const __grid_constant__ LseArgs a,
                  const __grid_constant__ PeerArgs pa) {
    extern __shared__ double sm[];        // [STAGED_WARPS][cap]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.cols, cap = a.cap;
    double *buf = sm + warp * cap + lane;
    const int iters = cap >> 5;           // multiple of 4
    const long long chain = blockIdx.y;
    const long long s0 = ((long long)blockIdx.x * STAGED_WARPS + warp) * a.rpw;
    const long long s1 = s0 + a.rpw < a.n_local ? s0 + a.rpw : a.n_local;
#ifdef B9GW_DIAG
    if (PEER && blockIdx.x == 0 && threadIdx.x == 0 && chain < 16)
        g_diag[DIAG_SLOT(*(volatile unsigned *)(pa.seq + chain) + 1u, chain) + 3] = b9gw::globaltimer_ns();
#endif
    if (blockIdx.x == 0 && warp == 0 && a.n_total < (1LL << a.vshift)) {
        zero_empty_shards(a, chain, lane);
        if (PEER && a.n_local == 0)        // nothing local to finish: this warp does the cross-rank half
            pull_and_total(a, pa, chain, a.tickets + chain * (a.n_shards + 1), lane);
    }

    for (long long s = s0; s < s1; ++s) {
        // pass 1: fetch/generate once, park, max
        double m = -INFINITY;
        auto fetch = [&](auto src) {
            if (n == cap) {               // no padding: no predicate (warp-uniform choice)
#pragma unroll (SRC == 0 ? 8 : 4)
                for (int k = 0; k < iters; ++k) {
                    const double v = src(lane + 32 * k);
                    buf[32 * k] = v;
                    m = max_keep(m, v);
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < iters; ++k) {
                    const int c = lane + 32 * k;
                    const double v = c < n ? src(c) : -INFINITY;
                    buf[32 * k] = v;
                    m = max_keep(m, v);
                }
            }
        };
        if constexpr (SRC == 0) fetch(MatrixRow{a.x + (chain * a.n_local + s) * n});
        else fetch(GeneratedRow(chain * a.n_total + a.star0 + s, n, a.inv_cols));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max_keep(m, __shfl_xor_sync(FULL, m, o));

        double r = -INFINITY;             // every term is exp(-inf) = 0 (also n == 0)
        if (m != -INFINITY) {             // warp-uniform
            // pass 2: lane l adds exp(term - m) over c = l, l+32, ... in increasing c
            double acc = 0.0;
            for (int k = 0; k < iters; k += PASS2_TERMS) {
                double neg[PASS2_TERMS], e[PASS2_TERMS];
                bool fast = true;
#pragma unroll
                for (int i = 0; i < PASS2_TERMS; ++i) {
                    // -708 < term - m <= 0 tested on the integer pipe: m - term is the exact
                    // negation of term - m and never negative, so its high word grows with its
                    // magnitude; +inf and NaN of either sign land above the bound
                    neg[i] = __dsub_rn(m, buf[32 * (k + i)]);
                    fast &= (unsigned)__double2hiint(neg[i]) < 0x40862000u;
                }
                if (__all_sync(FULL, fast)) {
                    exp_fast_path(K, neg, e);
                } else {
#pragma unroll
                    for (int i = 0; i < PASS2_TERMS; ++i) e[i] = exp(-neg[i]);
                }
#pragma unroll
                for (int i = 0; i < PASS2_TERMS; ++i) acc = __dadd_rn(acc, e[i]);
            }
            r = __dadd_rn(m, log(warp_add(acc)));
        }
        if (lane == 0) a.row_lse[chain * a.n_local + s] = r;
        __syncwarp();                     // the row's slice is reused by the next row
    }
    if (s0 < s1) finish_rows<PEER>(a, pa, chain, s0, s1, lane);
}

// Longer rows: one warp per row, two passes over the source.
template <int SRC, bool PEER>
__global__ void __launch_bounds__(CTA_THREADS, 4)
lse_stream_kernel(const __grid_constant__ LseArgs a, const __grid_constant__ PeerArgs pa) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long chain = blockIdx.y, cols = a.cols_ll;
    const long long s = (long long)blockIdx.x * STREAM_ROWS + warp;
    if (blockIdx.x == 0 && warp == 0 && a.n_total < (1LL << a.vshift)) {
        zero_empty_shards(a, chain, lane);
        if (PEER && a.n_local == 0)
            pull_and_total(a, pa, chain, a.tickets + chain * (a.n_shards + 1), lane);
    }
    if (s >= a.n_local) return;
    auto lse = [&](auto src) -> double {
        constexpr int B = 8;
        double m = -INFINITY;
        for (long long c0 = lane; c0 < cols; c0 += 32 * B) {
            double v[B];
#pragma unroll
            for (int k = 0; k < B; ++k) v[k] = c0 + 32 * k < cols ? src(c0 + 32 * k) : -INFINITY;
#pragma unroll
            for (int k = 0; k < B; ++k) m = fmax(m, v[k]);
        }
        m = warp_max(m);
        if (m == -INFINITY) return -INFINITY;
        double acc = 0.0;
#pragma unroll 4
        for (long long c = lane; c < cols; c += 32) acc = __dadd_rn(acc, exp(__dsub_rn(src(c), m)));
        return __dadd_rn(m, log(warp_add(acc)));
    };
    double r;
    if constexpr (SRC == 0) r = lse(MatrixRow{a.x + (chain * a.n_local + s) * cols});
    else r = lse(GeneratedRow(chain * a.n_total + a.star0 + s, cols, a.inv_cols));
    if (lane == 0) a.row_lse[chain * a.n_local + s] = r;
    finish_rows<PEER>(a, pa, chain, s, s + 1, lane);
}

__global__ void __launch_bounds__(256)
generate_terms_kernel(double *__restrict__ x, long long rows, long long cols, double inv_cols) {
    const long long n = rows * cols;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long r = i / cols, c = i - r * cols;
        x[i] = GeneratedRow(r, cols, inv_cols)(c);
    }
}

__global__ void no_stars_kernel(double *partials, long long n_partials, double *total, long long chains) {
    const long long stride = (long long)gridDim.x * blockDim.x;      // the empty sum, per chain
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_partials; i += stride)
        partials[i] = 0.0;
    if (total)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chains; i += stride)
            total[i] = 0.0;
}

inline int log2_of(int V) {
    int s = 0;
    while ((1 << s) < V) ++s;
    return s;
}

int env_int(const char *name, int lo, int hi, int dflt) {   // tuning aids, not an interface
    const char *e = getenv(name);
    const int n = e ? atoi(e) : 0;
    return n >= lo && n <= hi ? n : dflt;
}

// Rows a staged warp takes in sequence.  More rows per warp amortise the ticket round trip
// that ends a warp's work, but with fewer rows than ~8 per resident warp slot they only
// lengthen the critical path: measured on B200 (profiles/r02_groundwork.md), 1 is best at
// 10 000 rows and 4 at 160 000.
int rows_per_warp(long long rows) {
    static int forced = env_int("B9GW_LSE_ROWS_PER_WARP", 1, 64, 0);
    if (forced) return forced;
    static int slots = [] {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess)
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return sms * 24;                   // 6 CTAs of 4 warps fit one SM's shared memory at 1024 columns
    }();
    const long long per_slot = rows / slots;
    return per_slot >= 8 ? 4 : per_slot >= 4 ? 2 : 1;
}

template <int SRC, bool PEER = false>
cudaError_t launch_lse(cudaStream_t st, const double *x, const b9gw::LseJob &j, double *row_lse,
                       double *partials, double *total, unsigned *tickets, const PeerArgs &pa = PeerArgs{}) {
    const int vshift = log2_of(j.V);
    LseArgs a;
    a.x = x;
    a.row_lse = row_lse;
    a.partials = partials;
    a.total = PEER || j.n_shards == j.V ? total : nullptr;
    a.tickets = tickets;
    a.n_total = j.n_total;
    a.star0 = shard_lo(j.n_total, vshift, j.first_shard);
    a.n_local = shard_lo(j.n_total, vshift, j.first_shard + j.n_shards) - a.star0;
    a.chains = j.chains;
    a.inv_cols = j.cols > 0 ? 1.0 / (double)j.cols : 0.0;
    a.cols = (int)j.cols;                  // the staged kernel's; the stream kernel reads cols_ll
    a.cols_ll = j.cols;
    a.cap = 0;
    a.rpw = 1;
    a.vshift = vshift;
    a.first_shard = j.first_shard;
    a.n_shards = j.n_shards;
#ifdef B9GW_DIAG
    a.diag = 0;
    if (const char *e = getenv("B9GW_DIAG_FLAGS")) a.diag = atoi(e);
#endif
    if (j.chains == 0 || j.n_shards == 0) return cudaSuccess;
    if (j.n_total == 0) {
        // every shard on every rank is empty: the sum is +0 and no rank pushes or waits
        no_stars_kernel<<<32, 256, 0, st>>>(partials, (long long)j.n_shards * j.chains, a.total, j.chains);
    } else if (j.cols <= B9GW_LSE_STAGED_COLS) {
        a.cap = (int)((j.cols + 127) / 128 * 128);
        a.rpw = rows_per_warp(a.n_local * j.chains);
        const long long per_cta = (long long)STAGED_WARPS * a.rpw;
        const long long gx = a.n_local > 0 ? (a.n_local + per_cta - 1) / per_cta : 1;   // CTA 0 zeroes empty shards
        lse_staged_kernel<SRC, PEER><<<dim3((unsigned)gx, (unsigned)j.chains), STAGED_WARPS * 32,
                                       sizeof(double) * STAGED_WARPS * a.cap, st>>>(exp_constants(), a, pa);
    } else {
        const long long gx = a.n_local > 0 ? (a.n_local + STREAM_ROWS - 1) / STREAM_ROWS : 1;
        lse_stream_kernel<SRC, PEER><<<dim3((unsigned)gx, (unsigned)j.chains), CTA_THREADS, 0, st>>>(a, pa);
    }
    return cudaGetLastError();
}

int check_job(const b9gw::LseJob &j) {
    if (j.n_total < 0 || j.cols < 0 || j.chains < 0)
        return fail(B9GW_E_ARG, "need n_stars_total>=0, cols>=0, chains>=0");
    if (j.V < 4 || j.V > B9GW_MAX_VSHARDS || (j.V & (j.V - 1)))
        return fail(B9GW_E_ARG, "n_vshards must be a power of two in [4,128]");
    if (j.first_shard < 0 || j.n_shards < 0 || j.first_shard + j.n_shards > j.V)
        return fail(B9GW_E_ARG, "bad shard range");
    if (j.chains > B9GW_LSE_MAX_CHAINS) return fail(B9GW_E_ARG, "chains > B9GW_LSE_MAX_CHAINS");
    // the generator takes chain * n_stars_total + star as a 32-bit row number
    if (!b9gw::product_ok(j.chains, j.n_total) || j.chains * j.n_total > (1LL << 31) - 1)
        return fail(B9GW_E_ARG, "chains * n_stars_total > 2^31-1");
    if (!b9gw::product_ok(j.chains * j.n_total, j.cols) || j.cols > (1LL << 40))
        return fail(B9GW_E_ARG, "chains*n_stars_total*cols overflows (or cols > 2^40)");
    return B9GW_OK;
}

// Shared host body: SRC 0 uploads x_host, SRC 1 has no input at all.
template <int SRC>
int run_lse(int device, const double *x_host, long long rows, long long cols, int V, int warmup,
            int reps, double *row_lse_host, double *partials_host, double *total_host,
            float *ms_per_launch) {
    int rc = B9GW_OK;
    double *dx = nullptr, *dr = nullptr, *dp = nullptr, *dt = nullptr;
    unsigned *dtickets = nullptr;
    cudaStream_t st = nullptr;
    b9gw::Timer tm;
    float ms = 0.f;
    const b9gw::LseJob job{rows, cols, 1, V, 0, V};
    if (warmup < 0 || reps < 1) return fail(B9GW_E_ARG, "need warmup>=0, reps>=1");
    if ((rc = check_job(job)) != B9GW_OK) return rc;
    if (SRC == 0 && rows * cols > 0 && !x_host) return fail(B9GW_E_ARG, "x_host is null");
    if (!total_host) return fail(B9GW_E_ARG, "total_host is null");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    {
        const long long n = rows * cols;
        if (SRC == 0) {
            CK(cudaMalloc(&dx, (n > 0 ? n : 1) * sizeof(double)));
            if (n > 0) CK(cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice));
        }
        CK(cudaMalloc(&dr, (rows > 0 ? rows : 1) * sizeof(double)));
        CK(cudaMalloc(&dp, V * sizeof(double)));
        CK(cudaMalloc(&dt, sizeof(double)));
        CK(cudaMalloc(&dtickets, (V + 1) * sizeof(unsigned)));
        CK(cudaMemset(dtickets, 0, (V + 1) * sizeof(unsigned)));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            CK(launch_lse<SRC>(st, dx, job, dr, dp, dt, dtickets));
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (row_lse_host && rows > 0)
            CK(cudaMemcpy(row_lse_host, dr, rows * sizeof(double), cudaMemcpyDeviceToHost));
        if (partials_host) CK(cudaMemcpy(partials_host, dp, V * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(total_host, dt, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (ms_per_launch) *ms_per_launch = ms / reps;
done:
    if (st) cudaStreamDestroy(st);
    if (dx) cudaFree(dx);
    if (dr) cudaFree(dr);
    if (dp) cudaFree(dp);
    if (dt) cudaFree(dt);
    if (dtickets) cudaFree(dtickets);
    return rc;
}

}  // namespace

namespace b9gw {

long long lse_ticket_bytes(const LseJob &j) {
    return (long long)sizeof(unsigned) * j.chains * (j.n_shards + 1);
}

int launch_lse_generated(cudaStream_t st, const LseJob &j, double *row_lse, double *partials,
                         double *total, unsigned *tickets) {
    int rc = check_job(j);
    if (rc != B9GW_OK) return rc;
    const cudaError_t e = launch_lse<1>(st, nullptr, j, row_lse, partials, total, tickets);
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "lse kernel launch", e);
    return B9GW_OK;
}

int launch_lse_generated_step(cudaStream_t st, const LseJob &j, const PeerArgs &pa, double *row_lse,
                              double *partials, double *total, unsigned *tickets) {
    int rc = check_job(j);
    if (rc != B9GW_OK) return rc;
    if (j.chains > pa.max_chains) return fail(B9GW_E_ARG, "chains > the comm's max_chains");
    if (j.V % pa.world || j.n_shards != j.V / pa.world || j.first_shard != pa.rank * j.n_shards)
        return fail(B9GW_E_ARG, "the job's local shards are not this rank's");
    const cudaError_t e = launch_lse<1, true>(st, nullptr, j, row_lse, partials, total, tickets, pa);
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "fused lse step launch", e);
    return B9GW_OK;
}

}  // namespace b9gw

extern "C" {

int b9gw_lse_rows(int device, const double *x_host, long long rows, long long cols, int n_vshards,
                  int warmup, int reps, double *row_lse_host, double *partials_host,
                  double *total_host, float *ms_per_launch) {
    return run_lse<0>(device, x_host, rows, cols, n_vshards, warmup, reps, row_lse_host,
                      partials_host, total_host, ms_per_launch);
}

int b9gw_lse_generated(int device, long long rows, long long cols, int n_vshards, int warmup,
                       int reps, double *row_lse_host, double *partials_host, double *total_host,
                       float *ms_per_launch) {
    return run_lse<1>(device, nullptr, rows, cols, n_vshards, warmup, reps, row_lse_host,
                      partials_host, total_host, ms_per_launch);
}

#ifdef B9GW_DIAG
int b9gw_diag_dump(int device, long long chains, unsigned long long *out) {   // tools/fused_diag.py only
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    cudaDeviceSynchronize();
    (void)chains;
    cudaError_t e = cudaMemcpyFromSymbol(out, g_diag, sizeof(g_diag));
    return e == cudaSuccess ? B9GW_OK : fail(B9GW_E_CUDA, "cudaMemcpyFromSymbol", e);
}
#endif

long long b9gw_lse_workspace_bytes(long long chains, int n_shards) {
    if (chains < 0 || chains > B9GW_LSE_MAX_CHAINS || n_shards < 0 || n_shards > B9GW_MAX_VSHARDS)
        return fail(B9GW_E_ARG, "need 0<=chains<=B9GW_LSE_MAX_CHAINS, 0<=n_shards<=B9GW_MAX_VSHARDS");
    const long long b = b9gw::lse_ticket_bytes(b9gw::LseJob{0, 0, chains, B9GW_MAX_VSHARDS, 0, n_shards});
    return b > 0 ? b : (long long)sizeof(unsigned);
}

int b9gw_lse_generated_shards(int device, long long n_stars_total, long long cols, long long chains,
                              int n_vshards, int first_shard, int n_shards, double *row_lse_dev,
                              double *partial_dev, double *total_dev, void *workspace_dev,
                              void *cuda_stream) {
    const b9gw::LseJob job{n_stars_total, cols, chains, n_vshards, first_shard, n_shards};
    int rc = check_job(job);
    if (rc != B9GW_OK) return rc;
    if (chains > 0 && n_shards > 0 && (!row_lse_dev || !partial_dev || !workspace_dev))
        return fail(B9GW_E_ARG, "null device buffer");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    return b9gw::launch_lse_generated((cudaStream_t)cuda_stream, job, row_lse_dev, partial_dev,
                                      total_dev, (unsigned *)workspace_dev);
}

int b9gw_generate_terms(int device, long long rows, long long cols, double *x_host) {
    int rc = B9GW_OK, sms = 0;
    double *dx = nullptr;
    if (rows < 0 || cols < 0 || !b9gw::product_ok(rows, cols) || cols > (1LL << 40))
        return fail(B9GW_E_ARG, "need rows>=0, 0<=cols<=2^40 and rows*cols representable");
    if (rows * cols > 0 && !x_host) return fail(B9GW_E_ARG, "x_host is null");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    if (rows * cols == 0) return B9GW_OK;
    if ((rc = b9gw::sm_count_of(device, &sms)) != B9GW_OK) return rc;
    CK(cudaMalloc(&dx, rows * cols * sizeof(double)));
    generate_terms_kernel<<<sms * 8, 256>>>(dx, rows, cols, 1.0 / (double)cols);
    CK(cudaGetLastError());
    CK(cudaMemcpy(x_host, dx, rows * cols * sizeof(double), cudaMemcpyDeviceToHost));
done:
    if (dx) cudaFree(dx);
    return rc;
}

}  // extern "C"
