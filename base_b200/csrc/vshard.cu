// vshard.cu — world-size-independent sum of per-chain FP64 scalars over ranks,
// as ONE kernel over NVLink peer memory.  sm_100a.
//
// Reference-independent groundwork (the BASE-9 hot path is BLOCKED).
// north_star's only collective is "a single NCCL allreduce of per-chain
// log-likelihood scalars over NVLink per step".  Two things are wrong with
// taking that literally on the step path of a chain whose accept/reject compares
// the sum with a uniform draw:
//   (1) the bits of a rank-ordered (or ring/tree) sum depend on the world size,
//       because the per-rank partials group the stars differently at W = 2 and 8;
//   (2) an NCCL call costs a launch per collective plus a reduction kernel, and
//       the payload is only chains x 8 B: latency is everything.
// So the sum is DEFINED over V fixed virtual shards (include/b9_groundwork.h,
// "world-size-independent sum"), and implemented here as a single kernel:
//
//   push : thread (chain, q) takes its rank's partials P[v][chain] for the owned
//          shards v = first + q, first + q + 4, ... and stores each one into
//          EVERY rank's mailbox slot [parity][v][chain] as a 16-byte packet
//          {lo32, step, hi32, step}.  The two 8-byte halves are each atomic and
//          each carries the step number, so a packet is its own arrival flag:
//          no __threadfence_system, no separate flag write, no barrier.
//   pull : the four lanes (chain, 0..3) poll V/4 local slots each (all loads of a
//          lane in flight together) until every packet shows this step;
//   add  : lane 0 adds its V/4 shards left to right starting from +0, hands the
//          running sum to lane 1 by shuffle, and so on: exactly
//          (((0 + P[0]) + P[1]) + ...) + P[V-1].  Lane 3 stores total[chain].
//
// Mailboxes are double-buffered by step parity.  That is enough: a rank can be at
// most one step ahead of a peer (it cannot finish step s+1 without the peer's
// s+1 packets, which the peer sends only after its own step-s kernel — all its
// reads of parity s — has retired), so parity s is never overwritten while read.
// The step counter lives in device memory (one per chain, bumped by the kernel), so
// a captured launch replays correctly from a CUDA graph.
// A peer that never arrives costs `timeout_ns`, not a hang: the waiting thread
// gives up, stores NaN and raises the comm's sticky status.

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace {

using b9gw::fail;

constexpr unsigned FULL = 0xffffffffu;
constexpr int STEP_THREADS = 256;          // 64 chains per CTA, 4 lanes per chain
constexpr int PART_WARPS = 8;
constexpr long long MAX_CHAINS = 1LL << 22;

using b9gw::PeerArgs;
using b9gw::st_packet;
using b9gw::ld_packet;
using b9gw::globaltimer_ns;

struct StepArgs : PeerArgs {
    const double *partial;                 // [V/world][chains]
    double *out;                           // [chains]
    long long chains;
};

template <int SLOTS>                       // V = 4 * SLOTS virtual shards
__global__ void __launch_bounds__(STEP_THREADS)
vshard_step_kernel(const StepArgs a) {
    constexpr int V = 4 * SLOTS;
    const long long tid = (long long)blockIdx.x * STEP_THREADS + threadIdx.x;
    const long long chain = tid >> 2;
    const int q = (int)(tid & 3);
    const int lane = threadIdx.x & 31;
    const bool active = chain < a.chains;
    const unsigned step = (active ? a.seq[chain] : 0u) + 1u;   // a chain's four lanes read it here,
    const size_t parity_base = (size_t)(step & 1u) * V;        // before lane 3 moves it below
    bool ok = true;
    double val[SLOTS];
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) val[i] = 0.0;

    const uint4 *mine = nullptr;
    unsigned long long t0 = 0;
    if (active) {
        // ---- push: own shards to every rank, remote peers first, self last
        const int per = V / a.world, first = a.rank * per;
        for (int k = q; k < per; k += 4) {
            const double p = a.partial[(long long)k * a.chains + chain];
            const unsigned lo = (unsigned)__double2loint(p), hi = (unsigned)__double2hiint(p);
            const size_t slot = (parity_base + first + k) * (size_t)a.max_chains + (size_t)chain;
            for (int d = 1; d <= a.world; ++d) {
                int peer = a.rank + d;
                if (peer >= a.world) peer -= a.world;
                st_packet(a.mail[peer] + slot, lo, hi, step);
            }
        }
        // ---- pull: shards q*SLOTS .. q*SLOTS+SLOTS-1 of this chain, from the local mailbox
        mine = a.mail[a.rank] + (parity_base + (size_t)q * SLOTS) * (size_t)a.max_chains + (size_t)chain;
        // a comm that has already timed out is out of step with its peers for good: do not
        // spend another timeout per launch on it
        t0 = *(volatile int *)a.status ? ~0ULL : 0;
    }
    // The warp polls as a whole and leaves the loop on a vote, so that it is still converged
    // for the shuffles below (lanes leaving one by one made everything behind a poll loop run
    // an order of magnitude slower in the fused kernel: profiles/r02b_groundwork.md).
    bool got = !active;
    for (;;) {
        if (!got) {
            uint4 r[SLOTS];
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) r[i] = ld_packet(mine + (size_t)i * a.max_chains);
            bool ready = true;
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) ready &= (r[i].y == step) & (r[i].w == step);
            if (ready) {
#pragma unroll
                for (int i = 0; i < SLOTS; ++i) val[i] = __hiloint2double((int)r[i].z, (int)r[i].x);
                got = true;
            }
        }
        if (__all_sync(FULL, got)) break;
        bool expired = false;
        if (!got) {
            expired = t0 == ~0ULL;
            if (!expired) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else expired = now - t0 > a.timeout_ns;
            }
        }
        if (__any_sync(FULL, expired)) { ok = got; break; }
    }

    // ---- add: strictly left to right over v = 0 .. V-1, handed lane to lane
    double acc = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const double in = __shfl_up_sync(FULL, acc, 1);
        if (q == g) {
            if (g > 0) acc = in;
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) acc = __dadd_rn(acc, val[i]);
        }
    }
    const unsigned bad = __ballot_sync(FULL, !ok);
    if (active && q == 3) {
        const bool chain_bad = ((bad >> (lane & ~3)) & 0xFu) != 0;
        a.out[chain] = chain_bad ? __longlong_as_double(0x7ff8000000000000LL) : acc;
        if (chain_bad) atomicExch(a.status, 1);
        a.seq[chain] = step;               // after the shuffles above: the chain's lanes have read it
    }
}

// floor(v*n/V) for V a power of two (shift = log2 V), v <= V <= 128 and n <= 2^48: fits 64 bits.
constexpr long long MAX_STARS = 1LL << 48;
__host__ __device__ inline long long shard_lo(long long n, int shift, int v) {
    return (long long)(((unsigned long long)v * (unsigned long long)n) >> shift);
}
inline int log2_of(int V) {
    int s = 0;
    while ((1 << s) < V) ++s;
    return s;
}

// P[k][chain] for local shards k = 0 .. n_shards-1: one warp per (k, chain), warp order.
__global__ void __launch_bounds__(PART_WARPS * 32)
shard_partials_kernel(const double *__restrict__ values, long long chains, long long ld,
                      long long n_total, int vshift, int first_shard, int n_shards,
                      double *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * PART_WARPS + (threadIdx.x >> 5);
    if (w >= (long long)n_shards * chains) return;      // whole warp
    const int k = (int)(w / chains);
    const long long chain = w - (long long)k * chains;
    const long long base = shard_lo(n_total, vshift, first_shard);
    const long long lo = shard_lo(n_total, vshift, first_shard + k) - base;
    const long long hi = shard_lo(n_total, vshift, first_shard + k + 1) - base;
    const double *row = values + chain * ld;
    double s = 0.0;
    for (long long i0 = lo + lane; i0 < hi; i0 += 32 * 4) {     // 4 loads in flight, adds in order
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = i0 + 32 * u < hi ? __ldg(row + i0 + 32 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < hi) s = __dadd_rn(s, x[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(FULL, s, o));
    if (lane == 0) partial[(long long)k * chains + chain] = s;
}

bool vshards_ok(int V) { return V >= 4 && V <= B9GW_MAX_VSHARDS && (V & (V - 1)) == 0; }

}  // namespace

struct b9gw_comm {
    int device = -1, rank = 0, world = 1, V = 0;
    long long max_chains = 0;
    uint4 *mail[B9GW_MAX_WORLD] = {};
    bool mapped[B9GW_MAX_WORLD] = {};
    unsigned *seq = nullptr;
    int *status = nullptr;
    unsigned long long timeout_ns = 2000ULL * 1000 * 1000;
    bool connected = false;
};

namespace {

PeerArgs peer_args(const b9gw_comm *c) {
    PeerArgs a;
    for (int r = 0; r < B9GW_MAX_WORLD; ++r) a.mail[r] = c->mail[r];
    a.seq = c->seq;
    a.status = c->status;
    a.max_chains = c->max_chains;
    a.timeout_ns = c->timeout_ns;
    a.rank = c->rank;
    a.world = c->world;
    return a;
}

int launch_step(b9gw_comm *c, const double *partial_dev, double *out_dev, long long chains,
                cudaStream_t st) {
    StepArgs a;
    static_cast<PeerArgs &>(a) = peer_args(c);
    a.partial = partial_dev;
    a.out = out_dev;
    a.chains = chains;
    if (chains == 0) return B9GW_OK;
    const unsigned grid = (unsigned)((chains * 4 + STEP_THREADS - 1) / STEP_THREADS);
    switch (c->V) {
        case 4:   vshard_step_kernel<1><<<grid, STEP_THREADS, 0, st>>>(a); break;
        case 8:   vshard_step_kernel<2><<<grid, STEP_THREADS, 0, st>>>(a); break;
        case 16:  vshard_step_kernel<4><<<grid, STEP_THREADS, 0, st>>>(a); break;
        case 32:  vshard_step_kernel<8><<<grid, STEP_THREADS, 0, st>>>(a); break;
        case 64:  vshard_step_kernel<16><<<grid, STEP_THREADS, 0, st>>>(a); break;
        case 128: vshard_step_kernel<32><<<grid, STEP_THREADS, 0, st>>>(a); break;
        default: return fail(B9GW_E_STATE, "comm has an unsupported shard count");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "vshard_step_kernel launch", e);
    return B9GW_OK;
}

int launch_partials(const double *values_dev, long long chains, long long ld, long long n_total,
                    int V, int first_shard, int n_shards, double *partial_dev, cudaStream_t st) {
    const long long warps = (long long)n_shards * chains;
    if (warps == 0) return B9GW_OK;
    const long long grid = (warps + PART_WARPS - 1) / PART_WARPS;
    if (grid > 0x7fffffffLL) return fail(B9GW_E_ARG, "n_shards*chains too large for one launch");
    shard_partials_kernel<<<(unsigned)grid, PART_WARPS * 32, 0, st>>>(
        values_dev, chains, ld, n_total, log2_of(V), first_shard, n_shards, partial_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B9GW_E_CUDA, "shard_partials_kernel launch", e);
    return B9GW_OK;
}

}  // namespace

extern "C" {

int b9gw_vshard_bounds(long long n_stars, int n_vshards, int shard, long long *lo, long long *hi) {
    if (!vshards_ok(n_vshards) || shard < 0 || shard >= n_vshards || n_stars < 0 || n_stars > MAX_STARS)
        return fail(B9GW_E_ARG, "need a power-of-two shard count in [4,128], 0<=shard<V, 0<=n_stars<=2^48");
    if (lo) *lo = shard_lo(n_stars, log2_of(n_vshards), shard);
    if (hi) *hi = shard_lo(n_stars, log2_of(n_vshards), shard + 1);
    return B9GW_OK;
}

int b9gw_shard_partials(int device, const double *values_dev, long long chains, long long ld,
                        long long n_stars_total, int n_vshards, int first_shard, int n_shards,
                        double *partial_dev, void *cuda_stream) {
    if (!vshards_ok(n_vshards) || first_shard < 0 || n_shards < 0 ||
        first_shard + n_shards > n_vshards)
        return fail(B9GW_E_ARG, "bad shard range");
    if (chains < 0 || chains > MAX_CHAINS || ld < 0 || n_stars_total < 0 ||
        n_stars_total > MAX_STARS || !b9gw::product_ok(chains, ld))
        return fail(B9GW_E_ARG, "bad chains / ld / n_stars_total");
    const int sh = log2_of(n_vshards);
    const long long n_local = shard_lo(n_stars_total, sh, first_shard + n_shards) -
                              shard_lo(n_stars_total, sh, first_shard);
    if (ld < n_local) return fail(B9GW_E_ARG, "ld is smaller than the local star count");
    if ((long long)n_shards * chains > 0 && (!partial_dev || (n_local > 0 && !values_dev)))
        return fail(B9GW_E_ARG, "null device buffer");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    return launch_partials(values_dev, chains, ld, n_stars_total, n_vshards, first_shard, n_shards,
                           partial_dev, (cudaStream_t)cuda_stream);
}

int b9gw_comm_create(int device, int rank, int world, int n_vshards, long long max_chains,
                     b9gw_comm **comm, void *handle_out) {
    int rc = B9GW_OK;
    if (!comm || !handle_out) return fail(B9GW_E_ARG, "comm / handle_out is null");
    *comm = nullptr;
    if (world < 1 || world > B9GW_MAX_WORLD || rank < 0 || rank >= world)
        return fail(B9GW_E_ARG, "need 1<=world<=B9GW_MAX_WORLD and 0<=rank<world");
    if (!vshards_ok(n_vshards) || n_vshards % world != 0)
        return fail(B9GW_E_ARG, "n_vshards must be a power of two in [4,128] and a multiple of world");
    if (max_chains < 1 || max_chains > MAX_CHAINS)
        return fail(B9GW_E_ARG, "need 1<=max_chains<=2^22");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    b9gw_comm *c = new b9gw_comm;
    c->device = device;
    c->rank = rank;
    c->world = world;
    c->V = n_vshards;
    c->max_chains = max_chains;
    {
        const size_t mail_bytes = (size_t)2 * n_vshards * (size_t)max_chains * sizeof(uint4);
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof h);
        static_assert(sizeof(cudaIpcMemHandle_t) == B9GW_IPC_HANDLE_BYTES, "handle size");
        CK(cudaMalloc(&c->mail[rank], mail_bytes));
        CK(cudaMemset(c->mail[rank], 0, mail_bytes));
        CK(cudaMalloc(&c->seq, max_chains * sizeof(unsigned)));
        CK(cudaMemset(c->seq, 0, max_chains * sizeof(unsigned)));
        CK(cudaMalloc(&c->status, sizeof(int)));
        CK(cudaMemset(c->status, 0, sizeof(int)));
        CK(cudaDeviceSynchronize());       // zeroed before any peer can learn the handle
        if (world > 1) CK(cudaIpcGetMemHandle(&h, c->mail[rank]));
        memcpy(handle_out, &h, sizeof h);
        c->connected = world == 1;
    }
    *comm = c;
    return B9GW_OK;
done:
    if (c->mail[rank]) cudaFree(c->mail[rank]);
    if (c->seq) cudaFree(c->seq);
    if (c->status) cudaFree(c->status);
    delete c;
    return rc;
}

int b9gw_comm_connect(b9gw_comm *c, const void *all_handles) {
    int rc = B9GW_OK;
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    if (c->connected) return B9GW_OK;
    if (!all_handles) return fail(B9GW_E_ARG, "all_handles is null");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * B9GW_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->mail[r] = (uint4 *)p;
        c->mapped[r] = true;
    }
    c->connected = true;
    return B9GW_OK;
done:
    for (int r = 0; r < c->world; ++r)
        if (c->mapped[r]) {
            cudaIpcCloseMemHandle(c->mail[r]);
            c->mail[r] = nullptr;
            c->mapped[r] = false;
        }
    return rc;
}

int b9gw_ordered_allreduce(b9gw_comm *c, const double *partial_dev, double *out_dev,
                           long long chains, void *cuda_stream) {
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    if (!c->connected) return fail(B9GW_E_STATE, "comm is not connected (call b9gw_comm_connect)");
    if (chains < 0 || chains > c->max_chains) return fail(B9GW_E_ARG, "chains outside [0, max_chains]");
    if (chains > 0 && (!partial_dev || !out_dev)) return fail(B9GW_E_ARG, "null device buffer");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    return launch_step(c, partial_dev, out_dev, chains, (cudaStream_t)cuda_stream);
}

int b9gw_comm_set_timeout_ms(b9gw_comm *c, int ms) {
    if (!c || ms < 1) return fail(B9GW_E_ARG, "need a comm and ms>=1");
    c->timeout_ns = (unsigned long long)ms * 1000ULL * 1000ULL;
    return B9GW_OK;
}

int b9gw_comm_status(b9gw_comm *c, int *timed_out, unsigned long long *steps) {
    int rc = B9GW_OK, flag = 0;
    unsigned s = 0;
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&flag, c->status, sizeof flag, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&s, c->seq, sizeof s, cudaMemcpyDeviceToHost));
    if (timed_out) *timed_out = flag;
    if (steps) *steps = s;
    if (flag) rc = fail(B9GW_E_TIMEOUT, "a peer did not arrive within the comm's timeout; outputs of that step are NaN");
done:
    return rc;
}

int b9gw_allreduce_latency(b9gw_comm *c, long long chains, int warmup, int reps,
                           float *us_stream, float *us_graph) {
    constexpr int G = 8;                   // steps per graph launch
    int rc = B9GW_OK, flag = 0;
    double *dp = nullptr, *dout = nullptr;
    cudaStream_t st = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool capturing = false;
    b9gw::Timer tm;
    float ms = 0.f;
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    if (!c->connected) return fail(B9GW_E_STATE, "comm is not connected");
    if (chains < 1 || chains > c->max_chains || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need 1<=chains<=max_chains, warmup>=0, reps>=1");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    {
        const int per = c->V / c->world;
        const int launches = (reps + G - 1) / G;
        CK(cudaMalloc(&dp, (size_t)per * chains * sizeof(double)));
        CK(cudaMemset(dp, 0, (size_t)per * chains * sizeof(double)));
        CK(cudaMalloc(&dout, chains * sizeof(double)));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        for (int i = 0; i < warmup + reps; ++i) {
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            if ((rc = launch_step(c, dp, dout, chains, st)) != B9GW_OK) goto done;
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (us_stream) *us_stream = ms * 1e3f / reps;

        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        for (int i = 0; i < G; ++i)
            if ((rc = launch_step(c, dp, dout, chains, st)) != B9GW_OK) goto done;
        capturing = false;
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphLaunch(exec, st));     // one untimed replay
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(tm.a, st));
        for (int i = 0; i < launches; ++i) CK(cudaGraphLaunch(exec, st));
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (us_graph) *us_graph = ms * 1e3f / (launches * G);
        CK(cudaMemcpy(&flag, c->status, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag) rc = fail(B9GW_E_TIMEOUT, "a peer did not arrive within the comm's timeout");
    }
done:
    if (capturing) {
        cudaGraph_t dead = nullptr;
        cudaStreamEndCapture(st, &dead);
        if (dead) cudaGraphDestroy(dead);
        cudaGetLastError();
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (st) cudaStreamDestroy(st);
    if (dp) cudaFree(dp);
    if (dout) cudaFree(dout);
    return rc;
}

int b9gw_lse_generated_step(b9gw_comm *c, long long n_stars_total, long long cols, long long chains,
                            double *row_lse_dev, double *partial_dev, double *total_dev,
                            void *workspace_dev, void *cuda_stream) {
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    if (!c->connected) return fail(B9GW_E_STATE, "comm is not connected (call b9gw_comm_connect)");
    if (chains < 0 || chains > c->max_chains) return fail(B9GW_E_ARG, "chains outside [0, max_chains]");
    if (chains > 0 && (!row_lse_dev || !partial_dev || !total_dev || !workspace_dev))
        return fail(B9GW_E_ARG, "null device buffer");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    const int per = c->V / c->world;
    const b9gw::LseJob job{n_stars_total, cols, chains, c->V, c->rank * per, per};
    return b9gw::launch_lse_generated_step((cudaStream_t)cuda_stream, job, peer_args(c), row_lse_dev,
                                           partial_dev, total_dev, (unsigned *)workspace_dev);
}

int b9gw_sharded_step(b9gw_comm *c, long long n_stars_total, long long cols, long long chains,
                      int warmup, int reps, double *total_host, double *total_fused_host,
                      float *us_step, float *us_lse_alone, float *us_fused_step) {
    int rc = B9GW_OK, flag = 0;
    double *drow = nullptr, *dp = nullptr, *dout = nullptr;
    unsigned *dtk = nullptr;
    cudaStream_t st = nullptr;
    b9gw::Timer tm;
    float ms = 0.f;
    if (!c) return fail(B9GW_E_ARG, "comm is null");
    if (!c->connected) return fail(B9GW_E_STATE, "comm is not connected");
    if (chains < 1 || chains > c->max_chains || warmup < 0 || reps < 1)
        return fail(B9GW_E_ARG, "need 1<=chains<=max_chains, warmup>=0, reps>=1");
    if (n_stars_total < 0 || n_stars_total > MAX_STARS) return fail(B9GW_E_ARG, "bad n_stars_total");
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    {
        const int per = c->V / c->world, sh = log2_of(c->V);
        const b9gw::LseJob job{n_stars_total, cols, chains, c->V, c->rank * per, per};
        const long long n_local = shard_lo(n_stars_total, sh, job.first_shard + per) -
                                  shard_lo(n_stars_total, sh, job.first_shard);
        const long long tk_bytes = b9gw::lse_ticket_bytes(job);
        CK(cudaMalloc(&drow, (size_t)(n_local * chains > 0 ? n_local * chains : 1) * sizeof(double)));
        CK(cudaMalloc(&dp, (size_t)per * chains * sizeof(double)));
        CK(cudaMalloc(&dout, chains * sizeof(double)));
        CK(cudaMalloc(&dtk, tk_bytes));
        CK(cudaMemset(dtk, 0, tk_bytes));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(tm.init());
        // with one rank every shard is local and the LSE launch could add the shards itself;
        // the step is kept the same two kernels at every world size so the times compare
        for (int i = 0; i < warmup + reps; ++i) {           // the LSE share alone (no peer involved)
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            if ((rc = b9gw::launch_lse_generated(st, job, drow, dp, nullptr, dtk)) != B9GW_OK) goto done;
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (us_lse_alone) *us_lse_alone = ms * 1e3f / reps;
        for (int i = 0; i < warmup + reps; ++i) {           // the step: LSE share, then the cross-rank sum
            if (i == warmup) {
                CK(cudaStreamSynchronize(st));
                CK(cudaEventRecord(tm.a, st));
            }
            if ((rc = b9gw::launch_lse_generated(st, job, drow, dp, nullptr, dtk)) != B9GW_OK) goto done;
            if ((rc = launch_step(c, dp, dout, chains, st)) != B9GW_OK) goto done;
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (us_step) *us_step = ms * 1e3f / reps;
        if (total_host) CK(cudaMemcpy(total_host, dout, chains * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemset(dout, 0xff, chains * sizeof(double)));
        {
            const b9gw::PeerArgs pa = peer_args(c);
            for (int i = 0; i < warmup + reps; ++i) {       // the same step as ONE kernel
                if (i == warmup) {
                    CK(cudaStreamSynchronize(st));
                    CK(cudaEventRecord(tm.a, st));
                }
                if ((rc = b9gw::launch_lse_generated_step(st, job, pa, drow, dp, dout, dtk)) != B9GW_OK) goto done;
            }
        }
        CK(cudaEventRecord(tm.b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, tm.a, tm.b));
        if (us_fused_step) *us_fused_step = ms * 1e3f / reps;
        if (total_fused_host)
            CK(cudaMemcpy(total_fused_host, dout, chains * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&flag, c->status, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag) rc = fail(B9GW_E_TIMEOUT, "a peer did not arrive within the comm's timeout");
    }
done:
    if (st) cudaStreamDestroy(st);
    if (drow) cudaFree(drow);
    if (dp) cudaFree(dp);
    if (dout) cudaFree(dout);
    if (dtk) cudaFree(dtk);
    return rc;
}

int b9gw_comm_destroy(b9gw_comm *c) {
    if (!c) return B9GW_OK;
    b9gw::DeviceGuard guard(c->device);
    if (guard.rc() == B9GW_OK) {
        cudaDeviceSynchronize();
        for (int r = 0; r < c->world; ++r)
            if (c->mapped[r]) cudaIpcCloseMemHandle(c->mail[r]);
        if (c->mail[c->rank]) cudaFree(c->mail[c->rank]);
        if (c->seq) cudaFree(c->seq);
        if (c->status) cudaFree(c->status);
        cudaGetLastError();
    }
    delete c;
    return B9GW_OK;
}

int b9gw_vshard_total(int device, const double *values_host, long long chains, long long n_stars,
                      int n_vshards, double *partials_host, double *total_host) {
    int rc = B9GW_OK;
    double *dv = nullptr, *dp = nullptr, *dout = nullptr;
    b9gw_comm *c = nullptr;
    char handle[B9GW_IPC_HANDLE_BYTES];
    if (!vshards_ok(n_vshards)) return fail(B9GW_E_ARG, "n_vshards must be a power of two in [4,128]");
    if (chains < 0 || chains > MAX_CHAINS || n_stars < 0 || n_stars > MAX_STARS ||
        !b9gw::product_ok(chains, n_stars))
        return fail(B9GW_E_ARG, "bad chains / n_stars");
    if (chains > 0 && (!total_host || (n_stars > 0 && !values_host)))
        return fail(B9GW_E_ARG, "null host buffer");
    b9gw::DeviceGuard guard(device);
    if (guard.rc() != B9GW_OK) return guard.rc();
    if (chains == 0) return B9GW_OK;
    if ((rc = b9gw_comm_create(device, 0, 1, n_vshards, chains, &c, handle)) != B9GW_OK) return rc;
    {
        const size_t nv = (size_t)chains * (size_t)n_stars;
        CK(cudaMalloc(&dv, (nv ? nv : 1) * sizeof(double)));
        if (nv) CK(cudaMemcpy(dv, values_host, nv * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dp, (size_t)n_vshards * chains * sizeof(double)));
        CK(cudaMalloc(&dout, chains * sizeof(double)));
        if ((rc = launch_partials(dv, chains, n_stars, n_stars, n_vshards, 0, n_vshards, dp, nullptr)) != B9GW_OK)
            goto done;
        if ((rc = launch_step(c, dp, dout, chains, nullptr)) != B9GW_OK) goto done;
        if ((rc = b9gw_comm_status(c, nullptr, nullptr)) != B9GW_OK) goto done;
        if (partials_host)
            CK(cudaMemcpy(partials_host, dp, (size_t)n_vshards * chains * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(total_host, dout, chains * sizeof(double), cudaMemcpyDeviceToHost));
    }
done:
    if (dv) cudaFree(dv);
    if (dp) cudaFree(dp);
    if (dout) cudaFree(dout);
    b9gw_comm_destroy(c);
    return rc;
}

}  // extern "C"
