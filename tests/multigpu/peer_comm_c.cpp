// peer_comm_c.cpp — the cross-rank sum driven from plain C++: no Python, no torch, no NCCL,
// no CUDA headers.  This is what a C++ chain driver's binding looks like:
// it links libb9_groundwork.so, forks one process per GPU, passes the 64-byte handles through
// shared memory, and calls b9gw_ordered_allreduce once per step.
//
//   g++ -std=c++17 -O2 -Iinclude tests/multigpu/peer_comm_c.cpp -o /tmp/peer_comm_c
//       -Lbase_b200 -lb9_groundwork -Loracle -lb9_groundwork_ref -Wl,-rpath,$PWD/base_b200:$PWD/oracle
//   /tmp/peer_comm_c <world>          (world <= number of GPUs: ranks that wait on one another
//                                      must not share a device)
//
// Each rank checks its totals, bit for bit, against the CPU checker's world-independent sum
// (oracle/groundwork_ref.c — test infrastructure; a product driver would not link it).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <vector>

#include "b9_groundwork.h"

extern "C" void b9ref_vshard_total(const double *values, long long chains, long long n, int V,
                                   double *partials, double *total);

namespace {

constexpr int V = 64, CHAINS = 257, STEPS = 200;
constexpr long long N_STARS = 5003;

struct Shared {
    char handles[B9GW_MAX_WORLD][B9GW_IPC_HANDLE_BYTES];
    int arrived, generation;
    int ok[B9GW_MAX_WORLD];
    float us_stream[B9GW_MAX_WORLD], us_graph[B9GW_MAX_WORLD], us_sharded[B9GW_MAX_WORLD], us_fused[B9GW_MAX_WORLD];
};

void barrier(Shared *s, int world) {                    // sense-reversing, across processes
    const int gen = __atomic_load_n(&s->generation, __ATOMIC_ACQUIRE);
    if (__atomic_add_fetch(&s->arrived, 1, __ATOMIC_ACQ_REL) == world) {
        __atomic_store_n(&s->arrived, 0, __ATOMIC_RELAXED);
        __atomic_add_fetch(&s->generation, 1, __ATOMIC_ACQ_REL);
    } else {
        for (int waited = 0; __atomic_load_n(&s->generation, __ATOMIC_ACQUIRE) == gen; waited += 50) {
            if (waited > 30 * 1000 * 1000) {             // a rank died: do not wait for it forever
                fprintf(stderr, "barrier: a rank never arrived; giving up\n");
                _exit(3);
            }
            usleep(50);
        }
    }
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        const int rc_ = (call);                                                       \
        if (rc_ != B9GW_OK) {                                                         \
            fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc_, b9gw_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int run_rank(int rank, int world, Shared *sh) {
    // the same global per-star values on every rank (a 64-bit LCG; magnitudes over 12 decades)
    std::vector<double> values((size_t)CHAINS * N_STARS);
    uint64_t x = 88172645463325252ULL;
    for (auto &v : values) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        const double u = (double)(x >> 11) * 0x1p-53 - 0.5;
        v = u * ((x >> 7 & 1) ? 1e6 : 1e-6) * (double)(1 + (x >> 3 & 15));
    }
    std::vector<double> want(CHAINS);
    b9ref_vshard_total(values.data(), CHAINS, N_STARS, V, nullptr, want.data());

    const int per = V / world, first = rank * per;
    long long lo, hi, tmp;
    CHECK(b9gw_vshard_bounds(N_STARS, V, first, &lo, &tmp));
    CHECK(b9gw_vshard_bounds(N_STARS, V, first + per - 1, &tmp, &hi));
    const long long n_local = hi - lo;
    std::vector<double> local((size_t)CHAINS * n_local);
    for (int c = 0; c < CHAINS; ++c)
        memcpy(&local[(size_t)c * n_local], &values[(size_t)c * N_STARS + lo], n_local * sizeof(double));

    b9gw_comm *comm = nullptr;
    CHECK(b9gw_comm_create(rank, rank, world, V, 1024, &comm, sh->handles[rank]));
    barrier(sh, world);                                  // every handle is published
    CHECK(b9gw_comm_connect(comm, sh->handles));
    barrier(sh, world);                                  // everyone mapped before anyone pushes

    void *d_values, *d_partial, *d_out;
    CHECK(b9gw_dev_malloc(rank, (long long)local.size() * 8, &d_values));
    CHECK(b9gw_dev_malloc(rank, (long long)per * CHAINS * 8, &d_partial));
    CHECK(b9gw_dev_malloc(rank, CHAINS * 8, &d_out));
    CHECK(b9gw_memcpy_h2d(rank, d_values, local.data(), (long long)local.size() * 8));

    std::vector<double> got(CHAINS);
    int ok = 1;
    for (int step = 0; step < STEPS; ++step) {           // one launch pair per step, default stream
        CHECK(b9gw_shard_partials(rank, (const double *)d_values, CHAINS, n_local, N_STARS, V, first, per,
                                  (double *)d_partial, nullptr));
        CHECK(b9gw_ordered_allreduce(comm, (const double *)d_partial, (double *)d_out, CHAINS, nullptr));
        if (step % 50 == 49 || step == 0) {
            CHECK(b9gw_memcpy_d2h(rank, got.data(), d_out, CHAINS * 8));
            ok &= memcmp(got.data(), want.data(), CHAINS * 8) == 0;
            if (!ok) break;                              // a wrong sum will not get better
        }
    }
    int timed_out = 0;
    unsigned long long steps = 0;
    CHECK(b9gw_comm_status(comm, &timed_out, &steps));
    ok &= !timed_out && steps == (unsigned long long)STEPS;
    if (ok) CHECK(b9gw_allreduce_latency(comm, CHAINS, 20, 400, &sh->us_stream[rank], &sh->us_graph[rank]));
    if (ok) {
        // a star-sharded step end to end (this rank's shards' log-sum-exp, then the cross-rank sum)
        // against the same job done by this rank alone from all V shards in one launch
        const long long sn = 3001, sc = 160, sch = 9;
        std::vector<double> stepped(sch), fused(sch), alone(sch);
        float us_lse = 0;
        CHECK(b9gw_sharded_step(comm, sn, sc, sch, 2, 10, stepped.data(), fused.data(), &sh->us_sharded[rank],
                                &us_lse, &sh->us_fused[rank]));
        void *d_rows, *d_p, *d_t, *d_ws;
        CHECK(b9gw_dev_malloc(rank, sch * sn * 8, &d_rows));
        CHECK(b9gw_dev_malloc(rank, sch * V * 8, &d_p));
        CHECK(b9gw_dev_malloc(rank, sch * 8, &d_t));
        CHECK(b9gw_dev_malloc(rank, b9gw_lse_workspace_bytes(sch, V), &d_ws));
        CHECK(b9gw_lse_generated_shards(rank, sn, sc, sch, V, 0, V, (double *)d_rows, (double *)d_p,
                                        (double *)d_t, d_ws, nullptr));
        CHECK(b9gw_memcpy_d2h(rank, alone.data(), d_t, sch * 8));
        ok &= memcmp(stepped.data(), alone.data(), sch * 8) == 0;
        ok &= memcmp(fused.data(), alone.data(), sch * 8) == 0;
        b9gw_dev_free(rank, d_rows);
        b9gw_dev_free(rank, d_p);
        b9gw_dev_free(rank, d_t);
        b9gw_dev_free(rank, d_ws);
    }
    sh->ok[rank] = ok;
    barrier(sh, world);                                  // nobody frees a mailbox a peer still writes
    b9gw_dev_free(rank, d_values);
    b9gw_dev_free(rank, d_partial);
    b9gw_dev_free(rank, d_out);
    b9gw_comm_destroy(comm);
    return ok ? 0 : 2;
}

}  // namespace

int main(int argc, char **argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 1;
    if (world < 1 || world > B9GW_MAX_WORLD || V % world) {
        fprintf(stderr, "usage: %s <world dividing %d>\n", argv[0], V);
        return 64;
    }
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return 65;
    memset(sh, 0, sizeof *sh);
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r) {                    // fork BEFORE any CUDA call
        const pid_t p = fork();
        if (p == 0) _exit(run_rank(r, world, sh));
        kids.push_back(p);
    }
    int bad = 0;
    for (pid_t p : kids) {
        int st = 0;
        waitpid(p, &st, 0);
        bad |= !WIFEXITED(st) || WEXITSTATUS(st) != 0;
    }
    float us_s = 0, us_g = 0, us_sh = 0, us_fu = 0;
    for (int r = 0; r < world; ++r) {
        bad |= !sh->ok[r];
        us_s = sh->us_stream[r] > us_s ? sh->us_stream[r] : us_s;
        us_g = sh->us_graph[r] > us_g ? sh->us_graph[r] : us_g;
        us_sh = sh->us_sharded[r] > us_sh ? sh->us_sharded[r] : us_sh;
        us_fu = sh->us_fused[r] > us_fu ? sh->us_fused[r] : us_fu;
    }
    printf("{\"world\": %d, \"bits_equal_checker\": %s, \"steps\": %d, \"us_stream_max\": %.2f, "
           "\"us_graph_max\": %.2f, \"sharded_step_us_max\": %.2f, \"fused_step_us_max\": %.2f, \"driver\": \"C++ (no Python, no NCCL)\"}\n",
           world, bad ? "false" : "true", STEPS, us_s, us_g, us_sh, us_fu);
    return bad;
}
