// b9_spec_chain.hpp — batched log-posterior evaluation under an UNMODIFIED sequential MCMC
// step, with a bit-identical accepted chain.  Header-only C++17, no dependencies.
//
// Reference-independent (the BASE-9 hot path is BLOCKED).  This file knows nothing
// about BASE-9's sampler — not its proposal, its acceptance rule, its adaptation nor how many
// random draws a step consumes — and does not imitate it.  It answers the one question about
// the chain driver that can be answered without the source ("same-seed
// accepted-chain parity with batched proposals"): a launch costs ~9 us before it does any work
// and a dependent host round trip another ~14 us, so a small cluster — or a big one sharded over
// 8 GPUs — leaves the device idle unless several proposals, and several chains, share a launch;
// but a Metropolis chain is sequential.  How can a driver evaluate many proposals per launch and
// still produce, for the same seed, exactly the chain the sequential driver produces?
//
// By never letting the batch decide anything.  The caller hands over
//   Ctx    everything a step reads or writes: current parameters and their log-posterior, the
//          random generator, adaptation counters ...  Copyable; a copy is a snapshot.
//   step   ONE step of the sequential sampler, as it is, except that it obtains log-posteriors
//          through the callable it is given:  void step(Ctx&, const Eval&).
//   batch  log-posteriors of n candidate parameter vectors in one call (one kernel launch);
//          each candidate is tagged with the chain it belongs to, so independent chains (which
//          may fit different data) share the launch.
// and each round does three things, for all chains together:
//   1. speculate  run `step` up to `depth` times on a COPY of the context with an evaluator that
//                 answers "-inf" (the proposal is rejected, the chain stays where it is) for
//                 every parameter vector it has not seen, and notes that vector down;
//   2. evaluate   one `batch` call for everything noted down; results go into a small cache
//                 keyed by the bits of the parameter vector;
//   3. replay     run `step` on the REAL context with an evaluator that answers from the cache.
//                 A step that asks for something the cache does not hold is undone (the context
//                 is restored from the snapshot taken before it) and the round ends.
// The real context only ever advances through complete steps whose every log-posterior was the
// true one, requested in the sequential order: the chain, the generator and every counter are
// bit for bit those of the sequential run, whatever the step does inside (short-circuited
// uniform draws, several evaluations per step, adaptation, delayed rejection ...).  A wrong
// guess costs evaluations, never correctness.  The guess "rejected" is right for the fraction
// (1 - a) of steps at acceptance rate a, so a round of depth K completes (1 - (1-a)^K) / a
// steps on average — about 3.5 per launch at a = 0.28, K = 16 (tests/cpp/spec_chain_test.cpp
// measures 3.53), with C independent chains C times that per launch.  Whether depth pays depends on
// what a proposal costs: base_b200/roofline.py:speculative_steps_per_s puts the rounds on the
// measured launch curve — about 3.4x at 100 stars, 2.3x at 1 250 (depth 6), 1.2x at 10 000 stars on
// one GPU (depth 3; depth 16 loses) — and independent chains pay until the launch is throughput-bound.
//
// Requirement on `batch`: the value for a parameter vector must not depend on what else is in
// the batch nor on the batch's size — which is what fixed-order reductions buy (lse.cu, vshard.cu).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace b9 {

struct SpecStats {
    std::uint64_t steps = 0;       // steps completed on the real context
    std::uint64_t rounds = 0;      // speculate / evaluate / replay rounds
    std::uint64_t launches = 0;    // batch calls (rounds that had something new to evaluate)
    std::uint64_t evaluated = 0;   // parameter vectors sent to batch
    std::uint64_t used = 0;        // cache answers consumed by completed steps
    std::uint64_t undone = 0;      // replayed steps undone because of a miss
};

template <class Ctx>
class SpeculativeDriver {
public:
    // What a step calls to get a log-posterior: n_par doubles in, one double out.
    using Eval = std::function<double(const double *params)>;
    // One step of the sequential sampler.  It must keep ALL its state in Ctx (a step that is
    // undone is undone by restoring a copy) and get every log-posterior from the Eval.
    using Step = std::function<void(Ctx &, const Eval &)>;
    // One launch: params is [n][n_par] row-major, chain[i] says whose data candidate i is to be
    // evaluated against (independent chains may fit different clusters), out receives n values.
    using Batch = std::function<void(const double *params, const std::uint32_t *chain, std::size_t n, double *out)>;
    using AfterStep = std::function<void(const Ctx &, std::size_t chain, std::uint64_t step)>;

    SpeculativeDriver(Step step, Batch batch, std::size_t n_par, std::size_t depth,
                      std::size_t cache_entries_per_chain = 1024)
        : step_(std::move(step)), batch_(std::move(batch)), n_par_(n_par), depth_(depth),
          cache_cap_(cache_entries_per_chain) {
        if (!step_ || !batch_) throw std::invalid_argument("b9::SpeculativeDriver: step and batch are required");
        if (n_par_ == 0 || depth_ == 0) throw std::invalid_argument("b9::SpeculativeDriver: n_par and depth must be >= 1");
        if (cache_cap_ < 4 * depth_) cache_cap_ = 4 * depth_;
    }

    // Advances every context by exactly `steps` steps of the sequential sampler; the chains are
    // independent and share each round's launch.  after_step(ctx, c, i), if given, runs after
    // chain c's i-th completed step (0-based within this call) — where the sequential driver
    // writes its output line.
    void run(std::vector<Ctx> &chains, std::uint64_t steps, const AfterStep &after_step = nullptr) {
        const std::size_t C = chains.size();
        if (caches_.size() != C) caches_.assign(C, Cache{});
        std::vector<std::uint64_t> done(C, 0);
        std::size_t unfinished = steps ? C : 0;
        while (unfinished) {
            ++stats_.rounds;
            params_.clear();
            owner_.clear();
            slots_.assign(C, {});
            for (std::size_t c = 0; c < C; ++c)
                if (done[c] < steps) speculate(chains[c], c, want(steps - done[c]));
            const bool launched = evaluate();
            bool stepped = false;
            for (std::size_t c = 0; c < C; ++c) {
                if (done[c] >= steps) continue;
                const std::uint64_t n = replay(chains[c], c, want(steps - done[c]), done[c], after_step);
                stepped |= n > 0;
                done[c] += n;
                if (done[c] >= steps) --unfinished;
            }
            if (!launched && !stepped)
                throw std::logic_error("b9::SpeculativeDriver: no progress — does step() call the evaluator it is given?");
        }
    }

    void run(Ctx &ctx, std::uint64_t steps, const std::function<void(const Ctx &, std::uint64_t)> &after_step = nullptr) {
        std::vector<Ctx> one;
        one.push_back(std::move(ctx));
        run(one, steps, after_step ? AfterStep([&](const Ctx &x, std::size_t, std::uint64_t i) { after_step(x, i); })
                                   : AfterStep(nullptr));
        ctx = std::move(one[0]);
    }

    const SpecStats &stats() const { return stats_; }
    void reset_stats() { stats_ = SpecStats{}; }

private:
    using Key = std::vector<std::uint64_t>;
    struct KeyHash {
        std::size_t operator()(const Key &k) const noexcept {
            std::uint64_t h = 1469598103934665603ULL;                  // FNV-1a over the words
            for (std::uint64_t w : k) {
                h ^= w;
                h *= 1099511628211ULL;
            }
            return (std::size_t)h;
        }
    };
    using Cache = std::unordered_map<Key, double, KeyHash>;
    static constexpr double kRejected = -std::numeric_limits<double>::infinity();

    std::uint64_t want(std::uint64_t left) const { return left < depth_ ? left : depth_; }

    Key key(const double *p) const {                                   // the BITS: -0.0 != +0.0, NaNs by payload
        Key k(n_par_);
        std::memcpy(k.data(), p, n_par_ * sizeof(double));
        return k;
    }

    // 1. run ahead on a copy, answering "rejected" for the unknown and noting it down
    void speculate(const Ctx &ctx, std::size_t c, std::uint64_t n_steps) {
        Ctx guess = ctx;
        Cache &cache = caches_[c];
        auto &slots = slots_[c];
        const Eval noting = [&](const double *p) -> double {
            Key k = key(p);
            const auto it = cache.find(k);
            if (it != cache.end()) return it->second;
            if (slots.emplace(std::move(k), owner_.size()).second) {
                params_.insert(params_.end(), p, p + n_par_);
                owner_.push_back((std::uint32_t)c);
            }
            return kRejected;
        };
        for (std::uint64_t j = 0; j < n_steps; ++j) step_(guess, noting);
    }

    // 2. one launch for everything every chain noted down
    bool evaluate() {
        const std::size_t n = owner_.size();
        if (n == 0) return false;
        out_.assign(n, 0.0);
        batch_(params_.data(), owner_.data(), n, out_.data());
        ++stats_.launches;
        stats_.evaluated += n;
        for (std::size_t c = 0; c < slots_.size(); ++c) {
            if (caches_[c].size() + slots_[c].size() > cache_cap_) caches_[c].clear();   // cheap to refill: no LRU
            for (auto &kv : slots_[c]) caches_[c].emplace(kv.first, out_[kv.second]);
        }
        return true;
    }

    // 3. the real context advances through complete steps fed with true values only
    std::uint64_t replay(Ctx &ctx, std::size_t c, std::uint64_t n_steps, std::uint64_t first,
                         const AfterStep &after_step) {
        const Cache &cache = caches_[c];
        std::uint64_t n = 0;
        for (; n < n_steps; ++n) {
            Ctx snapshot = ctx;
            bool missed = false;
            std::uint64_t hits = 0;
            const Eval from_cache = [&](const double *p) -> double {
                if (missed) return kRejected;
                const auto it = cache.find(key(p));
                if (it == cache.end()) {
                    missed = true;
                    return kRejected;
                }
                ++hits;
                return it->second;
            };
            step_(ctx, from_cache);
            if (missed) {
                ctx = std::move(snapshot);
                ++stats_.undone;
                break;
            }
            stats_.used += hits;
            ++stats_.steps;
            if (after_step) after_step(ctx, c, first + n);
        }
        return n;
    }

    Step step_;
    Batch batch_;
    std::size_t n_par_, depth_, cache_cap_;
    std::vector<Cache> caches_;                                        // per chain: chains may fit different data
    std::vector<std::unordered_map<Key, std::size_t, KeyHash>> slots_; // per chain: noted key -> row of this round's batch
    std::vector<double> params_, out_;
    std::vector<std::uint32_t> owner_;
    SpecStats stats_;
};

}  // namespace b9
