// spec_chain_test.cpp — b9::SpeculativeDriver against plain sequential runs of three toy
// Metropolis samplers.  Reference-independent: the samplers and the target are made up here for
// the test (they are deliberately awkward, not BASE-9's).  Prints one JSON line per case; exit
// code 1 if any speculative chain differs in any bit from its sequential twin.
//
//   g++ -std=c++17 -O2 -Wall -Wextra -Werror -Iinclude tests/cpp/spec_chain_test.cpp -o spec_chain_test
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "b9_spec_chain.hpp"

namespace {

constexpr std::size_t P = 4;

struct Ctx {                               // everything a step touches
    std::array<double, P> cur{};
    double cur_lp = std::nan("");
    std::mt19937_64 rng;
    std::normal_distribution<double> gauss{0.0, 1.0};     // keeps a spare draw inside: part of the snapshot
    double scale = 0.5;
    unsigned long long steps = 0, accepted = 0, window_acc = 0;
};

// The "likelihood": depends on whose data it is (chain), has a region of true -inf.
double target(const double *p, unsigned chain) {
    double s = 0.0;
    for (std::size_t i = 0; i < P; ++i) {
        if (std::fabs(p[i]) > 6.0) return -INFINITY;
        const double mu = 0.3 * (double)i - 0.1 * (double)(chain % 7), sd = 0.5 + 0.25 * (double)i;
        s -= 0.5 * (p[i] - mu) * (p[i] - mu) / (sd * sd);
    }
    return s + std::log1p(std::sin(p[0] * p[1]) * std::sin(p[0] * p[1]));
}

using Eval = b9::SpeculativeDriver<Ctx>::Eval;

void adapt(Ctx &c) {                       // step-size adaptation on the acceptance count of a window
    if (c.steps % 25 == 0) {
        c.scale *= c.window_acc > 10 ? 1.25 : c.window_acc < 4 ? 0.8 : 1.0;
        c.window_acc = 0;
    }
}

// mode 0: textbook — propose all parameters, always draw the uniform.
// mode 1: the uniform is drawn only when the proposal is worse (draw count depends on the values).
// mode 2: two blocks per step, the second proposed from whatever the first left behind; every 100
//         steps the current point's log-posterior is asked for again.
template <int MODE>
void step(Ctx &c, const Eval &eval) {
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    if (std::isnan(c.cur_lp)) c.cur_lp = eval(c.cur.data());
    if (MODE == 2 && c.steps % 100 == 99) c.cur_lp = eval(c.cur.data());
    const int blocks = MODE == 2 ? 2 : 1;
    for (int b = 0; b < blocks; ++b) {
        std::array<double, P> cand = c.cur;
        const std::size_t lo = MODE == 2 ? (std::size_t)b * 2 : 0, hi = MODE == 2 ? lo + 2 : P;
        for (std::size_t i = lo; i < hi; ++i) cand[i] += c.scale * c.gauss(c.rng);
        const double lp = eval(cand.data());
        bool take;
        if (MODE == 1) take = lp >= c.cur_lp || std::log(unif(c.rng)) < lp - c.cur_lp;
        else {
            const double u = unif(c.rng);
            take = std::log(u) < lp - c.cur_lp;
        }
        if (take) {
            c.cur = cand;
            c.cur_lp = lp;
            ++c.accepted;
            ++c.window_acc;
        }
    }
    ++c.steps;
    adapt(c);
}

struct Trace {
    std::vector<double> v;                 // [chain][step][P + 1]
    void put(std::size_t chains, std::uint64_t steps, std::size_t c, std::uint64_t i, const Ctx &x) {
        if (v.empty()) v.assign(chains * steps * (P + 1), 0.0);
        double *d = &v[(c * steps + i) * (P + 1)];
        std::memcpy(d, x.cur.data(), P * sizeof(double));
        d[P] = x.cur_lp;
    }
};

std::vector<Ctx> fresh(std::size_t chains, unsigned seed) {
    std::vector<Ctx> v(chains);
    for (std::size_t c = 0; c < chains; ++c) {
        v[c].rng.seed(seed + 1000u * (unsigned)c);
        for (std::size_t i = 0; i < P; ++i) v[c].cur[i] = 0.1 * (double)(i + c % 3);
    }
    return v;
}

template <int MODE>
bool run_case(std::size_t chains, std::uint64_t steps, std::size_t depth, unsigned seed) {
    // the sequential twin: the same step, fed directly
    std::vector<Ctx> seq = fresh(chains, seed);
    Trace want;
    unsigned long long seq_evals = 0;
    for (std::size_t c = 0; c < chains; ++c) {
        const Eval direct = [&](const double *p) { ++seq_evals; return target(p, (unsigned)c); };
        for (std::uint64_t i = 0; i < steps; ++i) {
            step<MODE>(seq[c], direct);
            want.put(chains, steps, c, i, seq[c]);
        }
    }
    // the speculative driver
    std::vector<Ctx> spec = fresh(chains, seed);
    Trace got;
    std::size_t largest = 0;
    b9::SpeculativeDriver<Ctx> drv(
        step<MODE>,
        [&](const double *params, const std::uint32_t *chain, std::size_t n, double *out) {
            largest = n > largest ? n : largest;
            for (std::size_t k = n; k-- > 0;) out[k] = target(params + k * P, chain[k]);   // any order
        },
        P, depth);
    // in two calls, to show a run can be resumed
    const std::uint64_t first = steps / 3;
    drv.run(spec, first, [&](const Ctx &x, std::size_t c, std::uint64_t i) { got.put(chains, steps, c, i, x); });
    drv.run(spec, steps - first,
            [&](const Ctx &x, std::size_t c, std::uint64_t i) { got.put(chains, steps, c, first + i, x); });

    bool same = want.v.size() == got.v.size() && std::memcmp(want.v.data(), got.v.data(), want.v.size() * sizeof(double)) == 0;
    unsigned long long acc = 0;
    for (std::size_t c = 0; c < chains; ++c) {
        same = same && seq[c].steps == spec[c].steps && seq[c].accepted == spec[c].accepted &&
               std::memcmp(&seq[c].scale, &spec[c].scale, sizeof(double)) == 0 &&
               seq[c].rng() == spec[c].rng() &&                       // the generators are in the same state
               seq[c].gauss(seq[c].rng) == spec[c].gauss(spec[c].rng);
        acc += seq[c].accepted;
    }
    const b9::SpecStats &s = drv.stats();
    std::printf("{\"mode\": %d, \"chains\": %zu, \"steps\": %llu, \"depth\": %zu, \"identical\": %s, "
                "\"acceptance\": %.4f, \"sequential_evals\": %llu, \"launches\": %llu, \"evaluated\": %llu, "
                "\"used\": %llu, \"undone\": %llu, \"largest_batch\": %zu, \"chain_steps_per_launch\": %.3f}\n",
                MODE, chains, (unsigned long long)steps, depth, same ? "true" : "false",
                (double)acc / (double)(chains * steps * (MODE == 2 ? 2 : 1)), seq_evals,
                (unsigned long long)s.launches, (unsigned long long)s.evaluated, (unsigned long long)s.used,
                (unsigned long long)s.undone, largest, (double)steps / (double)s.launches);
    return same && s.steps == chains * steps;
}

}  // namespace

int main() {
    bool ok = true;
    for (std::size_t depth : {1, 2, 5, 16, 64}) {
        ok &= run_case<0>(1, 3000, depth, 11);
        ok &= run_case<1>(1, 3000, depth, 12);
        ok &= run_case<2>(1, 3000, depth, 13);
    }
    ok &= run_case<0>(64, 400, 16, 21);     // many independent chains share each launch
    ok &= run_case<1>(64, 400, 16, 22);
    ok &= run_case<2>(64, 400, 8, 23);
    ok &= run_case<2>(3, 7, 64, 24);        // deeper than the run is long
    // a step that never asks for anything cannot be driven: that is reported, not looped on
    try {
        Ctx c;
        b9::SpeculativeDriver<Ctx> idle([](Ctx &x, const Eval &) { ++x.steps; },
                                        [](const double *, const std::uint32_t *, std::size_t, double *) {}, P, 4);
        idle.run(c, 10);
        ok &= c.steps == 10;               // ... although a step that needs no evaluation does complete
        std::printf("{\"idle_steps\": %llu}\n", c.steps);
    } catch (const std::logic_error &e) {
        std::printf("{\"idle_error\": \"%s\"}\n", e.what());
        ok = false;
    }
    return ok ? 0 : 1;
}
