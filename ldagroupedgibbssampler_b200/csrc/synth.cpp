// synth.cpp -- host-side synthetic corpus generator for benchmarks and tests; built into its own
// libldasynth.so (g++ only), so users of it -- bench.py's reference arm, the CPU tests -- map no GPU code.
//
// LDA generative model with fixed seeds: Phi_true (K_gen x V) ~ Dir(0.01 * V * m) where m is a
// Zipf(1.07) base measure over the vocabulary; per document a log-normal length clipped to
// [1, max_len], theta_d ~ Dir(0.1), tokens i.i.d. from the mixture; tokens of a document sorted
// by type id (the reference's bundled corpora are bags of words with equal types adjacent,
// src/main/resources/datasets/cats.txt).  Every document has its own counter-seeded generator, so
// the corpus does not depend on the number of host threads.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/ldasynth.h"

namespace {

struct Rng {   // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return ((double)(next() >> 11) + 0.5) * 0x1p-53; }
    double normal()
    {
        double u1 = uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    }
    double gamma(double a)
    {
        if (a < 1.0) return gamma(a + 1.0) * std::pow(uniform(), 1.0 / a);
        const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x, v;
            do { x = normal(); v = 1.0 + c * x; } while (v <= 0.0);
            v = v * v * v;
            double u = uniform();
            if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
            if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
        }
    }
};

inline uint64_t mix(uint64_t a, uint64_t b) { return Rng(a ^ (b * 0xD6E8FEB86659FD93ull)).next(); }

template <typename F> void parallel_for(int64_t n, F f)
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (n < 64) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=]() {
            int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
            for (int64_t i = lo; i < hi; ++i) f(i);
        });
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" int ldasynth_corpus(int64_t D, int64_t doc_first, int32_t V, int32_t K_gen, double mean_len,
                                   double sigma_len, int32_t max_len, uint64_t seed, int64_t *doc_offsets,
                                   int32_t *tokens, int64_t capacity, int64_t *n_tokens)
{
    if (D < 0 || doc_first < 0 || V < 1 || K_gen < 1 || !doc_offsets || !n_tokens || mean_len <= 0 || max_len < 1) return 1;
    // topic-word distributions as cumulative tables
    std::vector<double> base((size_t)V);
    double bs = 0.0;
    for (int32_t w = 0; w < V; ++w) { base[(size_t)w] = std::pow((double)w + 1.0, -1.07); bs += base[(size_t)w]; }
    std::vector<float> cdf((size_t)K_gen * V);
    parallel_for(K_gen, [&](int64_t k) {
        Rng r(mix(seed, 0x1000 + (uint64_t)k));
        std::vector<double> g((size_t)V);
        double s = 0.0;
        for (int32_t w = 0; w < V; ++w) { g[(size_t)w] = r.gamma(0.01 * V * base[(size_t)w] / bs); s += g[(size_t)w]; }
        double acc = 0.0;
        for (int32_t w = 0; w < V; ++w) { acc += g[(size_t)w] / s; cdf[(size_t)k * V + w] = (float)acc; }
        cdf[(size_t)k * V + V - 1] = 2.0f;   // sentinel: every uniform lands somewhere
    });
    // document lengths: log-normal with the requested mean
    const double mu = std::log(mean_len) - 0.5 * sigma_len * sigma_len;
    std::vector<int32_t> len((size_t)D);
    parallel_for(D, [&](int64_t d) {
        Rng r(mix(seed, 0x2000000 + (uint64_t)(doc_first + d)));
        double l = std::exp(mu + sigma_len * r.normal());
        int64_t li = (int64_t)std::llround(l);
        len[(size_t)d] = (int32_t)std::min<int64_t>(std::max<int64_t>(li, 1), max_len);
    });
    doc_offsets[0] = 0;
    for (int64_t d = 0; d < D; ++d) doc_offsets[d + 1] = doc_offsets[d] + len[(size_t)d];
    *n_tokens = doc_offsets[D];
    if (!tokens) return 0;   // sizing call
    if (doc_offsets[D] > capacity) return 2;
    parallel_for(D, [&](int64_t d) {
        Rng r(mix(seed, 0x4000000000ull + (uint64_t)(doc_first + d)));
        std::vector<double> th((size_t)K_gen);
        double s = 0.0;
        for (int k = 0; k < K_gen; ++k) { th[(size_t)k] = r.gamma(0.1); s += th[(size_t)k]; }
        double acc = 0.0;
        for (int k = 0; k < K_gen; ++k) { acc += th[(size_t)k] / s; th[(size_t)k] = acc; }
        th[(size_t)K_gen - 1] = 2.0;
        int32_t *out = tokens + doc_offsets[d];
        for (int32_t i = 0; i < len[(size_t)d]; ++i) {
            int k = (int)(std::lower_bound(th.begin(), th.end(), r.uniform()) - th.begin());
            const float *c = cdf.data() + (size_t)k * V;
            out[i] = (int32_t)(std::lower_bound(c, c + V, (float)r.uniform()) - c);
        }
        std::sort(out, out + len[(size_t)d]);
    });
    return 0;
}
