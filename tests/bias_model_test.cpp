// CPU check: sfb200::BiasModel (sailfish_b200/host/sfb200_host.hpp) builds the fragment-length cdf table the way the reference's
// EmpiricalDistribution does -- compared bit for bit with the oracle's restatement (oracle/orc_empdist.hpp via orc_fld_cdf), which is
// pinned to the reference's own class (tests/test_oracle_bias.py).  Built and run by tests/test_host_adaptors.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../sailfish_b200/host/sfb200_host.hpp"

extern "C" uint32_t orc_fld_cdf(const uint32_t* fld_counts, uint32_t n_fld, float* cdf_out, uint32_t cap, uint32_t* max_value);

static int check(const std::vector<uint32_t>& fld, const char* what) {
    std::vector<uint32_t> rb(4096, 1), og(101, 1);
    sfb200::BiasModel m(true, rb, og, 10, 20, fld, 3);
    std::vector<float> want(fld.size() + 1);
    uint32_t mx = 0;
    const uint32_t n = orc_fld_cdf(fld.data(), (uint32_t)fld.size(), want.data(), (uint32_t)want.size(), &mx);
    const sfb200_bias_model* c = m.get();
    if (c->n_cdf != n || c->fld_max != mx || std::memcmp(c->fld_cdf, want.data(), n * sizeof(float)) != 0) {
        std::fprintf(stderr, "FAIL %s: n_cdf %u vs %u, fld_max %u vs %u\n", what, c->n_cdf, n, c->fld_max, mx);
        return 1;
    }
    if (c->mode != 2 || c->gc_samp != 3 || c->num_fwd != 10 || c->num_rc != 20 || c->read_bias[7] != 1 || c->observed_gc[100] != 1) { std::fprintf(stderr, "FAIL %s: fields\n", what); return 1; }
    return 0;
}

int main() {
    std::mt19937_64 rng(5);
    std::vector<uint32_t> g(1000), sparse(1000, 0), flat(300, 7), spike(1000, 0), noisy(1000);
    for (uint32_t x = 0; x < 1000; ++x) g[x] = (uint32_t)std::lround(30000.0 * std::exp(-0.5 * std::pow((x - 190.0) / 30.0, 2)));
    for (int k = 0; k < 40; ++k) sparse[rng() % 1000] += 1 + (uint32_t)(rng() % 50);
    spike[250] = 10000;
    for (auto& v : noisy) v = (uint32_t)(rng() % 1000);
    int bad = check(g, "normal") + check(sparse, "sparse") + check(flat, "flat") + check(spike, "spike") + check(noisy, "noisy");
    std::vector<uint32_t> one(1, 5);
    bad += check(one, "single bin");
    if (bad) return 1;
    std::printf("bias model ok\n");
    return 0;
}
