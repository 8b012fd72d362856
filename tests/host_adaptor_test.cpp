// Exercises sailfish_b200/host/sfb200_host.hpp the way mainQuantify uses the reference classes
// (src/SailfishQuantify.cpp:1322-1410) against mock ReadExperiment / Transcript / SailfishOpts types that carry the same
// members as the reference's (include/ReadExperiment.hpp, Transcript.hpp, SailfishOpts.hpp).
// Input  (argv[1]): text file: T, then T lines "len seq"; n reads then n lines "r1 r2"
// Output (stdout):  counters, class count, then T lines "estCount mass", then bootstrap / gibbs row sums.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../sailfish_b200/host/sfb200_host.hpp"

struct Transcript {                       // include/Transcript.hpp:55-69
    uint32_t RefLength = 0; double EffectiveLength = 0;
    void setEstCount(double v) { est_ = v; }
    void setMass(double v) { mass_ = v; }
    double estCount() const { return est_; }
    double mass() const { return mass_; }
private:
    double est_ = 0, mass_ = 0;
};
struct SailfishOpts {                     // include/SailfishOpts.hpp:9-41 (the members the path reads)
    bool useVBOpt = false, noEffectiveLengthCorrection = false;
    uint32_t numBootstraps = 0, numGibbsSamples = 0;
};
struct ReadExperiment {                   // include/ReadExperiment.hpp:65-97
    std::vector<Transcript> txps; uint64_t mapped = 0;
    std::vector<Transcript>& transcripts() { return txps; }
    uint64_t numMappedFragments() const { return mapped; }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    size_t T; in >> T;
    std::string seq; std::vector<uint64_t> off(T); std::vector<uint32_t> len(T);
    ReadExperiment exp; exp.txps.resize(T);
    for (size_t t = 0; t < T; ++t) {
        std::string s; double eff; in >> eff >> s;
        off[t] = seq.size(); len[t] = static_cast<uint32_t>(s.size()); seq += s;
        exp.txps[t].RefLength = len[t]; exp.txps[t].EffectiveLength = eff;
    }
    size_t n; in >> n;
    std::vector<std::string> r1(n), r2(n);
    for (size_t i = 0; i < n; ++i) in >> r1[i] >> r2[i];
    try {
        sfb200::Device dev(0);
        dev.buildIndex(seq, off, len, 31);
        sfb200::EquivalenceClassBuilder eqb(dev);
        sfb200_map_opts mo = {200, 1000, 10000, /*IU*/ 1 | (2 << 1) | (4 << 3), 0, 1, 0, 0, 0, 1000};
        eqb.start(mo);
        sfb200::GpuQuasiMapper mapper(dev, 3000);                       // small flush threshold: exercises several device batches
        for (size_t j0 = 0; j0 < n; j0 += 1000) {                       // parser jobs of 1000 reads (SailfishQuantify.cpp:73)
            const size_t m = std::min<size_t>(1000, n - j0);
            mapper.processReads(m, [&](size_t i) -> const std::string& { return r1[j0 + i]; }, [&](size_t i) -> const std::string& { return r2[j0 + i]; });
        }
        mapper.flush();
        eqb.finish();
        exp.mapped = eqb.numMappedFragments();
        std::printf("%llu %llu %llu %llu %llu %llu\n", (unsigned long long)eqb.numObservedFragments(), (unsigned long long)eqb.numMappedFragments(),
                    (unsigned long long)eqb.numFragHits(), (unsigned long long)eqb.upperBoundHits(), (unsigned long long)eqb.numFwd(), (unsigned long long)eqb.numRC());
        auto& vec = eqb.eqVec();
        uint64_t tot = 0;
        for (auto& kv : vec) tot += kv.second.count;
        std::printf("%zu %llu\n", vec.size(), (unsigned long long)tot);
        SailfishOpts sopt;
        sfb200::CollapsedEMOptimizer opt(dev);
        if (!opt.optimize(exp, sopt, 0.01, 10000)) { std::fprintf(stderr, "optimize failed: %s\n", opt.lastError().c_str()); return 1; }
        for (auto& t : exp.txps) std::printf("%.17g %.17g\n", t.estCount(), t.mass());
        sopt.numBootstraps = 3;
        std::function<bool(const std::vector<double>&)> wb = [&](const std::vector<double>& a) { double s = 0; for (double v : a) s += v; std::printf("boot %.6f\n", s); return true; };
        if (!opt.gatherBootstraps(exp, sopt, wb, 0.01, 10000)) return 1;
        sfb200::CollapsedGibbsSampler gs(dev);
        std::function<bool(const std::vector<int>&)> ws = [&](const std::vector<int>& a) { long long s = 0; for (int v : a) s += v; std::printf("gibbs %lld\n", s); return true; };
        if (!gs.sample(exp, sopt, ws, 3)) return 1;
    } catch (const sfb200::Error& e) {
        std::fprintf(stderr, "sfb200 error %d: %s\n", e.code, e.what());
        return 3;
    }
    return 0;
}
