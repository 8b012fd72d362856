// sfb200_quant.cpp -- `sailfish quant`-shaped command line driver in C++ on top of libsfb200.
//
//   sfb200-quant [quant] -t transcripts.fa -l IU -1 r_1.fq[.gz] -2 r_2.fq[.gz] -o out_dir [options]
//   sfb200-quant [quant] -t transcripts.fa -l U  -r reads.fq -o out_dir
//   sfb200-quant index -t transcripts.fa -o index_dir [-k 31] [-f]        then        sfb200-quant quant -i index_dir ...
//   sfb200-quant genes -g genes.gtf|txp2gene.tsv -q out_dir/quant.sf      (what `quant -g` does after quantification)
//
// Option names follow the reference's `sailfish quant` (src/SailfishQuantify.cpp:1066-1150); the index is built on the GPU
// from the transcript sequences at start-up (a fraction of a second for 200k transcripts), so the `index` command (reference
// src/SailfishIndexer.cpp:66-237) only parses the FASTA once and stores names, lengths and sequence next to versionInfo.json
// (include/SailfishIndexVersionInfo.hpp:22-50) and header.json; RapMap's own on-disk format is not part of the reference tree.  The host side keeps the reference's structure: reader threads parse
// FASTA/FASTQ into batches (fastx_reader.hpp), the batches go through GpuQuasiMapper / EquivalenceClassBuilder /
// CollapsedEMOptimizer / CollapsedGibbsSampler (sfb200_host.hpp: the reference's class names over the C ABI), effective
// lengths are computed on the host as in quasiMapReads' tail (SailfishQuantify.cpp:648-838,937-992,1034-1043) and the writers
// reproduce GZipWriter's formats (src/GZipWriter.cpp:51-92,163-284).  There is no CPU fallback: without a CUDA device the
// program stops with an error.
#include <zlib.h>

#include <sys/stat.h>

#include <cctype>
#include <cerrno>
#include <cstring>
#include <stdexcept>
#include <unistd.h>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "fastx_reader.hpp"
#include "gene_agg.hpp"
#include "sfb200_host.hpp"

namespace {

// ---- library format (include/LibraryFormat.hpp:7-9,89-98; parseLibraryFormatStringNew, src/SailfishUtils.cpp:63-153) ------------
bool parse_library_format(std::string s, int32_t& id, bool& paired) {
    for (char& ch : s) ch = (char)toupper(ch);
    struct E { const char* n; int t, o, st; };
    // type: 0 single 1 paired; orientation: 0 same 1 away 2 toward 3 none; strandedness: 0 SA 1 AS 2 S 3 A 4 U
    static const E tab[] = {{"IU", 1, 2, 4}, {"ISF", 1, 2, 0}, {"ISR", 1, 2, 1}, {"OU", 1, 1, 4}, {"OSF", 1, 1, 0}, {"OSR", 1, 1, 1},
                            {"MU", 1, 0, 4}, {"MSF", 1, 0, 2}, {"MSR", 1, 0, 3}, {"U", 0, 3, 4}, {"SF", 0, 3, 2}, {"SR", 0, 3, 3}};
    for (const E& e : tab)
        if (s == e.n) { id = (e.t & 1) | ((e.o & 3) << 1) | ((e.st & 7) << 3); paired = e.t == 1; return true; }
    return false;
}

// ---- effective lengths (host, O(T + maxFragLen)) ---------------------------------------------------------------------------------
// getNormalFragLengthDist + correction factors of a discretised normal (SailfishQuantify.cpp:648-704)
std::vector<double> normal_correction_factors(uint32_t maxLen, double mean, double sd) {
    std::vector<double> cf(maxLen, 0.0);
    const double inv = 1.0 / sd;
    double cumMass = 0.0, cumDens = 0.0;
    for (uint32_t i = 0; i < maxLen; ++i) {
        const double x = inv * (static_cast<double>(i) - mean);
        const double dens = std::exp(-0.5 * x * x) * inv;
        cumMass += static_cast<double>(i) * dens;
        cumDens += dens;
        if (cumDens > 0) cf[i] = cumMass / cumDens;
    }
    return cf;
}
// getNormalFragLengthCounts (SailfishQuantify.cpp:675-704): the prior normal as rounded counts, what setFragLengthDist receives when
// too few fragment lengths were observed
std::vector<uint32_t> normal_frag_length_counts(uint32_t maxLen, int32_t totalCount, double mean, double sd) {
    std::vector<uint32_t> dist(maxLen, 0);
    const double inv = 1.0 / sd;
    auto kernel = [&](double p) { const double x = inv * (p - mean); return std::exp(-0.5 * x * x) * inv; };
    double totalMass = 0.0;
    for (uint32_t i = 0; i < maxLen; ++i) totalMass += kernel(static_cast<double>(i));
    if (totalMass > 0) for (uint32_t i = 0; i < maxLen; ++i) dist[i] = static_cast<uint32_t>(static_cast<int>(std::round(kernel(static_cast<double>(i)) * totalCount / totalMass)));
    return dist;
}
// correctionFactorsFromCounts (SailfishQuantify.cpp:769-807): running mean of the observed fragment lengths
std::vector<double> correction_factors_from_counts(const std::vector<uint32_t>& hist) {
    const size_t n = hist.size();
    std::vector<double> cf(n, 0.0);
    double acc = 0.0;
    uint32_t mult = n ? hist[0] : 0;
    for (size_t i = 1; i < n; ++i) {
        acc = static_cast<double>(static_cast<uint64_t>(hist[i]) * i) + acc;
        mult += hist[i];
        if (mult > 0) cf[i] = acc / static_cast<double>(mult);
    }
    return cf;
}
// computeSmoothedEffectiveLengths (:809-838) / setEffectiveLengthsDirect (:707-715) and the mode selection (:937-992, :1034-1043)
// computeEmpiricalEffectiveLengths (--unsmoothedFLD, :717-767): sum over fragment lengths of pdf(l) * (RefLength - l + 1), with the
// pdf EmpiricalDistribution builds from jointMap (every length 0 .. maxFragLen-1 with its count, :944-947): float, truncated where
// the cumulative mass passes 1 - 1e-6 (src/EmpiricalDistribution.cpp:29-94); transcripts not longer than the median keep RefLength
std::vector<float> empirical_pdf(const std::vector<uint32_t>& fld) {
    const size_t n = fld.size();
    double total = 0.0;
    for (uint32_t c : fld) total += c;
    size_t last = 0, maxval = 1;
    double cum = 0.0;
    for (; last < n; ++last) { cum += fld[last] / total; maxval = last; if (cum > 1.0 - 1e-6) break; }
    double kept = 0.0;
    for (size_t i = 0; i < last && i < n; ++i) kept += fld[i];
    std::vector<float> pdf(n ? maxval : 0);
    for (size_t v = 0; v < pdf.size(); ++v) pdf[v] = static_cast<float>(fld[v] / kept);
    return pdf;
}
std::vector<double> empirical_effective_lengths(const std::vector<uint32_t>& lens, const std::vector<uint32_t>& fld) {
    const size_t n = fld.size();
    std::vector<double> eff(lens.size());
    uint64_t observed = 0;
    for (uint32_t c : fld) observed += c;
    if (observed == 0) { for (size_t t = 0; t < lens.size(); ++t) eff[t] = lens[t]; return eff; }   // nothing observed: no distribution to correct with
    const std::vector<float> pdf = empirical_pdf(fld);
    size_t i = 0, j = n ? n - 1 : 0;                                          // the median by walking in from both ends (:83-93)
    unsigned u = n ? fld[0] : 0, v = n ? fld[n - 1] : 0;
    while (i < j) { if (u <= v) { v -= u; u = fld[++i]; } else { u -= v; v = fld[--j]; } }
    const float median = static_cast<float>(i);
    const uint32_t minVal = 0, maxVal = n ? static_cast<uint32_t>(n - 1) : 0;
    for (size_t t = 0; t < lens.size(); ++t) {
        const double refLen = lens[t];
        if (refLen <= median || !(maxVal > minVal)) { eff[t] = refLen; continue; }
        double e = 0.0;
        for (size_t l = minVal; l <= std::min(lens[t], maxVal); ++l) e += (l < pdf.size() ? pdf[l] : 0.0f) * (lens[t] - l + 1.0);
        eff[t] = e;
    }
    return eff;
}

std::vector<double> effective_lengths(const std::vector<uint32_t>& lens, const std::vector<uint32_t>& fld, uint32_t maxFragLen,
                                      int32_t numFragSamples, bool singleEnd, bool noCorrection, double priorMean, double priorSD,
                                      bool unsmoothed = false) {
    std::vector<double> eff(lens.size());
    if (noCorrection) { for (size_t i = 0; i < lens.size(); ++i) eff[i] = lens[i]; return eff; }
    uint64_t nSamp = 0;
    for (uint32_t c : fld) nSamp += c;
    const bool enough = !singleEnd && nSamp >= static_cast<uint64_t>(numFragSamples);
    if (enough && unsmoothed) return empirical_effective_lengths(lens, fld);  // :985-986
    const std::vector<double> cf = enough ? correction_factors_from_counts(fld) : normal_correction_factors(maxFragLen, priorMean, priorSD);
    for (size_t i = 0; i < lens.size(); ++i) {
        const uint32_t idx = std::min<uint32_t>(lens[i], maxFragLen - 1);
        const double e = static_cast<double>(lens[i]) - cf[idx] + 1.0;
        eff[i] = e < 1.0 ? static_cast<double>(lens[i]) : e;
    }
    return eff;
}

// ---- the experiment as the adaptors see it (Transcript / ReadExperiment members they touch) ------------------------------------
struct Transcript {
    std::string RefName;
    uint32_t RefLength = 0;
    double EffectiveLength = 0.0;
    double estCount_ = 0.0, mass_ = 0.0;
    void setEstCount(double v) { estCount_ = v; }
    void setMass(double v) { mass_ = v; }
    double mass() const { return mass_; }
    double estCount() const { return estCount_; }
};
struct ReadExperiment {
    std::vector<Transcript> txps;
    uint64_t numMapped = 0;
    std::vector<Transcript>& transcripts() { return txps; }
    uint64_t numMappedFragments() const { return numMapped; }
};
struct SailfishOpts {
    bool useVBOpt = false, noEffectiveLengthCorrection = false;
    uint32_t numBootstraps = 0, numGibbsSamples = 0;
    bool useUnsmoothedFLD = false;                                            // --unsmoothedFLD (SailfishQuantify.cpp:1109)
    bool biasCorrect = false, gcBiasCorrect = false;                          // SailfishQuantify.cpp:1089-1090
    int32_t numBiasSamples = 1000000;                                         // :1131
    uint32_t pdfSampFactor = 1;                                               // --gcSpeedSamp (:1103)
};

// mkdir -p; throws when a component cannot be created or is not a directory (the reference creates the output directory up front,
// SailfishQuantify.cpp:1320-1335, so an unwritable -o fails before any work is done)
void make_dir(const std::string& p) {
    if (p.empty()) return;
    for (size_t at = 1; at <= p.size(); ++at) {
        if (at != p.size() && p[at] != '/') continue;
        const std::string part = p.substr(0, at);
        if (mkdir(part.c_str(), 0755) != 0 && errno != EEXIST) throw std::runtime_error("cannot create directory " + part + ": " + strerror(errno));
    }
    struct stat st;
    if (stat(p.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) throw std::runtime_error(p + " is not a directory");
    if (access(p.c_str(), W_OK) != 0) throw std::runtime_error("directory " + p + " is not writable");
}

std::string json_escape(const std::string& v) {
    std::string o;
    for (char ch : v) { if (ch == '"' || ch == '\\') o.push_back('\\'); if ((unsigned char)ch >= 0x20) o.push_back(ch); }
    return o;
}
// cmd_info.json (SailfishQuantify.cpp:1262-1276): sf_version, then every option as given, under its long name; one value is a
// string, several (or none, for a switch) an array
void write_cmd_info(const std::string& path, int argc, char** argv) {
    static const char* const shortNames[][2] = {{"-t", "transcripts"}, {"-i", "index"}, {"-l", "libType"}, {"-o", "output"}, {"-r", "unmatedReads"},
                                                {"-1", "mates1"}, {"-2", "mates2"}, {"-p", "threads"}, {"-g", "geneMap"}, {"-k", "kmerLen"},
                                                {"-w", "maxReadOcc"}, {"-f", "force"}, {"-q", "quantFile"}};
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return;
    fprintf(f, "{\n    \"sf_version\": \"0.10.0-b200\"");
    for (int i = 1; i < argc; ++i) {
        const std::string o = argv[i];
        if (o.size() < 2 || o[0] != '-') continue;                             // the sub-command word
        std::string key = o.compare(0, 2, "--") == 0 ? o.substr(2) : o;
        for (const auto& sn : shortNames) if (o == sn[0]) key = sn[1];
        std::vector<std::string> vals;
        auto is_option = [](const char* t) {                                   // "-1" / "-2" are options, other "-<digit>..." tokens are numbers
            return t[0] == '-' && t[1] != '\0' && (!std::isdigit((unsigned char)t[1]) || ((t[1] == '1' || t[1] == '2') && t[2] == '\0'));
        };
        while (i + 1 < argc && !is_option(argv[i + 1])) vals.push_back(argv[++i]);
        fprintf(f, ",\n    \"%s\": ", json_escape(key).c_str());
        if (vals.size() == 1) fprintf(f, "\"%s\"", json_escape(vals[0]).c_str());
        else {
            fprintf(f, "[");
            for (size_t v = 0; v < vals.size(); ++v) fprintf(f, "%s\"%s\"", v ? ", " : "", json_escape(vals[v]).c_str());
            fprintf(f, "]");
        }
    }
    fprintf(f, "\n}\n");
    fclose(f);
}

// writeVectorToFile (src/GZipWriter.cpp:23-43): the raw elements, gzip level 6
template <typename T>
bool write_vector_gz(const std::string& path, const std::vector<T>& v) {
    gzFile f = gzopen(path.c_str(), "wb6");
    if (!f) return false;
    const bool ok = v.empty() || gzwrite(f, v.data(), (unsigned)(v.size() * sizeof(T))) > 0;
    gzclose(f);
    return ok;
}
// EmpiricalDistribution::realize (src/EmpiricalDistribution.cpp:127-144): 10000 draws from the pdf, as counts per length 0 .. maxValue.
// The reference seeds from std::random_device (aux/fld.gz differs from run to run); here the seed is fixed.
std::vector<int32_t> realize_fld(const std::vector<uint32_t>& fldCounts, uint32_t numSamp = 10000) {
    const std::vector<float> pdf = empirical_pdf(fldCounts);
    std::vector<double> padded(fldCounts.size(), 0.0);
    for (size_t i = 0; i < padded.size() && i < pdf.size(); ++i) padded[i] = pdf[i];
    std::vector<int32_t> samples(padded.size(), 0);
    double mass = 0.0;
    for (double x : padded) mass += x;
    if (padded.empty() || !(mass > 0.0)) return samples;                      // no observations at all (also catches NaN)
    std::mt19937 gen(0x5f3759df);
    std::discrete_distribution<int32_t> d(padded.begin(), padded.end());
    for (uint32_t i = 0; i < numSamp; ++i) ++samples[d(gen)];
    return samples;
}

std::string fmt_g(double x) { char b[64]; snprintf(b, sizeof b, "%g", x); return b; }   // cppformat's `{}` for doubles

// GZipWriter::writeAbundances (src/GZipWriter.cpp:194-248)
void write_quant_sf(const std::string& path, ReadExperiment& ex) {
    const double numMapped = static_cast<double>(ex.numMapped);
    double denom = 0.0;
    std::vector<double> tfrac(ex.txps.size(), 0.0);
    for (size_t i = 0; i < ex.txps.size(); ++i) {
        const Transcript& t = ex.txps[i];
        if (numMapped > 0 && t.EffectiveLength > 0) tfrac[i] = (t.estCount() / numMapped) / t.EffectiveLength;
        denom += tfrac[i];
    }
    FILE* f = fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    fprintf(f, "Name\tLength\tEffectiveLength\tTPM\tNumReads\n");
    for (size_t i = 0; i < ex.txps.size(); ++i) {
        const Transcript& t = ex.txps[i];
        const double tpm = denom > 0 ? tfrac[i] / denom * 1e6 : 0.0;
        fprintf(f, "%s\t%u\t%s\t%s\t%s\n", t.RefName.c_str(), t.RefLength, fmt_g(t.EffectiveLength).c_str(), fmt_g(tpm).c_str(),
                fmt_g(t.estCount()).c_str());
    }
    fclose(f);
}

// GZipWriter::writeEquivCounts (src/GZipWriter.cpp:51-92)
void write_eq_classes(const std::string& path, ReadExperiment& ex, sfb200::EquivalenceClassBuilder& eq) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    auto& vec = eq.eqVec();
    fprintf(f, "%zu\n%zu\n", ex.txps.size(), vec.size());
    for (const Transcript& t : ex.txps) fprintf(f, "%s\n", t.RefName.c_str());
    for (auto& kv : vec) {
        fprintf(f, "%zu\t", kv.first.txps.size());
        for (uint32_t t : kv.first.txps) fprintf(f, "%u\t", t);
        fprintf(f, "%llu\n", (unsigned long long)kv.second.count);
    }
    fclose(f);
}

// ---- index directory: versionInfo.json (same two fields as SailfishIndexVersionInfo), header.json, txpInfo.bin, seq.bin ---------
constexpr uint32_t kIndexVersion = 2;                                      // sailfish::indexVersion (include/SailfishConfig.hpp:33)

void write_index_dir(const std::string& dir, int k, const std::vector<std::string>& names, const std::string& seq,
                     const std::vector<uint32_t>& lens) {
    make_dir(dir);
    FILE* f = fopen((dir + "/versionInfo.json").c_str(), "w");
    if (!f) throw std::runtime_error("cannot write into " + dir);
    fprintf(f, "{\n    \"indexVersion\": %u,\n    \"kmerLength\": %d\n}\n", kIndexVersion, k);
    fclose(f);
    f = fopen((dir + "/header.json").c_str(), "w");
    fprintf(f, "{\n    \"IndexType\": \"sfb200-text\",\n    \"IndexVersion\": \"b200-1\",\n    \"k\": %d,\n    \"NumTranscripts\": %zu,\n"
               "    \"TextLength\": %zu,\n    \"note\": \"suffix array, k-mer table and presence filter are rebuilt on the GPU at load time\"\n}\n",
            k, names.size(), seq.size());
    fclose(f);
    f = fopen((dir + "/txpInfo.bin").c_str(), "wb");
    const uint64_t n = names.size();
    fwrite(&n, 8, 1, f);
    for (const std::string& s : names) { const uint32_t l = (uint32_t)s.size(); fwrite(&l, 4, 1, f); fwrite(s.data(), 1, l, f); }
    fwrite(lens.data(), 4, lens.size(), f);
    fclose(f);
    f = fopen((dir + "/seq.bin").c_str(), "wb");
    const uint64_t sl = seq.size();
    fwrite(&sl, 8, 1, f);
    if (fwrite(seq.data(), 1, seq.size(), f) != seq.size()) { fclose(f); throw std::runtime_error("short write to " + dir + "/seq.bin"); }
    fclose(f);
}

int read_index_dir(const std::string& dir, std::vector<std::string>& names, std::string& seq, std::vector<uint64_t>& off,
                   std::vector<uint32_t>& lens) {
    FILE* f = fopen((dir + "/versionInfo.json").c_str(), "r");
    if (!f) throw std::invalid_argument("Error: The index version file " + dir + "/versionInfo.json doesn't seem to exist.  Please try re-building the sailfish index.");
    unsigned ver = 0; int k = 0;
    char buf[256]; std::string txt;
    while (fgets(buf, sizeof buf, f)) txt += buf;
    fclose(f);
    const size_t pv = txt.find("\"indexVersion\""), pk = txt.find("\"kmerLength\"");
    if (pv == std::string::npos || pk == std::string::npos) throw std::runtime_error("malformed " + dir + "/versionInfo.json");
    ver = (unsigned)atoi(txt.c_str() + txt.find(':', pv) + 1); k = atoi(txt.c_str() + txt.find(':', pk) + 1);
    if (ver != kIndexVersion) throw std::runtime_error("index version " + std::to_string(ver) + " is not the version this program reads (" + std::to_string(kIndexVersion) + ")");
    f = fopen((dir + "/txpInfo.bin").c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + dir + "/txpInfo.bin");
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) { fclose(f); throw std::runtime_error("truncated txpInfo.bin"); }
    names.resize(n); lens.resize(n); off.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t l = 0;
        if (fread(&l, 4, 1, f) != 1) { fclose(f); throw std::runtime_error("truncated txpInfo.bin"); }
        names[i].resize(l);
        if (l && fread(&names[i][0], 1, l, f) != l) { fclose(f); throw std::runtime_error("truncated txpInfo.bin"); }
    }
    if (n && fread(lens.data(), 4, n, f) != n) { fclose(f); throw std::runtime_error("truncated txpInfo.bin"); }
    fclose(f);
    f = fopen((dir + "/seq.bin").c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + dir + "/seq.bin");
    uint64_t sl = 0;
    if (fread(&sl, 8, 1, f) != 1) { fclose(f); throw std::runtime_error("truncated seq.bin"); }
    seq.resize(sl);
    if (sl && fread(&seq[0], 1, sl, f) != sl) { fclose(f); throw std::runtime_error("truncated seq.bin"); }
    fclose(f);
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; ++i) { off[i] = o; o += lens[i]; }
    if (o != sl) throw std::runtime_error("index directory is inconsistent (sequence length)");
    return k;
}

struct Args {
    std::string transcripts, index, libType, out, auxDir = "aux", geneMap, aggKey = "gene_id", quantFile;
    bool indexCmd = false, genesCmd = false, efflensCmd = false, efflensSingle = false, force = false, saveDeviceIndex = false;
    std::string lensFile, fldFile;
    std::vector<std::string> unmated, mates1, mates2;
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    int k = 31, device = 0;
    bool dumpEq = false, parseOnly = false, noChecksum = false, deviceParse = false;
    size_t batch = 1u << 21, blockBytes = 0;             // 0: 2 MB of text per parser thread and block
    SailfishOpts sopt;
    sfb200_map_opts mopt;
    double fldMean = 200.0, fldSD = 80.0;
    bool discardOrphans = false;
};

[[noreturn]] void usage(const char* msg) {
    if (msg) fprintf(stderr, "error: %s\n\n", msg);
    fprintf(stderr,
            "sfb200-quant [quant] -t <transcripts.fa> -l <libType> {-r <reads> | -1 <mates1> -2 <mates2>} -o <dir> [options]\n"
            "  -t, --transcripts FILE     transcript FASTA (the index is built on the GPU from it), or\n"
            "  -i, --index DIR            a directory written by `sfb200-quant index -t FILE -o DIR [-k K] [-f]`\n"
            "  -l, --libType STR          IU ISF ISR OU OSF OSR MU MSF MSR U SF SR\n"
            "  -r, --unmatedReads FILE..  single-end reads (FASTA/FASTQ, plain or .gz)\n"
            "  -1, --mates1 FILE.. / -2, --mates2 FILE..\n"
            "  -o, --output DIR           quant.sf and <auxDir>/ are written here\n"
            "  -p, --threads N  -k, --kmerLen K (31)  --device N\n"
            "  --useVBOpt  --numBootstraps N  --numGibbsSamples N  --dumpEq  --noEffectiveLengthCorrection\n"
            "  --unsmoothedFLD            effective lengths from the observed fragment length distribution itself\n"
            "  --biasCorrect | --gcBiasCorrect  [--numBiasSamples N (1000000)]  [--gcSpeedSamp N (1)]\n"
            "  -g, --geneMap FILE         transcript-to-gene map (.gtf, or `transcript gene` per line): also write quant.genes.sf\n"
            "  --txpAggregationKey KEY    GTF attribute that names the gene (gene_id)\n"
            "  --maxFragLen N (1000)  --numFragSamples N (10000)  --fldMean M (200)  --fldSD S (80)  -w, --maxReadOcc N (200)\n"
            "  --strictIntersect  --ignoreLibCompat  --enforceLibCompat  --allowDovetail  --discardOrphans  --auxDir NAME\n"
            "  --deviceParse              send the read files' text to the GPU as it is and find the reads there (plain four-line FASTQ or two-line FASTA)\n"
            "  --parseOnly                only parse the read files and print record / base counts (no GPU needed)\n");
    exit(msg ? 2 : 0);
}

Args parse_args(int argc, char** argv) {
    Args a;
    a.mopt.max_read_occs = 200; a.mopt.max_frag_len = 1000; a.mopt.num_frag_samples = 10000; a.mopt.lib_format_id = 0;
    a.mopt.strict_intersect = 0; a.mopt.allow_orphans = 1; a.mopt.allow_dovetail = 0; a.mopt.ignore_compat = 0;
    a.mopt.enforce_compat = 0; a.mopt.max_interval = 1000;
    int i = 1;
    if (i < argc && std::string(argv[i]) == "quant") ++i;
    else if (i < argc && std::string(argv[i]) == "index") { a.indexCmd = true; ++i; }
    else if (i < argc && std::string(argv[i]) == "genes") { a.genesCmd = true; ++i; }
    else if (i < argc && std::string(argv[i]) == "efflens") { a.efflensCmd = true; ++i; }
    auto need = [&](const std::string& o) -> std::string { if (i + 1 >= argc) usage(("missing value for " + o).c_str()); return argv[++i]; };
    auto multi = [&](std::vector<std::string>& v) { while (i + 1 < argc && argv[i + 1][0] != '-') v.push_back(argv[++i]); };
    for (; i < argc; ++i) {
        const std::string o = argv[i];
        if (o == "-h" || o == "--help") usage(nullptr);
        else if (o == "-t" || o == "--transcripts") a.transcripts = need(o);
        else if (o == "-i" || o == "--index") a.index = need(o);
        else if (o == "-f" || o == "--force") a.force = true;
        else if (o == "--saveDeviceIndex") a.saveDeviceIndex = true;
        else if (o == "-g" || o == "--geneMap") a.geneMap = need(o);
        else if (o == "--txpAggregationKey") a.aggKey = need(o);
        else if (o == "-q" || o == "--quantFile") a.quantFile = need(o);
        else if (o == "--kmerSize") a.k = atoi(need(o).c_str());
        else if (o == "-l" || o == "--libType") a.libType = need(o);
        else if (o == "-o" || o == "--output") a.out = need(o);
        else if (o == "-r" || o == "--unmatedReads") multi(a.unmated);
        else if (o == "-1" || o == "--mates1") multi(a.mates1);
        else if (o == "-2" || o == "--mates2") multi(a.mates2);
        else if (o == "-p" || o == "--threads") a.threads = (unsigned)std::max(1, atoi(need(o).c_str()));
        else if (o == "-k" || o == "--kmerLen") a.k = atoi(need(o).c_str());
        else if (o == "--device") a.device = atoi(need(o).c_str());
        else if (o == "--useVBOpt") a.sopt.useVBOpt = true;
        else if (o == "--numBootstraps") a.sopt.numBootstraps = (uint32_t)atoi(need(o).c_str());
        else if (o == "--numGibbsSamples") a.sopt.numGibbsSamples = (uint32_t)atoi(need(o).c_str());
        else if (o == "--dumpEq") a.dumpEq = true;
        else if (o == "--noEffectiveLengthCorrection") a.sopt.noEffectiveLengthCorrection = true;
        else if (o == "--unsmoothedFLD") a.sopt.useUnsmoothedFLD = true;
        else if (o == "--lensFile") a.lensFile = need(o);                      // `efflens` only
        else if (o == "--fldFile") a.fldFile = need(o);
        else if (o == "--singleEnd") a.efflensSingle = true;
        else if (o == "--biasCorrect") a.sopt.biasCorrect = true;
        else if (o == "--gcBiasCorrect") a.sopt.gcBiasCorrect = true;
        else if (o == "--numBiasSamples") a.sopt.numBiasSamples = atoi(need(o).c_str());
        else if (o == "--gcSpeedSamp") a.sopt.pdfSampFactor = (uint32_t)std::max(1, atoi(need(o).c_str()));
        else if (o == "--maxFragLen") a.mopt.max_frag_len = (uint32_t)atoi(need(o).c_str());
        else if (o == "--numFragSamples") a.mopt.num_frag_samples = atoi(need(o).c_str());
        else if (o == "--fldMean") a.fldMean = atof(need(o).c_str());
        else if (o == "--fldSD") a.fldSD = atof(need(o).c_str());
        else if (o == "-w" || o == "--maxReadOcc") a.mopt.max_read_occs = (uint32_t)atoi(need(o).c_str());
        else if (o == "--strictIntersect") a.mopt.strict_intersect = 1;
        else if (o == "--ignoreLibCompat") a.mopt.ignore_compat = 1;
        else if (o == "--enforceLibCompat") a.mopt.enforce_compat = 1;
        else if (o == "--allowDovetail") a.mopt.allow_dovetail = 1;
        else if (o == "--discardOrphans") a.discardOrphans = true;
        else if (o == "--auxDir") a.auxDir = need(o);
        else if (o == "--batchReads") a.batch = (size_t)std::max(1, atoi(need(o).c_str()));
        else if (o == "--blockBytes") a.blockBytes = (size_t)std::max(0, atoi(need(o).c_str()));    // parser block size (tests); 0 = default
        else if (o == "--parseOnly") a.parseOnly = true;
        else if (o == "--deviceParse") a.deviceParse = true;                   // FASTQ text is parsed on the GPU (sfb200_map_fastq)
        else if (o == "--noChecksum") a.noChecksum = true;                     // with --parseOnly: count only (parser throughput)
        else usage(("unknown option " + o).c_str());
    }
    if (a.discardOrphans) a.mopt.allow_orphans = 0;                       // SailfishQuantify.cpp:1204
    return a;
}

// ---- read ingestion on the device (--deviceParse): the host only reads the files; sfb200_map_fastq finds the records -----------------
// One mate's text, the files one after another (records do not span files; a file whose last line lacks its newline gets one).
// A reader thread fills page-locked blocks ahead of the device (H2D from page-locked memory runs at the PCIe rate, and the next
// block is read while the current one is extracted and mapped).  Every block has `head` free bytes in front of its data: the
// part of the previous block that sfb200_map_fastq did not consume (an incomplete record, or records the other mate had no
// partner for yet) is copied there, so the device always sees one contiguous text that starts at a record boundary.
class RawTextStream {
public:
    struct Block { char* mem = nullptr; size_t cap = 0, head = 0, len = 0; };
    RawTextStream(const std::vector<std::string>& files, size_t block, unsigned threads)
        : files_(files), block_(block), threads_(threads ? threads : 1), th_(&RawTextStream::produce, this) {}
    ~RawTextStream() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
        for (Block* b : all_) { if (b->mem) sfb200_host_free(b->mem); delete b; }
    }
    RawTextStream(const RawTextStream&) = delete;
    RawTextStream& operator=(const RawTextStream&) = delete;

    const char* data() const { return cur_ ? cur_->mem + start_ : nullptr; }
    size_t size() const { return len_; }
    bool exhausted() const { return drained_; }
    void consume(size_t n) { start_ += n; len_ -= n; }
    bool only_whitespace() const { for (size_t i = 0; i < len_; ++i) { const char ch = data()[i]; if (ch != '\n' && ch != '\r' && ch != ' ') return false; } return true; }
    // make more text available behind what is left: at least `min_len` bytes unless the files end first; false = nothing at all is left
    bool fill(size_t min_len) {
        while (len_ < min_len && !drained_) {
            Block* nb = pop();
            if (!nb) { drained_ = true; break; }
            if (len_ > nb->head) nb = grow(nb, len_);                          // the carried text does not fit in front: rare, a bigger block
            if (len_) std::memcpy(nb->mem + nb->head - len_, cur_->mem + start_, len_);
            if (cur_) recycle(cur_);
            cur_ = nb; start_ = nb->head - len_; len_ += nb->len;
        }
        return len_ > 0;
    }

private:
    Block* make(size_t head, size_t data_cap) {
        Block* b = new Block();
        b->cap = head + data_cap + files_.size() + 2;                          // a newline may be appended per file
        b->mem = static_cast<char*>(sfb200_host_alloc(b->cap));
        if (!b->mem) { delete b; throw std::runtime_error("cannot allocate page-locked host memory for the FASTQ text"); }
        b->head = head; b->len = 0;
        std::lock_guard<std::mutex> lk(mu_);
        all_.push_back(b);
        return b;
    }
    Block* grow(Block* nb, size_t need_head) {
        Block* g = make(need_head, nb->len);
        std::memcpy(g->mem + g->head, nb->mem + nb->head, nb->len);
        g->len = nb->len;
        recycle(nb);
        return g;
    }
    Block* pop() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !ready_.empty() || eof_ || !err_.empty(); });
        if (!err_.empty()) throw std::runtime_error(err_);
        if (ready_.empty()) return nullptr;
        Block* b = ready_.front();
        ready_.erase(ready_.begin());
        cv_.notify_all();
        return b;
    }
    void recycle(Block* b) {
        std::lock_guard<std::mutex> lk(mu_);
        if (b->head == block_ && free_.size() < 3) free_.push_back(b);         // odd-sized (grown) blocks are simply kept until the end
        cv_.notify_all();
    }
    void produce() {
        try {
            if (files_.empty()) { std::lock_guard<std::mutex> lk(mu_); eof_ = true; cv_.notify_all(); return; }
            int fd = -1;
            size_t next = 0;
            uint64_t pos = 0, fsize = 0;
            char last_char = '\n';
            for (;;) {
                Block* b = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return stop_ || ready_.size() < 2; });
                    if (stop_) break;
                    if (!free_.empty()) { b = free_.back(); free_.pop_back(); }
                }
                if (!b) b = make(block_, block_);
                b->len = 0;
                bool more = true;
                while (b->len < block_) {
                    if (fd < 0) {
                        if (next >= files_.size()) { more = false; break; }
                        fd = ::open(files_[next].c_str(), O_RDONLY);
                        if (fd < 0) throw std::runtime_error("cannot open " + files_[next]);
                        unsigned char magic[2] = {0, 0};
                        if (::pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b)
                            throw std::runtime_error(files_[next] + ": --deviceParse reads plain FASTQ text; inflate gzipped files first (or drop the option)");
                        struct stat st;
                        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) throw std::runtime_error(files_[next] + ": --deviceParse needs a regular file");
                        fsize = (uint64_t)st.st_size; pos = 0;
                    }
                    const size_t room = (size_t)std::min<uint64_t>(block_ - b->len, fsize - pos);
                    char* dst = b->mem + b->head + b->len;
                    read_exact(fd, dst, room, pos, files_[next]);
                    if (room) last_char = dst[room - 1];
                    b->len += room; pos += room;
                    if (pos == fsize) {                                        // end of this file
                        if (fsize > 0 && last_char != '\n') { b->mem[b->head + b->len++] = '\n'; last_char = '\n'; }
                        // blank lines at the end of a file (the host parser skips them) would shift every record of the next file
                        char* m0 = b->mem + b->head;
                        while (b->len >= 2 && m0[b->len - 1] == '\n' &&
                               (m0[b->len - 2] == '\n' || (b->len >= 3 && m0[b->len - 2] == '\r' && m0[b->len - 3] == '\n')))
                            b->len -= (m0[b->len - 2] == '\r') ? 2 : 1;
                        ::close(fd); fd = -1; ++next;
                    }
                }
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    if (b->len) ready_.push_back(b); else free_.push_back(b);
                    if (!more) eof_ = true;
                }
                cv_.notify_all();
                if (!more) break;
            }
            if (fd >= 0) ::close(fd);
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(mu_);
            err_ = e.what(); eof_ = true;
            cv_.notify_all();
        }
    }
    // large reads are split over a few threads (one pread stream does not saturate a page cache, let alone an NVMe array)
    void read_exact(int fd, char* dst, size_t n, uint64_t pos, const std::string& name) {
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, n >> 24));      // at least 16 MB per thread
        std::vector<std::thread> th;
        std::vector<int> bad(nt, 0);
        auto part = [&](unsigned t) {
            size_t a = n * t / nt; const size_t b = n * (t + 1) / nt;
            while (a < b) {
                const ssize_t r = ::pread(fd, dst + a, b - a, (off_t)(pos + a));
                if (r <= 0) { bad[t] = 1; return; }
                a += (size_t)r;
            }
        };
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(part, t);
        part(0);
        for (std::thread& x : th) x.join();
        for (int b : bad) if (b) throw std::runtime_error("read error in " + name);
    }

    std::vector<std::string> files_;
    size_t block_;
    unsigned threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Block*> ready_, free_, all_;
    bool eof_ = false, stop_ = false;
    std::string err_;
    // consumer side
    Block* cur_ = nullptr;
    size_t start_ = 0, len_ = 0;
    bool drained_ = false;
    std::thread th_;                                                           // last member: starts when everything else exists
};

// all reads of a library through sfb200_map_fastq; returns the number of fragments
uint64_t map_fastq_files(sfb200::Device& dev, const std::vector<std::string>& f1, const std::vector<std::string>& f2, size_t blockBytes, unsigned threads) {
    const bool paired = !f2.empty();
    const size_t block = blockBytes ? blockBytes : (size_t)64 << 20;          // x 2 (room in front) x up to 4 blocks per mate, page-locked
    RawTextStream s1(f1, block, threads), s2(f2, block, threads);
    size_t want = block;                                                       // a mate that is ahead (text left over) takes no new block
    uint64_t total = 0;
    for (;;) {
        const bool have1 = s1.fill(want), have2 = paired ? s2.fill(want) : false;
        if (!have1 && !have2) break;
        uint64_t n = 0, c1 = 0, c2 = 0;
        if (have1 && (!paired || have2))
            dev.check(sfb200_map_fastq(dev.get(), s1.data(), s1.size(), paired ? s2.data() : nullptr, paired ? s2.size() : 0, 0, &n, &c1, paired ? &c2 : nullptr));
        if (n == 0) {
            if (paired && ((s1.exhausted() && s1.only_whitespace()) != (s2.exhausted() && s2.only_whitespace())) &&
                (s1.exhausted() || s2.exhausted()) && (s1.only_whitespace() != s2.only_whitespace()))
                throw std::runtime_error("mate files hold different numbers of reads");
            if (s1.exhausted() && (!paired || s2.exhausted())) {               // no complete record is left
                if (!s1.only_whitespace() || (paired && !s2.only_whitespace()))
                    throw std::runtime_error(paired && s1.only_whitespace() != s2.only_whitespace() ? "mate files hold different numbers of reads" : "truncated record at end of file");
                break;
            }
            want += block;                                                     // a record longer than the block (or one mate far behind)
            continue;
        }
        want = block;
        total += n;
        s1.consume((size_t)c1);
        if (paired) s2.consume((size_t)c2);
    }
    return total;
}

// do all n reads of a batch have one length?  (off[0] == 0: ReadBatch::clear)
static bool one_length(const uint64_t* off, size_t n, uint32_t* len) {
    if (n == 0 || off[0] != 0 || off[1] > 0xFFFFFFFFull) return false;
    const uint64_t L = off[1];
    uint64_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad |= (off[i + 1] - off[i]) ^ L;
    *len = (uint32_t)L;
    return bad == 0;
}

// ---- read ingestion: a producer thread parses the next batch while the current one is copied to the device and mapped ------------
struct PairBatch { sfb200::ReadBatch m1, m2; bool last = false; };

class BatchPipe {
public:
    BatchPipe(const std::vector<std::string>& f1, const std::vector<std::string>& f2, size_t batch, unsigned threads, size_t block)
        : files1_(f1), files2_(f2), batch_(batch), block_(block), threads_(threads), th_(&BatchPipe::produce, this) {}
    // a consumer that stops early (device error while mapping) must not leave the producer waiting for room in the queue
    ~BatchPipe() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }
    // blocks until a batch is ready; returns nullptr after the last one
    std::unique_ptr<PairBatch> pop() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || done_; });
        if (!err_.empty()) throw std::runtime_error(err_);
        if (q_.empty()) return nullptr;
        std::unique_ptr<PairBatch> b = std::move(q_.front());
        q_.erase(q_.begin());
        cv_.notify_all();
        return b;
    }
    // hand a consumed batch back: its buffers are reused for a later batch (no allocation, no page faults after warm-up)
    void recycle(std::unique_ptr<PairBatch> b) {
        std::lock_guard<std::mutex> lk(mu_);
        if (free_.size() < 4) free_.push_back(std::move(b));
    }

private:
    void produce() {
        try {
            const bool paired = !files2_.empty();
            for (size_t fi = 0; fi < files1_.size(); ++fi) {
                const unsigned t1 = paired ? std::max(1u, threads_ / 2) : threads_;
                sfb200::FastxReader r1(files1_[fi], t1, block_);
                struct Report { sfb200::FastxReader& r; ~Report() { if (getenv("SFB200_PARSE_TIMING")) { const double* t = r.phase_seconds();
                    fprintf(stderr, "[sfb200-quant] parser phases (s): read %.3f  index %.3f  lengths+offsets %.3f  copy %.3f  tail %.3f\n", t[0], t[1], t[2], t[3], t[4]); } } } report{r1};
                std::unique_ptr<sfb200::FastxReader> r2;
                if (paired) r2.reset(new sfb200::FastxReader(files2_[fi], t1, block_));
                for (;;) {
                    std::unique_ptr<PairBatch> b;
                    { std::lock_guard<std::mutex> lk(mu_); if (!free_.empty()) { b = std::move(free_.back()); free_.pop_back(); } }
                    if (!b) b.reset(new PairBatch());
                    b->m1.clear(); b->m2.clear();
                    size_t n1 = 0, n2 = 0;
                    if (paired) {                                             // the two mates are parsed side by side
                        std::string err2;                                       // an exception must not leave the helper thread
                        std::thread other([&] { try { n2 = r2->next(b->m2, batch_); } catch (const std::exception& e) { err2 = e.what(); } });
                        try { n1 = r1.next(b->m1, batch_); } catch (...) { other.join(); throw; }
                        other.join();
                        if (!err2.empty()) throw std::runtime_error(err2);
                        if (n1 != n2) throw std::runtime_error("mate files " + files1_[fi] + " / " + files2_[fi] + " hold different numbers of reads");
                    } else {
                        n1 = r1.next(b->m1, batch_);
                    }
                    if (n1 == 0) break;
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return stop_ || q_.size() < 2; });
                    if (stop_) return;
                    q_.push_back(std::move(b));
                    cv_.notify_all();
                }
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(mu_);
            err_ = e.what();
        }
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
        cv_.notify_all();
    }
    std::vector<std::string> files1_, files2_;
    size_t batch_, block_;
    unsigned threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::unique_ptr<PairBatch>> q_, free_;
    bool done_ = false, stop_ = false;
    std::string err_;
    std::thread th_;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char** argv) {
    try {
        Args a = parse_args(argc, argv);
        auto gene_level = [&](const std::string& quantPath) {                  // generateGeneLevelEstimates (SailfishUtils.cpp:1043-1088)
            size_t nt = 0, ng = 0, nu = 0;
            fprintf(stderr, "Computing gene-level abundance estimates\n");
            const std::string outp = sfb200::generate_gene_level_estimates(a.geneMap, quantPath, a.aggKey, &nt, &ng, &nu);
            fprintf(stderr, "There were %zu transcripts mapping to %zu genes\n", nt, ng);
            if (nu) fprintf(stderr, "WARNING: %zu transcripts of %s are not in the map; each is reported as its own gene\n", nu, quantPath.c_str());
            fprintf(stderr, "[sfb200-quant] wrote %s\n", outp.c_str());
        };
        if (a.efflensCmd) {
            // test hook, no GPU: the effective lengths `quant` would compute for transcript lengths (--lensFile, one per line) and a
            // fragment length histogram (--fldFile, counts for 0 .. maxFragLen-1), printed with 17 significant digits
            if (a.lensFile.empty() || a.fldFile.empty()) usage("efflens needs --lensFile and --fldFile");
            std::vector<uint32_t> lens, fld;
            { std::ifstream f(a.lensFile); for (uint64_t x; f >> x;) lens.push_back((uint32_t)x); }
            { std::ifstream f(a.fldFile); for (uint64_t x; f >> x;) fld.push_back((uint32_t)x); }
            if (fld.size() != a.mopt.max_frag_len) usage("--fldFile must hold --maxFragLen counts");
            const std::vector<double> eff = effective_lengths(lens, fld, a.mopt.max_frag_len, a.mopt.num_frag_samples, a.efflensSingle,
                                                              a.sopt.noEffectiveLengthCorrection, a.fldMean, a.fldSD, a.sopt.useUnsmoothedFLD);
            for (double e : eff) printf("%.17g\n", e);
            if (!a.out.empty()) {                                              // and aux/fld.gz as `quant` would write it
                uint64_t nSamp = 0;
                for (uint32_t c : fld) nSamp += c;
                const bool enough = !a.efflensSingle && nSamp >= static_cast<uint64_t>(a.mopt.num_frag_samples);
                make_dir(a.out);
                if (!write_vector_gz(a.out + "/fld.gz", realize_fld(enough ? fld : normal_frag_length_counts(a.mopt.max_frag_len, a.mopt.num_frag_samples, a.fldMean, a.fldSD)))) return 1;
            }
            return 0;
        }
        if (a.genesCmd) {
            if (a.geneMap.empty() || a.quantFile.empty()) usage("genes needs -g <map> and -q <quant.sf>");
            gene_level(a.quantFile);
            return 0;
        }
        if (a.indexCmd) {                                                     // `sailfish index` (src/SailfishIndexer.cpp:66-237); no GPU needed
            if (a.transcripts.empty() || a.out.empty()) usage("index needs -t <transcripts.fa> and -o <dir>");
            if (a.k % 2 == 0 || a.k > 31 || a.k < 3) {
                fprintf(stderr, "k-mer length should be odd to avoid a k-mer being it's own reverse complement\nplease specify an odd value of k (<= 31)\n");
                return 1;
            }
            struct stat st;
            if (!a.force && stat((a.out + "/header.json").c_str(), &st) == 0) {
                fprintf(stderr, "All index files seem up-to-date.\nTo force a rebuild of the index, use the --force option.\n");
                return 0;
            }
            std::string seq; std::vector<std::string> names; std::vector<uint64_t> off; std::vector<uint32_t> lens;
            sfb200::read_transcripts(a.transcripts, names, seq, off, lens);
            if (names.empty()) throw std::runtime_error("no transcripts in " + a.transcripts);
            write_index_dir(a.out, a.k, names, seq, lens);
            fprintf(stderr, "[sfb200-quant] index: %zu transcripts, %.1f Mnt, k = %d -> %s\n", names.size(), seq.size() / 1e6, a.k, a.out.c_str());
            if (a.saveDeviceIndex) {                                           // the device structures as well (needs the GPU): quant -i loads them
                std::vector<uint64_t> off(lens.size());
                uint64_t at = 0;
                for (size_t i = 0; i < lens.size(); ++i) { off[i] = at; at += lens[i]; }
                sfb200::Device dev(a.device);
                dev.buildIndex(seq, off, lens, a.k);
                dev.saveIndex(a.out + "/device_index.bin");
                fprintf(stderr, "[sfb200-quant] index: device structures written to %s/device_index.bin\n", a.out.c_str());
            }
            return 0;
        }
        const bool paired_files = !a.mates1.empty() || !a.mates2.empty();
        const std::vector<std::string>& f1 = paired_files ? a.mates1 : a.unmated;
        if (f1.empty()) usage("no read files given");
        if (paired_files && a.mates1.size() != a.mates2.size()) usage("--mates1 and --mates2 need the same number of files");
        if (!a.parseOnly && !a.out.empty()) { make_dir(a.out); make_dir(a.out + "/" + a.auxDir); }   // fail fast, before the index is built
        if (a.parseOnly) {
            BatchPipe pipe(f1, a.mates2, a.batch, a.threads, a.blockBytes);
            // FNV-1a of the mate-1 bases, the mate-2 bases and the read lengths of either mate (independent of the batching)
            uint64_t n = 0, bases1 = 0, bases2 = 0, x[4] = {1469598103934665603ULL, 1469598103934665603ULL, 1469598103934665603ULL, 1469598103934665603ULL};
            auto mix = [&](int k, uint64_t v) { x[k] ^= v; x[k] *= 1099511628211ULL; };
            while (std::unique_ptr<PairBatch> b = pipe.pop()) {
                n += b->m1.size(); bases1 += b->m1.bases.size(); bases2 += b->m2.bases.size();
                if (a.noChecksum) { pipe.recycle(std::move(b)); continue; }
                for (char ch : b->m1.bases) mix(0, (unsigned char)ch);
                for (char ch : b->m2.bases) mix(1, (unsigned char)ch);
                for (size_t i = 0; i + 1 < b->m1.off.size(); ++i) mix(2, (b->m1.off[i + 1] - b->m1.off[i]) & 0xFFFF);
                for (size_t i = 0; i + 1 < b->m2.off.size(); ++i) mix(3, (b->m2.off[i + 1] - b->m2.off[i]) & 0xFFFF);
                pipe.recycle(std::move(b));
            }
            printf("{\"records\": %llu, \"bases1\": %llu, \"bases2\": %llu, \"fnv1a\": [\"%016llx\", \"%016llx\", \"%016llx\", \"%016llx\"]}\n",
                   (unsigned long long)n, (unsigned long long)bases1, (unsigned long long)bases2, (unsigned long long)x[0],
                   (unsigned long long)x[1], (unsigned long long)x[2], (unsigned long long)x[3]);
            return 0;
        }
        if ((a.transcripts.empty() == a.index.empty()) || a.libType.empty() || a.out.empty()) usage("one of -t / -i, and -l and -o are required");
        if (!a.geneMap.empty()) {                                              // verified before any real work (SailfishQuantify.cpp:1208-1218)
            struct stat gst;
            if (stat(a.geneMap.c_str(), &gst) != 0) {
                fprintf(stderr, "Could not find transcript <=> gene map file %s\nExiting now: please either omit the 'geneMap' option or provide a valid file\n", a.geneMap.c_str());
                return 1;
            }
        }
        if (a.sopt.biasCorrect && a.sopt.gcBiasCorrect) {                     // SailfishQuantify.cpp:1293-1297
            fprintf(stderr, "Enabling both sequence-specific and fragment GC bias correction simultaneously is not yet supported. Please disable one of these options.\n");
            return 1;
        }
        if (a.sopt.gcBiasCorrect && !paired_files) {                          // :1298-1309
            fprintf(stderr, "[sfb200-quant] Fragment GC bias correction is currently only implemented for paired-end libraries. It is being disabled\n");
            a.sopt.gcBiasCorrect = false;
        }
        const bool doBias = a.sopt.biasCorrect || a.sopt.gcBiasCorrect;
        if (doBias && a.sopt.noEffectiveLengthCorrection) usage("--biasCorrect / --gcBiasCorrect need the effective length correction (the reference has no fragment length distribution without it)");
        if (a.sopt.numBootstraps && a.sopt.numGibbsSamples) usage("--numBootstraps and --numGibbsSamples are mutually exclusive (SailfishQuantify.cpp:1281-1287)");
        bool lib_paired = false;
        if (!parse_library_format(a.libType, a.mopt.lib_format_id, lib_paired)) usage(("unknown library type " + a.libType).c_str());
        if (lib_paired != paired_files) usage("the library type does not match the read files given");

        const double t_start = now_s();
        std::string run_start;                                                // SailfishQuantify.cpp:1204-1206
        { const std::time_t now = std::time(nullptr); run_start = std::asctime(std::localtime(&now)); if (!run_start.empty()) run_start.pop_back(); }
        // ---- transcripts + index (what SailfishIndex::load + ReadExperiment's constructor do, ReadExperiment.hpp:59-131)
        ReadExperiment ex;
        std::string seq; std::vector<std::string> names; std::vector<uint64_t> off; std::vector<uint32_t> lens;
        if (!a.index.empty()) a.k = read_index_dir(a.index, names, seq, off, lens);
        else sfb200::read_transcripts(a.transcripts, names, seq, off, lens);
        if (names.empty()) throw std::runtime_error("no transcripts in " + (a.index.empty() ? a.transcripts : a.index));
        ex.txps.resize(names.size());
        for (size_t i = 0; i < names.size(); ++i) { ex.txps[i].RefName = names[i]; ex.txps[i].RefLength = lens[i]; }
        sfb200::Device dev(a.device, /*bindHost=*/true);         // parser / reader threads start later and inherit the CPU set
        struct stat ist;
        if (!a.index.empty() && stat((a.index + "/device_index.bin").c_str(), &ist) == 0) dev.loadIndex(a.index + "/device_index.bin", (uint32_t)lens.size());
        else dev.buildIndex(seq, off, lens, a.k);
        const double t_index = now_s();
        fprintf(stderr, "[sfb200-quant] %zu transcripts, %.1f Mnt, index built in %.2f s\n", names.size(), seq.size() / 1e6, t_index - t_start);

        // ---- quasi-mapping -> equivalence classes (quasiMapReads, SailfishQuantify.cpp:864-1047)
        sfb200::EquivalenceClassBuilder eqBuilder(dev);
        eqBuilder.start(a.mopt);
        if (doBias) eqBuilder.collectBias(a.sopt.biasCorrect, a.sopt.gcBiasCorrect, a.sopt.numBiasSamples);
        if (a.deviceParse) {
            map_fastq_files(dev, f1, a.mates2, a.blockBytes, a.threads);
        } else {
            BatchPipe pipe(f1, a.mates2, a.batch, a.threads, a.blockBytes);
            while (std::unique_ptr<PairBatch> b = pipe.pop()) {
                const size_t n = b->m1.size();
                b->m1.bases.push_back('\0');
                // reads of one length (the usual case): no offsets over the host link (sfb200_map_batch_fixed)
                uint32_t len1 = 0, len2 = 0;
                const bool fixed = one_length(b->m1.off.data(), n, &len1) && (!paired_files || one_length(b->m2.off.data(), n, &len2));
                if (paired_files) {
                    b->m2.bases.push_back('\0');
                    if (fixed) dev.check(sfb200_map_batch_fixed(dev.get(), b->m1.bases.data(), len1, b->m2.bases.data(), len2, n));
                    else dev.check(sfb200_map_batch(dev.get(), b->m1.bases.data(), b->m1.off.data(), b->m2.bases.data(), b->m2.off.data(), n));
                } else {
                    if (fixed) dev.check(sfb200_map_batch_fixed(dev.get(), b->m1.bases.data(), len1, nullptr, 0, n));
                    else dev.check(sfb200_map_batch(dev.get(), b->m1.bases.data(), b->m1.off.data(), nullptr, nullptr, n));
                }
                pipe.recycle(std::move(b));                            // sfb200_map_batch returns when the host buffers are free again
            }
        }
        eqBuilder.finish();
        ex.numMapped = eqBuilder.numMappedFragments();
        const double t_map = now_s();
        fprintf(stderr, "[sfb200-quant] %llu fragments, %llu mapped (%.2f%%), %llu equivalence classes, %.2f s\n",
                (unsigned long long)eqBuilder.numObservedFragments(), (unsigned long long)ex.numMapped,
                100.0 * ex.numMapped / std::max<uint64_t>(1, eqBuilder.numObservedFragments()), (unsigned long long)eqBuilder.numClasses(),
                t_map - t_index);
        if (const uint64_t nclip = sfb200_map_clipped(dev.get()))
            fprintf(stderr, "[sfb200-quant] WARNING: %llu mates are longer than 256 bases and were mapped by their first 256 (the reference maps the whole read)\n",
                    (unsigned long long)nclip);

        // ---- effective lengths, inference
        const std::vector<double> eff = effective_lengths(lens, eqBuilder.fragLengthCounts(), a.mopt.max_frag_len, a.mopt.num_frag_samples,
                                                          !paired_files, a.sopt.noEffectiveLengthCorrection, a.fldMean, a.fldSD, a.sopt.useUnsmoothedFLD);
        for (size_t i = 0; i < eff.size(); ++i) ex.txps[i].EffectiveLength = eff[i];
        make_dir(a.out);
        write_cmd_info(a.out + "/cmd_info.json", argc, argv);
        const std::string aux = a.out + "/" + a.auxDir;
        make_dir(aux);
        if (a.dumpEq) write_eq_classes(aux + "/eq_classes.txt", ex, eqBuilder);
        sfb200::CollapsedEMOptimizer optimizer(dev);
        // readExp.setFragLengthDist (:966-984, :1039): the observed histogram when enough fragments were sampled, else the rounded normal
        uint64_t nSamp = 0;
        for (uint32_t c : eqBuilder.fragLengthCounts()) nSamp += c;
        const bool enough = paired_files && nSamp >= static_cast<uint64_t>(a.mopt.num_frag_samples);
        const std::vector<uint32_t> fldCounts = enough ? eqBuilder.fragLengthCounts()
                                                       : normal_frag_length_counts(a.mopt.max_frag_len, a.mopt.num_frag_samples, a.fldMean, a.fldSD);
        bool opt_ok;
        if (doBias) {
            sfb200::BiasModel model(a.sopt.gcBiasCorrect, eqBuilder.readBiasCounts(), eqBuilder.observedGC(), eqBuilder.numFwd(), eqBuilder.numRC(),
                                    fldCounts, a.sopt.pdfSampFactor);
            opt_ok = optimizer.optimizeWithBias(ex, a.sopt, model, 0.01, 10000);
        } else {
            opt_ok = optimizer.optimize(ex, a.sopt, 0.01, 10000);             // SailfishQuantify.cpp:1341-1349
        }
        if (!opt_ok) {
            fprintf(stderr, "[sfb200-quant] %s\n", optimizer.lastError().c_str());
            return 1;
        }
        write_quant_sf(a.out + "/quant.sf", ex);
        const char* samp_type = "none";
        uint32_t n_samples = 0;
        if (a.sopt.numBootstraps || a.sopt.numGibbsSamples) {                 // :1376-1410
            make_dir(aux + "/bootstrap");
            gzFile nf = gzopen((aux + "/bootstrap/names.tsv.gz").c_str(), "wb");
            if (!nf) throw std::runtime_error("cannot open " + aux + "/bootstrap/names.tsv.gz for writing");
            for (size_t i = 0; i < names.size(); ++i) { gzputs(nf, names[i].c_str()); gzputs(nf, i + 1 < names.size() ? "\t" : "\n"); }
            gzclose(nf);
            gzFile bf = gzopen((aux + "/bootstrap/bootstraps.gz").c_str(), "wb");
            if (!bf) throw std::runtime_error("cannot open " + aux + "/bootstrap/bootstraps.gz for writing");
            bool ok = true;
            if (a.sopt.numBootstraps) {
                samp_type = "bootstrap"; n_samples = a.sopt.numBootstraps;
                std::function<bool(const std::vector<double>&)> w = [&](const std::vector<double>& row) {
                    return gzwrite(bf, row.data(), (unsigned)(row.size() * sizeof(double))) > 0;             // GZipWriter.cpp:266-270
                };
                ok = optimizer.gatherBootstraps(ex, a.sopt, w, 0.01, 10000);
            } else {
                samp_type = "gibbs"; n_samples = a.sopt.numGibbsSamples;
                sfb200::CollapsedGibbsSampler sampler(dev);
                std::function<bool(const std::vector<int>&)> w = [&](const std::vector<int>& row) {
                    return gzwrite(bf, row.data(), (unsigned)(row.size() * sizeof(int))) > 0;                // GZipWriter.cpp:279-283
                };
                ok = sampler.sample(ex, a.sopt, w, a.sopt.numGibbsSamples);
            }
            gzclose(bf);
            if (!ok) { fprintf(stderr, "[sfb200-quant] posterior sampling failed: %s\n", optimizer.lastError().c_str()); return 1; }
        }
        // the binary vectors of GZipWriter::writeMeta (src/GZipWriter.cpp:139-161).  The reference never fills the two "expected" vectors
        // (nothing calls setExpectedSeqBias / setExpectedGCBias), so they are what ReadExperiment's constructor leaves: all ones.
        if (!a.sopt.noEffectiveLengthCorrection) write_vector_gz(aux + "/fld.gz", realize_fld(fldCounts));   // fragLengthDist() is unset otherwise
        {
            std::vector<int32_t> obsBias(4096, 1), obsGC(101, 1);              // the initial count of one per bin
            if (doBias) {
                for (size_t i = 0; i < obsBias.size(); ++i) obsBias[i] = static_cast<int32_t>(eqBuilder.readBiasCounts()[i]);
                for (size_t i = 0; i < obsGC.size(); ++i) obsGC[i] = static_cast<int32_t>(eqBuilder.observedGC()[i]);
            }
            write_vector_gz(aux + "/expected_bias.gz", std::vector<double>(4096, 1.0));
            write_vector_gz(aux + "/observed_bias.gz", obsBias);
            write_vector_gz(aux + "/expected_gc.gz", std::vector<double>(101, 1.0));
            write_vector_gz(aux + "/observed_gc.gz", obsGC);
        }
        // meta_info.json (GZipWriter::writeMeta, src/GZipWriter.cpp:163-190)
        FILE* mf = fopen((aux + "/meta_info.json").c_str(), "w");
        if (mf) {
            fprintf(mf, "{\n    \"sf_version\": \"0.10.0-b200\",\n    \"samp_type\": \"%s\",\n    \"frag_dist_length\": %u,\n"
                        "    \"bias_correct\": %s,\n    \"num_bias_bins\": 4096,\n    \"num_targets\": %zu,\n    \"num_bootstraps\": %u,\n    \"num_processed\": %llu,\n"
                        "    \"num_mapped\": %llu,\n    \"percent_mapped\": %.10g,\n    \"call\": \"quant\",\n    \"start_time\": \"%s\",\n"
                        "    \"em_iterations\": %u,\n    \"elapsed_s\": %.3f\n}\n",
                    samp_type, a.mopt.max_frag_len - 1 /* fragLengthDist()->maxValue() */, a.sopt.biasCorrect ? "true" : "false", names.size(), n_samples, (unsigned long long)eqBuilder.numObservedFragments(),
                    (unsigned long long)ex.numMapped, 100.0 * ex.numMapped / std::max<uint64_t>(1, eqBuilder.numObservedFragments()),
                    run_start.c_str(), optimizer.lastIterations(), now_s() - t_start);
            fclose(mf);
        }
        if (!a.geneMap.empty()) gene_level(a.out + "/quant.sf");             // SailfishQuantify.cpp:1413-1422
        fprintf(stderr, "[sfb200-quant] EM: %u iterations; wrote %s/quant.sf (%.2f s in total)\n", optimizer.lastIterations(), a.out.c_str(),
                now_s() - t_start);
        return 0;
    } catch (const sfb200::Error& e) {
        fprintf(stderr, "[sfb200-quant] device error %d: %s\n", e.code, e.what());
        return 3;
    } catch (const std::exception& e) {
        fprintf(stderr, "[sfb200-quant] %s\n", e.what());
        return 1;
    }
}
