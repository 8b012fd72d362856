// gene_agg.hpp -- gene-level aggregation of quant.sf: `sailfish quant -g <map>` writes quant.genes.sf.
//
// Restates sailfish::utils::generateGeneLevelEstimates / aggregateEstimatesToGeneLevel (reference src/SailfishUtils.cpp:929-1088)
// and the two ways the reference builds its TranscriptGeneMap: the simple "transcript <white space> gene" format
// (readTranscriptToGeneMap, :438-507) and GTF (transcriptGeneMapFromGTF, :322-436, which uses libgff -- not in the tree; here a
// plain attribute scan: every feature line with a transcript_id contributes transcript -> <key attribute>, first occurrence wins).
// The reference's file type test is the extension ".gtf" (:1051-1059).
//
// Behaviour kept as it is in the reference, including two things one might not expect:
//   * the values are read back from the PRINTED quant.sf (6 significant digits), not taken from memory;
//   * inside the per-gene loop `totalTPM += expVals[tpmIdx]` adds the RUNNING sum (:1011-1016), so for a gene with TPMs a, b, c the
//     normaliser of the length average is a + (a+b) + (a+b+c), not a+b+c.  The gene's TPM and NumReads are the plain sums.
// One deliberate difference: the reference iterates an unordered_map (unspecified order); genes are written in order of first
// appearance in quant.sf.  Header-only, C++11.
#ifndef SFB200_GENE_AGG_HPP
#define SFB200_GENE_AGG_HPP

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace sfb200 {

// transcript name -> gene name; a transcript that is not in the map is its own gene (TranscriptGeneMap::geneName, :128-140)
class TranscriptGeneMap {
public:
    void add(const std::string& txp, const std::string& gene) { if (!map_.count(txp)) { map_[txp] = gene; if (!genes_.count(gene)) genes_[gene] = 1; } }
    std::string geneName(const std::string& txp, bool* found = nullptr) const {
        auto it = map_.find(txp);
        if (found) *found = it != map_.end();
        return it != map_.end() ? it->second : txp;
    }
    size_t numTranscripts() const { return map_.size(); }
    size_t numGenes() const { return genes_.size(); }

private:
    std::unordered_map<std::string, std::string> map_;
    std::unordered_map<std::string, int> genes_;
};

inline TranscriptGeneMap read_simple_gene_map(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open " + path);
    TranscriptGeneMap m;
    std::string t, g;
    while (f >> t >> g) m.add(t, g);
    return m;
}

// value of attribute `key` in a GTF attribute column (`key "value"; key2 "value2"; ...`; unquoted values are accepted too)
inline bool gtf_attribute(const std::string& attrs, const std::string& key, std::string& value) {
    size_t pos = 0;
    while (pos < attrs.size()) {
        while (pos < attrs.size() && (std::isspace((unsigned char)attrs[pos]) || attrs[pos] == ';')) ++pos;
        size_t k0 = pos;
        while (pos < attrs.size() && !std::isspace((unsigned char)attrs[pos]) && attrs[pos] != ';') ++pos;
        const std::string k = attrs.substr(k0, pos - k0);
        while (pos < attrs.size() && std::isspace((unsigned char)attrs[pos])) ++pos;
        std::string v;
        if (pos < attrs.size() && attrs[pos] == '"') {
            const size_t e = attrs.find('"', pos + 1);
            if (e == std::string::npos) return false;
            v = attrs.substr(pos + 1, e - pos - 1);
            pos = e + 1;
        } else {
            size_t v0 = pos;
            while (pos < attrs.size() && attrs[pos] != ';') ++pos;
            v = attrs.substr(v0, pos - v0);
            while (!v.empty() && std::isspace((unsigned char)v.back())) v.pop_back();
        }
        if (k == key) { value = v; return true; }
    }
    return false;
}

inline TranscriptGeneMap read_gtf_gene_map(const std::string& path, const std::string& key) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open " + path);
    TranscriptGeneMap m;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        size_t tab = 0, pos = 0;
        for (int c = 0; c < 8 && tab != std::string::npos; ++c) { tab = line.find('\t', pos); pos = tab == std::string::npos ? pos : tab + 1; }
        if (tab == std::string::npos) continue;                               // fewer than nine columns
        const std::string attrs = line.substr(pos);
        std::string txp, gene;
        if (gtf_attribute(attrs, "transcript_id", txp) && !txp.empty() && gtf_attribute(attrs, key, gene)) m.add(txp, gene);
    }
    return m;
}

inline bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }

struct ExpressionRecord { std::string target; uint32_t length; double effLength; std::vector<double> expVals; };

inline std::string fmt_default(double x) { char b[64]; snprintf(b, sizeof b, "%g", x); return b; }    // ostream << double

// quant.sf -> <same path with the extension replaced by .genes.sf>; returns the output path
inline std::string aggregate_to_gene_level(const TranscriptGeneMap& tgm, const std::string& quantPath, size_t* n_unmapped = nullptr) {
    constexpr double minTPM = std::numeric_limits<double>::denorm_min();
    std::ifstream in(quantPath);
    if (!in.is_open()) throw std::invalid_argument("Attempting to compute gene-level esimtates, but could not \nfind isoform-level file " + quantPath);
    std::vector<std::string> comments, order;
    std::unordered_map<std::string, std::vector<ExpressionRecord>> geneExps;
    std::string l;
    bool headerLine = true;
    size_t unmapped = 0;
    while (std::getline(in, l)) {
        auto it = std::find_if(l.begin(), l.end(), [](char c) { return !std::isspace((unsigned char)c); });
        if (it == l.end()) continue;
        if (*it == '#') { comments.push_back(l); continue; }
        if (headerLine) { comments.push_back(l); headerLine = false; continue; }   // the header line is treated as a comment (:972-975)
        std::istringstream ss(l);
        std::vector<std::string> toks;
        for (std::string t; ss >> t;) toks.push_back(t);
        if (toks.size() < 3) throw std::invalid_argument("Any expression line must contain at least 3 tokens");
        ExpressionRecord er;
        er.target = toks[0]; er.length = (uint32_t)std::stoi(toks[1]); er.effLength = std::stod(toks[2]);
        for (size_t i = 3; i < toks.size(); ++i) er.expVals.push_back(std::stod(toks[i]));
        bool found = false;
        const std::string gn = tgm.geneName(er.target, &found);
        if (!found) ++unmapped;
        if (!geneExps.count(gn)) order.push_back(gn);
        geneExps[gn].push_back(std::move(er));
    }
    in.close();
    std::string outPath = quantPath;
    const size_t dot = outPath.find_last_of('.'), slash = outPath.find_last_of('/');
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) outPath.erase(dot);
    outPath += ".genes.sf";
    FILE* out = fopen(outPath.c_str(), "w");
    if (!out) throw std::runtime_error("cannot write " + outPath);
    for (const std::string& c : comments) fprintf(out, "%s\n", c.c_str());
    for (const std::string& gn : order) {
        const std::vector<ExpressionRecord>& recs = geneExps[gn];
        double geneLength = 0.0, geneEffLength = 0.0;
        std::vector<double> expVals(recs.front().expVals.size(), 0.0);
        const size_t NE = expVals.size(), tpmIdx = 0;
        double totalTPM = 0.0;
        for (const ExpressionRecord& r : recs) {
            for (size_t i = 0; i < NE && i < r.expVals.size(); ++i) expVals[i] += r.expVals[i];
            if (NE) totalTPM += expVals[tpmIdx];                               // the running sum, as the reference has it (:1011-1016)
        }
        if (totalTPM > minTPM) {
            for (const ExpressionRecord& r : recs) {
                const double frac = (NE ? r.expVals[tpmIdx] : 0.0) / totalTPM;
                geneLength += r.length * frac; geneEffLength += r.effLength * frac;
            }
        } else {
            const double frac = 1.0 / recs.size();
            for (const ExpressionRecord& r : recs) { geneLength += r.length * frac; geneEffLength += r.effLength * frac; }
        }
        fprintf(out, "%s\t%s\t%s", gn.c_str(), fmt_default(geneLength).c_str(), fmt_default(geneEffLength).c_str());
        for (size_t i = 0; i < NE; ++i) fprintf(out, "\t%s", fmt_default(expVals[i]).c_str());
        fprintf(out, "\n");
    }
    fclose(out);
    if (n_unmapped) *n_unmapped = unmapped;
    return outPath;
}

// generateGeneLevelEstimates (:1043-1088): map type by extension, then aggregate <estDir>/quant.sf
inline std::string generate_gene_level_estimates(const std::string& geneMapPath, const std::string& quantPath, const std::string& aggKey,
                                                 size_t* n_txp = nullptr, size_t* n_genes = nullptr, size_t* n_unmapped = nullptr) {
    const TranscriptGeneMap tgm = ends_with(geneMapPath, ".gtf") ? read_gtf_gene_map(geneMapPath, aggKey) : read_simple_gene_map(geneMapPath);
    if (n_txp) *n_txp = tgm.numTranscripts();
    if (n_genes) *n_genes = tgm.numGenes();
    return aggregate_to_gene_level(tgm, quantPath, n_unmapped);
}

}  // namespace sfb200
#endif
