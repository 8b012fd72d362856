// Stand-in for RapMap's RapMapSAIndex.hpp: the four fields Sailfish reads (ReadExperiment.hpp:103-116) filled from a
// plain text file "txpinfo.txt" (one "name length" per line) and, when present, "seq.txt" (one transcript sequence per line,
// same order; without it every base is 'A').  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>
template <typename IndexT>
class RapMapSAIndex {
  public:
    bool load(const std::string& dir) {
        std::ifstream in(dir + "txpinfo.txt");
        if (!in) return false;
        std::string name; uint32_t len; IndexT off = 0;
        while (in >> name >> len) { txpNames.push_back(name); txpLens.push_back(len); txpOffsets.push_back(off); off += len + 1; }
        seq.assign(static_cast<size_t>(off) + 1, 'A');
        std::ifstream sq(dir + "seq.txt");
        if (sq) {
            std::string line;
            for (size_t t = 0; t < txpLens.size() && std::getline(sq, line); ++t) {
                if (line.size() != txpLens[t]) return false;
                seq.replace(static_cast<size_t>(txpOffsets[t]), line.size(), line);
                seq[static_cast<size_t>(txpOffsets[t]) + line.size()] = '$';
            }
        }
        return true;
    }
    std::vector<std::string> txpNames;
    std::vector<uint32_t> txpLens;
    std::vector<IndexT> txpOffsets;
    std::string seq;
};
