// TEST INFRASTRUCTURE ONLY.  EmpiricalDistribution restated (reference src/EmpiricalDistribution.cpp:29-125): pdf and cdf are
// stored as float, cdf(x) = 1 beyond the table.  Shared by orc_em.cpp (--unsmoothedFLD effective lengths) and orc_bias.cpp.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

struct EmpDist {
    std::vector<float> pdfvals, cdfvals; float med = 0; uint32_t minVal = 0, maxVal = 0;
    void build(const std::vector<uint32_t>& vals, const std::vector<uint32_t>& lens) {
        const size_t n = vals.size();
        minVal = std::numeric_limits<uint32_t>::max(); maxVal = 0;
        double valsum = 0;
        for (size_t i = 0; i < n; ++i) { minVal = std::min(minVal, vals[i]); maxVal = std::max(maxVal, vals[i]); valsum += lens[i]; }
        double cumpr = 0.0; unsigned lastval = 0, maxval = 1;
        for (; lastval < n; ++lastval) { cumpr += lens[lastval] / valsum; maxval = vals[lastval]; if (cumpr > 1.0 - 1e-6) break; }
        pdfvals.resize(maxval);
        valsum = 0.0;
        for (unsigned i = 0; i < lastval; ++i) valsum += lens[i];
        for (unsigned val = 0, i = 0; val < maxval;) {
            if (val == vals[i]) { pdfvals[val] = static_cast<float>(lens[i] / valsum); ++val; ++i; }
            else if (val < vals[i]) { pdfvals[val] = 0.0f; ++val; }
        }
        cdfvals.resize(maxval);                                                          // :77-81
        if (maxval) cdfvals[0] = pdfvals[0];
        for (unsigned val = 1; val < maxval; ++val) cdfvals[val] = cdfvals[val - 1] + pdfvals[val];
        size_t i = 0, j = n - 1; unsigned u = lens[0], v = lens[n - 1];
        while (i < j) { if (u <= v) { v -= u; u = lens[++i]; } else { u -= v; v = lens[--j]; } }
        med = static_cast<float>(vals[i]);
    }
    float pdf(unsigned x) const { return x < pdfvals.size() ? pdfvals[x] : 0.0f; }
    float cdf(unsigned x) const { return x < cdfvals.size() ? cdfvals[x] : 1.0f; }     // :121-124
};
