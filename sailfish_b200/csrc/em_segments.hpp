// em_segments.hpp -- host-side scheduling of an optimizer run that has to stop at given iteration numbers.
//
// CollapsedEMOptimizer::optimize recomputes the effective lengths at the TOP of iterations 50, 500 and 1000 when bias / GC correction
// is on (reference src/CollapsedEMOptimizer.cpp:820-840).  The device loops (em.cu, em_part.cuh, em_gather.cuh, em_dense.cuh) run a
// whole optimisation in one launch, so such a run is cut into segments: each segment is one ordinary launch with segment-local
// limits, the caller's `update` runs between two segments, and this file decides the limits and when the run is over.
//
// Contract of one launch with limits (min_iter, max_iter, fixed_iters), as all loops implement it:
//     n = 0
//     loop:  last = fixed ? n >= fixed : (n >= max_iter && n >= min_iter);           if last: stop
//            if !fixed && n > 0 && n >= min_iter && !(relDiff of iteration n-1 > tol):   stop
//            iterate; ++n
//     reports n and the relDiff of the last iteration (only evaluated for iterations >= min_iter, or the last one of a fixed run)
// Plain host C++ (no CUDA), so tests/em_segments_test.cpp can replay it against a straightforward loop with a hook.
#ifndef SFB200_EM_SEGMENTS_HPP
#define SFB200_EM_SEGMENTS_HPP

#include <algorithm>
#include <cstdint>
#include <limits>

namespace sfb {

struct SegLimits { uint32_t min_iter, max_iter, fixed_iters; };

inline bool seg_is_pause(uint32_t it, const uint32_t* pauses, int n_pauses) {
    for (int i = 0; i < n_pauses; ++i) if (pauses[i] == it) return true;
    return false;
}
inline uint32_t seg_next_pause(uint32_t it, const uint32_t* pauses, int n_pauses) {
    uint32_t best = std::numeric_limits<uint32_t>::max();
    for (int i = 0; i < n_pauses; ++i) if (pauses[i] > it && pauses[i] < best) best = pauses[i];
    return best;
}

// run(limits, global_iteration_of_the_segment's_first, &iters, &mrd) -> 0 or an error code; update(global_iteration) -> 0 or error.
// `update` runs at the top of every pause iteration the reference's loop would enter.  Returns the first error, else 0.
template <class Run, class Update>
int run_segments(uint32_t min_iter, uint32_t max_iter, uint32_t fixed_iters, double tol, const uint32_t* pauses, int n_pauses,
                 Run&& run, Update&& update, uint32_t* iters_out, double* mrd_out) {
    const bool fixed = fixed_iters > 0;
    const uint32_t end_abs = fixed ? fixed_iters : std::max(min_iter, max_iter);      // the loop never goes past this
    uint32_t g = 0;
    double mrd = -std::numeric_limits<double>::max();
    while (g < end_abs) {                                    // here the reference's loop condition holds at the top of iteration g
        if (seg_is_pause(g, pauses, n_pauses)) { const int rc = update(g); if (rc) return rc; }
        const uint32_t seg_end = std::min(end_abs, seg_next_pause(g, pauses, n_pauses));
        SegLimits lim;
        if (fixed) { lim.fixed_iters = seg_end - g; lim.min_iter = 0; lim.max_iter = seg_end - g; }
        else {
            lim.fixed_iters = 0;
            lim.max_iter = seg_end - g;
            lim.min_iter = min_iter > g ? std::min(min_iter - g, lim.max_iter) : 0;
        }
        uint32_t it = 0;
        const int rc = run(lim, g, &it, &mrd);
        if (rc) return rc;
        g += it;
        if (fixed) continue;
        if (it < lim.max_iter) break;                        // converged inside the segment
        if (g >= min_iter && !(mrd > tol)) break;            // converged on the segment's last iteration
    }
    *iters_out = g;
    *mrd_out = mrd;
    return 0;
}

}  // namespace sfb
#endif
