// CPU test of sailfish_b200/csrc/em_segments.hpp: an optimizer run cut into launches at iterations 50 / 500 / 1000 must perform exactly
// the iterations, the effective-length updates and the stop decision of one loop with a hook at the top of every iteration
// (the shape of CollapsedEMOptimizer::optimize, reference src/CollapsedEMOptimizer.cpp:820-861).  The "optimizer" is a toy contraction
// whose relative change decays geometrically, so the stop iteration can be steered onto any iteration, including the pauses.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>

#include "../sailfish_b200/csrc/em_segments.hpp"

namespace {

struct Toy {
    double x = 100.0, eff = 1.0, r = 0.9;
    int updates = 0;
    double step() {                                        // one iteration; returns its relative change
        const double nx = r * x + (1.0 - r) * (50.0 / eff);
        const double rel = std::fabs(nx - x) / nx;
        x = nx;
        return rel;
    }
    void update(uint32_t it) { eff *= (it == 50 ? 1.7 : it == 500 ? 0.6 : 1.2); ++updates; }
};

const uint32_t PAUSES[3] = {50, 500, 1000};
bool is_pause(uint32_t it) { return it == 50 || it == 500 || it == 1000; }

struct Result { uint32_t iters; double mrd, x; int updates; std::vector<double> trace; };

// the reference's loop shape
Result straight(Toy t, uint32_t min_iter, uint32_t max_iter, uint32_t fixed, double tol) {
    uint32_t it = 0;
    bool converged = false;
    double mrd = -std::numeric_limits<double>::max();
    std::vector<double> trace;
    auto go = [&]() { return fixed ? it < fixed : (it < min_iter || (it < max_iter && !converged)); };
    while (go()) {
        if (is_pause(it)) t.update(it);
        mrd = t.step();
        trace.push_back(mrd);
        converged = !(mrd > tol);
        ++it;
    }
    return {it, mrd, t.x, t.updates, trace};
}

// one device launch, following the contract in em_segments.hpp; relDiff is poisoned where the loops do not evaluate it
int launch(Toy& t, const sfb::SegLimits& l, double tol, uint32_t* iters, double* mrd_out) {
    const bool fixed = l.fixed_iters > 0;
    uint32_t n = 0;
    double mrd = 1e300;                                    // "not evaluated"
    for (;;) {
        const bool last = fixed ? n >= l.fixed_iters : (n >= l.max_iter && n >= l.min_iter);
        if (last) break;
        if (!fixed && n > 0 && n >= l.min_iter && !(mrd > tol)) break;
        const double rel = t.step();
        ++n;
        const bool evaluated = fixed ? n >= l.fixed_iters : n >= l.min_iter;
        mrd = evaluated ? rel : 1e300;
    }
    *iters = n; *mrd_out = mrd;
    return 0;
}

int fails = 0;
void check(uint32_t min_iter, uint32_t max_iter, uint32_t fixed, double tol, double r, int* n_launches = nullptr) {
    Toy a; a.r = r;
    Toy b = a;
    const Result want = straight(a, min_iter, max_iter, fixed, tol);
    uint32_t iters = 0; double mrd = 0.0;
    int launches = 0;
    uint32_t expect_first = 0;
    bool order_ok = true;
    const int rc = sfb::run_segments(min_iter, max_iter, fixed, tol, PAUSES, 3,
        [&](const sfb::SegLimits& l, uint32_t first, uint32_t* it, double* m) {
            if (first != expect_first) order_ok = false;
            const int e = launch(b, l, tol, it, m);
            expect_first = first + *it; ++launches;
            return e;
        },
        [&](uint32_t it) { if (it != expect_first) order_ok = false; b.update(it); return 0; }, &iters, &mrd);
    if (n_launches) *n_launches = launches;
    const bool mrd_matters = want.iters > 0 && (fixed || want.iters >= min_iter);
    const bool ok = rc == 0 && order_ok && iters == want.iters && b.x == want.x && b.updates == want.updates && (!mrd_matters || mrd == want.mrd);
    if (!ok) {
        ++fails;
        if (fails < 10) fprintf(stderr, "MISMATCH min %u max %u fixed %u tol %g r %g: iters %u/%u x %.17g/%.17g updates %d/%d mrd %g/%g\n", min_iter, max_iter,
                                fixed, tol, r, iters, want.iters, b.x, want.x, b.updates, want.updates, mrd, want.mrd);
    }
}

}  // namespace

int main() {
    std::mt19937_64 rng(11);
    // 1. fixed-iteration runs on and around the pauses
    for (uint32_t f : {1u, 2u, 49u, 50u, 51u, 499u, 500u, 501u, 999u, 1000u, 1001u, 1300u}) check(50, 10000, f, 0.01, 0.9);
    // 2. stop iteration steered onto chosen iterations: tol = relDiff of iteration k-1 exactly, so the run stops with k iterations
    for (double r : {0.9, 0.99, 0.995}) {
        Toy t; t.r = r;
        const Result full = straight(t, 0, 1400, 1400, 0.0);
        for (uint32_t k : {1u, 2u, 30u, 49u, 50u, 51u, 52u, 100u, 499u, 500u, 501u, 502u, 700u, 999u, 1000u, 1001u, 1002u, 1399u}) {
            const double tol = full.trace[k - 1];
            for (uint32_t mn : {0u, 1u, 10u, 50u, 60u, 500u, 600u}) for (uint32_t mx : {1u, 40u, 50u, 51u, 450u, 500u, 1000u, 10000u}) check(mn, mx, 0, tol, r);
        }
    }
    // 3. random limits
    int max_launches = 0;
    for (int rep = 0; rep < 4000; ++rep) {
        const uint32_t mn = (uint32_t)(rng() % 1200), mx = 1 + (uint32_t)(rng() % 1500);
        const uint32_t fixed = (rng() % 4 == 0) ? 1 + (uint32_t)(rng() % 1200) : 0;
        const double tol = std::pow(10.0, -1.0 - (double)(rng() % 600) / 100.0);
        const double r = 0.5 + 0.499 * (double)(rng() % 1000) / 1000.0;
        int nl = 0;
        check(mn, mx, fixed, tol, r, &nl);
        if (nl > max_launches) max_launches = nl;
    }
    if (max_launches != 4) { fprintf(stderr, "expected some run to need 4 launches, saw at most %d\n", max_launches); ++fails; }
    if (fails) { fprintf(stderr, "%d mismatches\n", fails); return 1; }
    printf("em_segments: ok\n");
    return 0;
}
