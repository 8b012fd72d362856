// Host build of sailfish_b200/csrc/vb_math.hpp: reads x values (one per line) on stdin, prints "digamma(x) exp_digamma(x)" per line with
// 17 significant digits.  tests/test_vb_math.py compares them with mpmath.
#include <cstdio>

#include "../sailfish_b200/csrc/vb_math.hpp"

int main() {
    double x;
    while (scanf("%lf", &x) == 1) printf("%.17g %.17g\n", sfb_digamma(x), sfb_exp_digamma(x));
    return 0;
}
