// crt_spectra.cuh -- spectral binning shared by the CUDA kernel and the host-compiled test harness.
#pragma once

#include "crt_core.cuh"

namespace crt {

// Trapezoidally integrated average of y(x) over the bin [xl, xu]   (ref crt1d/spectra.py:221-258, `_smear_tuv_1`;
// the TUV / F0AM scheme behind `smear_tuv`, ref :261-300).  x ascending, n_x >= 2.  Trapezoids are visited in
// increasing k and accumulated with the reference's expression, starting at the first k with x[k+1] >= xl
// (found by bisection instead of the reference's linear skip) and stopping at the first x[k] > xu.
CRT_HD double smear_tuv_bin(const double* x, const double* y, int n_x, double xl, double xu) {
    int lo = 0, hi = n_x - 1;  // smallest k in [0, n_x - 1] with x[k+1] >= xl, i.e. smallest i = k+1 >= 1 with x[i] >= xl
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid + 1] < xl) lo = mid + 1; else hi = mid;
    }
    double area = 0.0;
    for (int k = lo; k < n_x - 1; ++k) {
        if (x[k + 1] < xl) continue;
        if (x[k] > xu) break;
        const double a1 = x[k] > xl ? x[k] : xl;
        const double a2 = x[k + 1] < xu ? x[k + 1] : xu;
        const double slope = (y[k + 1] - y[k]) / (x[k + 1] - x[k]);
        const double b1 = y[k] + slope * (a1 - x[k]);
        const double b2 = y[k] + slope * (a2 - x[k]);
        area = area + (a2 - a1) * (b2 + b1) / 2;
    }
    return area / (xu - xl);
}

}  // namespace crt
