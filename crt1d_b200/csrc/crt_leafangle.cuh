// crt_leafangle.cuh -- parametric leaf-angle projection functions G(psi) and the hemispherical
// quadratures built on them, for the batched path where the reference's Python callables
// `G_fn` / `K_b_fn` (crt1d/variables.yml:137-162) cannot be called.  Host + device.
#pragma once

#include "../../include/crt1d_b200.h"
#include "crt_core.cuh"

namespace crt {

// G(psi) of the closed families in the reference's leaf_angle.py:118-202.
CRT_HD double leaf_G(int family, double param, double psi) {
    switch (family) {
        case CRT1D_G_SPHERICAL: return 0.5;                               // leaf_angle.py:123-125
        case CRT1D_G_HORIZONTAL: return cos(psi);                         // :118-120
        case CRT1D_G_VERTICAL: return 2.0 / CRT_PI * sin(psi);            // :128-130
        case CRT1D_G_ELLIPSOIDAL_APPROX: {                                // :168-180 (Campbell 1990)
            const double x = param, tp = tan(psi);
            const double p1 = sqrt(x * x + tp * tp);
            const double p2 = x + 1.774 * pow(x + 1.182, -0.733);
            return p1 / p2 * cos(psi);
        }
        case CRT1D_G_ELLIPSOIDAL: {                                       // :133-165 (Campbell 1986)
            const double x = param;
            if (x == 1.0) return 0.5;
            const double tphi = tan(CRT_PI / 2.0 - psi);
            const double p1 = sqrt(x * x + 1.0 / (tphi * tphi));
            double p2;
            if (x > 1.0) {
                const double e1 = sqrt(1.0 - 1.0 / (x * x));
                p2 = x + 1.0 / (2.0 * e1 * x) * log((1.0 + e1) / (1.0 - e1));
            } else {
                const double e2 = sqrt(1.0 - x * x);
                p2 = x + asin(e2) / e2;
            }
            return p1 / p2 * cos(psi);
        }
        case CRT1D_G_ELLIPSOIDAL_APPROX_BONAN: {                          // :183-202 (Ross-Goudriaan)
            const double chil = fmin(fmax(param, -0.4), 0.6);
            const double phi1 = 0.5 - 0.633 * chil - 0.330 * chil * chil;
            const double phi2 = 0.877 * (1.0 - 2.0 * phi1);
            return phi1 + phi2 * cos(psi);
        }
        default: return 0.5;
    }
}

// Gauss-Legendre rule on [-1, 1] (n <= 128), passed to kernels by value.  n == 0 selects the
// reference's nine-sector "9sky" rule for tau_d (common.py:40-53).
struct QuadRule {
    int n;
    int panels;  // tau_d only: number of geometrically graded panels towards psi = pi/2
    double x[128];
    double w[128];
};

// integrand of tau_d: exp(-K_b(psi) L) sin(psi) cos(psi)   (common.py:36)
CRT_HD double tau_d_integrand(int family, double param, double L, double psi) {
    const double c = cos(psi);
    return exp(-(leaf_G(family, param, psi) / c) * L) * sin(psi) * c;
}

CRT_HD double tau_d_quadrature(int family, double param, const QuadRule& rule, double L) {
    if (rule.n == 0) {  // 9sky
        double acc = 0.0;
        for (int k = 0; k < 9; ++k) {
            const double psi = (5.0 + 10.0 * k) * (CRT_PI / 180.0);
            acc += tau_d_integrand(family, param, L, psi);
        }
        return acc * (2.0 * (10.0 * (CRT_PI / 180.0)));
    }
    // composite Gauss-Legendre on [0, pi/2]; panels halve towards pi/2, where exp(-G L / cos psi)
    // has a boundary layer of width ~L that a single panel resolves poorly for thin layers.
    double total = 0.0, a = 0.0;
    const int P = rule.panels < 1 ? 1 : rule.panels;
    for (int p = 0; p < P; ++p) {
        const double b = (p == P - 1) ? CRT_PI / 2.0 : CRT_PI / 2.0 * (1.0 - ldexp(1.0, -(p + 1)));
        const double hw = 0.5 * (b - a), mid = 0.5 * (b + a);
        double acc = 0.0;
        for (int q = 0; q < rule.n; ++q) acc += rule.w[q] * tau_d_integrand(family, param, L, mid + hw * rule.x[q]);
        total += acc * hw;
        a = b;
    }
    return 2.0 * total;
}

// which = 0: mu_bar = int_0^{pi/2} cos(a)/G(a) sin(a) da      (2s, _solve_2s.py:32)
// which = 1: int_0^{mu_s} G(arccos m) dm;  which = 2: int_{mu_s}^1 G(arccos m) dm   (4s, _solve_4s.py:148-149)
CRT_HD double leaf_integral(int family, double param, double mu_s, const QuadRule& rule, int which) {
    double a, b;
    if (which == 0) { a = 0.0; b = CRT_PI / 2.0; }
    else if (which == 1) { a = 0.0; b = mu_s; }
    else { a = mu_s; b = 1.0; }
    const double hw = 0.5 * (b - a), mid = 0.5 * (b + a);
    double acc = 0.0;
    for (int q = 0; q < rule.n; ++q) {
        const double t = mid + hw * rule.x[q];
        acc += rule.w[q] * (which == 0 ? cos(t) / leaf_G(family, param, t) * sin(t) : leaf_G(family, param, acos(t)));
    }
    return acc * hw;
}

// Gauss-Legendre nodes/weights by Newton iteration on P_n (host only; n <= 128).
inline void gauss_legendre(int n, double* x, double* w) {
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = cos(CRT_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-16) break;
        }
        x[i] = -z;
        x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

}  // namespace crt
