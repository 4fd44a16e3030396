// crt_core.cuh -- per-column arithmetic of every scheme, shared by the CUDA kernels (crt_kernels.cu)
// and by the host-compiled unit-test harness (tests/_hostcheck).  A "column" is one (scenario, band):
// all n_z interface levels of one wavelength band of one scenario.  VEC adjacent bands are processed
// together so that the kernels can issue 16-byte stores and have two independent dependency chains.
//
// Every function cites the reference lines (zmoon/crt1d, paths relative to crt1d/solvers/) whose
// arithmetic it reproduces.  Nothing here is copied: the reference loops over bands in Python and
// evaluates whole-profile numpy expressions; this code is organised as  scenario prologue (level
// tables) -> per-band coefficients -> level sweep, with level-independent factors folded.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CRT_HD __host__ __device__ __forceinline__
#else
#define CRT_HD inline
#endif

#if defined(__CUDACC__)
#define CRT_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define CRT_HD_NOINLINE static
#endif

namespace crt {

// Remember three values of one level (a canopy-end level, for the absorbed reduction) -- deliberately NOT inlined:
// as inline assignments under `if (j == 0)` the compiler turns them into per-level selects (zq_pa: 12 FSEL of 224
// instructions per layer.band); a call stays a branch that is taken twice per column.
CRT_HD_NOINLINE void keep3(double* dst, double a, double b, double c) {
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
}

// Output field slots of the column accessors.
enum Field : int { F_IDR = 0, F_DN = 1, F_UP = 2, F_F = 3, F_X0 = 4, F_X1 = 5, F_X2 = 6, N_FIELDS = 7 };

#ifndef CRT_PI
#define CRT_PI 3.14159265358979323846
#endif


// ---------------------------------------------------------------------------------------------
// exp_pm: e^{-x} and e^{+x} for x >= 0 from ONE range reduction.
// Every level sweep needs a decaying exponential and its reciprocal (2s: e^{-hL}, e^{+hL}; 4s:
// e^{-lambda x}, e^{-lambda (LAI - x)} = e^{-lambda LAI} e^{+lambda x}; bf: e^{-k_d L}, e^{-k_d (L_T - L)}).
// libm-style exp() + an IEEE division cost ~50 issue slots per pair on sm_100a (17 FP64 ops + 13
// constant moves + range checks for exp, MUFU.RCP64H + 5 DFMA + fix-up branch for the reciprocal).
// Here: x = n ln2 + r, |r| <= ln2/2;  e^{+-r} = cosh r +- sinh r, both even/odd Taylor polynomials in
// r^2 (truncation < 5e-18), and 2^{+-n} applied by integer adds on the exponent field: 21 FP64 ops and
// ~6 integer ops for BOTH values, no division, no branches on the fast path.  Max relative error
// ~2 ulp (tests/test_hostmath.py::test_exp_pm_accuracy).  x > 700 (results beyond the normal range)
// and NaN take the libm path.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define CRT_CONST __constant__
#else
#define CRT_CONST static const
#endif
// 1/(2k)! for k = 6..1 then 1/(2k+1)! for k = 6..1
CRT_CONST double kExpPm[12] = {
    1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5,
    1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0};

// rcp_nr: 1/x for the per-level reciprocals of the Thomas sweeps.  The compiler's IEEE division costs ~20
// issue slots on sm_100a (MUFU.RCP64H, 5 DFMA, exponent checks and a convergent slow-path branch); the
// pivots here are O(1) and never zero, subnormal or infinite, so the seed (rcp.approx.ftz.f64: the top
// ~20 bits) and two Newton steps suffice: 5 instructions, result within 1 ulp of the rounded quotient.
// The host build (tests/_hostcheck) uses the plain division.
CRT_HD double rcp_nr(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

CRT_HD double scale_pow2(double v, int n) {  // v * 2^n for results that stay normal
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) + (n << 20), __double2loint(v));
#else
    return ldexp(v, n);
#endif
}

CRT_HD void exp_pm(double x, double& em, double& ep) {
    if (!(x <= 700.0)) {  // also catches NaN
        em = exp(-x);
        ep = exp(x);
        return;
    }
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer
    const double nf = fma(x, 1.4426950408889634074, kMagic);
    const double nd = nf - kMagic;
#if defined(__CUDA_ARCH__)
    const int n = __double2loint(nf);
#else
    const int n = (int)nd;
#endif
    double r = fma(nd, -6.93147180559945286227e-01, x);   // ln2 hi
    r = fma(nd, -2.31904681384629955842e-17, r);          // ln2 lo
    const double r2 = r * r;
    double c = kExpPm[0], sn = kExpPm[6];
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        c = fma(c, r2, kExpPm[i]);
        sn = fma(sn, r2, kExpPm[6 + i]);
    }
    c = fma(c, r2, 1.0);              // cosh r
    sn = fma(sn * r2, r, r);          // sinh r = r + r^3 (1/6 + ...)
    ep = scale_pow2(c + sn, n);
    em = scale_pow2(c - sn, -n);
}

// e^{-x} for x >= 0 (the unused e^{+x} half of exp_pm is dead code the compiler drops).
CRT_HD double exp_neg(double x) {
    double em, ep;
    exp_pm(x, em, ep);
    return em;
}

// Band inputs of one column group.
template <int VEC>
struct BandIn {
    double leaf_r[VEC], leaf_t[VEC], soil_r[VEC], Idr0[VEC], Idf0[VEC];
};

// Canopy-integrated absorbed irradiance of a column from its ground (0) and top (n_z-1) level values:
// Sum over layers of  a = dI_dr + dI_df_d - dI_df_u  (ref ../model.py:606-609) telescopes to the ends.
CRT_HD double absorbed_from_ends(double Idr_top, double Idr_gnd, double dn_top, double dn_gnd,
                                 double up_top, double up_gnd) {
    return (Idr_top - Idr_gnd) + (dn_top - dn_gnd) + (up_gnd - up_top);
}

// =================================================================================================
// 2s  Dickinson-Sellers two-stream   (ref _solve_2s.py:11-163)
// =================================================================================================
struct Scen2s {
    double K;        // K_b_fn(psi), black-leaf extinction          (ref :26, :40)
    double inv_mu;   // 1/cos(psi)
    double mu_bar;   // (ref :32)
    double inv_mu_bar;
    double cos2_tl;  // cos^2(radians(mla))                         (ref :28, :68)
    double as_fac;   // 1 - mu log((mu+1)/mu)                       (ref :73)
    double L_T;      // total LAI = lai[0]                          (ref :38)
    double S2;       // exp(-K L_T)                                 (ref :91)
};

CRT_HD Scen2s scen_2s(double psi, double K, double mu_bar, double mla_deg, double L_T) {
    Scen2s s;
    const double mu = cos(psi);
    s.K = K;
    s.inv_mu = 1.0 / mu;
    s.mu_bar = mu_bar;
    s.inv_mu_bar = 1.0 / mu_bar;
    const double ct = cos(mla_deg * (CRT_PI / 180.0));
    s.cos2_tl = ct * ct;
    s.as_fac = 1.0 - mu * log((mu + 1.0) / mu);
    s.L_T = L_T;
    s.S2 = exp(-K * L_T);
    return s;
}

// Folded per-band coefficients:  I_up(L) = Au eK(L) + Bu e^{-hL} + Cu e^{+hL},  I_dn likewise.
struct Coef2s {
    double h, Au, Bu, Cu, Ad, Bd, Cd, Idr0;
};

// RCP: reciprocals by rcp_nr instead of IEEE divisions.  Measured A/B on one box: the reduced-diagnostic kernel gains
// 10 % (44.2 -> 40.0 ms per 10^6 scenarios), the full row-sweep kernel LOSES 1 % (0.961 -> 0.951 of HBM peak: the
// first-touch item shares 128 registers with the level sweep), so only the reduced-diagnostic instantiation uses it.
template <bool RCP = false>
CRT_HD Coef2s coef_2s(const Scen2s& s, double alpha, double tau, double rho_s, double Idr0, double Idf0) {
    auto rcp_coef = [](double x) { return RCP ? rcp_nr(x) : 1.0 / x; };
    // The reference divides by S1, D1, D2, sigma in a dozen places (ref :96-120); here each reciprocal is
    // formed once (1/S1 = e^{+h L_T} falls out of exp_pm; the others by division, or with RCP by rcp_nr: seed +
    // two Newton steps instead of the ~20-slot IEEE division sequence) -- same algebra, results within a few ulp.
    const double mu_bar = s.mu_bar, K = s.K;
    const double omega = alpha + tau;                                                   // ref :65
    const double beta = (0.5 * (alpha + tau + (alpha - tau) * s.cos2_tl)) * rcp_coef(omega);  // ref :68 (eq. 3)
    const double a_s = omega / 2.0 * s.as_fac;                                          // ref :73
    const double mK = mu_bar * K;
    const double beta_0 = (1.0 + mK) * rcp_coef(omega * mK) * a_s;                        // ref :76 (eq. 4)
    const double b = 1.0 - (1.0 - beta) * omega;                                        // ref :80-85
    const double c = omega * beta;
    const double d = omega * mK * beta_0;
    const double f = omega * mK * (1.0 - beta_0);
    const double h = sqrt(b * b - c * c) * s.inv_mu_bar;
    const double sigma = mK * mK + c * c - b * b;
    const double u1 = b - c / rho_s;                                                    // ref :87-97
    const double u2 = b - c * rho_s;
    const double u3 = f + c * rho_s;
    double S1, iS1;                                                                     // e^{-h L_T}, e^{+h L_T}
    exp_pm(h * s.L_T, S1, iS1);
    const double S2 = s.S2;
    const double mh = mu_bar * h;
    const double p1 = b + mh, p2 = b - mh, p3 = b + mK, p4 = b - mK;
    const double iD1 = rcp_coef(p1 * (u1 - mh) * iS1 - p2 * (u1 + mh) * S1);
    const double iD2 = rcp_coef((u2 + mh) * iS1 - (u2 - mh) * S1);
    const double isig = rcp_coef(sigma);
    const double h1s = (-d * p4 - c * f) * isig;                                        // h1 / sigma, ref :99
    const double t1 = d - h1s * p3;
    const double t2 = d - c - h1s * (u1 + mK);
    const double h2 = iD1 * (t1 * (u1 - mh) * iS1 - p2 * t2 * S2);
    const double h3 = -iD1 * (t1 * (u1 + mh) * S1 - p1 * t2 * S2);
    const double h4s = (-f * p3 - c * d) * isig;                                        // h4 / sigma (Sellers 1996), ref :109
    const double t3 = u3 - h4s * (u2 - mK);
    const double h5 = -iD2 * (h4s * (u2 + mh) * iS1 + t3 * S2);
    const double h6 = iD2 * (h4s * (u2 - mh) * S1 + t3 * S2);
    const double h7 = c * iD1 * (u1 - mh) * iS1;
    const double h8 = -c * iD1 * (u1 + mh) * S1;
    const double h9 = iD2 * (u2 + mh) * iS1;
    const double h10 = -iD2 * (u2 - mh) * S1;
    Coef2s k;  // fold  I_dr0 * (h1 eK/sigma + h2 em + h3 ep) + I_df0 * (h7 em + h8 ep)   (ref :125-135)
    k.h = h;
    k.Au = Idr0 * h1s;
    k.Bu = Idr0 * h2 + Idf0 * h7;
    k.Cu = Idr0 * h3 + Idf0 * h8;
    k.Ad = Idr0 * h4s;
    k.Bd = Idr0 * h5 + Idf0 * h9;
    k.Cd = Idr0 * h6 + Idf0 * h10;
    k.Idr0 = Idr0;
    return k;
}

// One level of one 2s column given em = exp(-hL), ep = exp(+hL), eK = exp(-K L)   (ref :125-156).
CRT_HD void level_2s_e(const Coef2s& k, double inv_mu, double eK, double em, double ep, double& Idr, double& dn,
                       double& up, double& F) {
    up = k.Au * eK + (k.Bu * em + k.Cu * ep);
    dn = k.Ad * eK + (k.Bd * em + k.Cd * ep);
    Idr = k.Idr0 * eK;                          // ref :150
    F = Idr * inv_mu + 2.0 * up + 2.0 * dn;     // ref :156
}

CRT_HD void level_2s(const Coef2s& k, double inv_mu, double L, double eK, double& Idr, double& dn, double& up,
                     double& F) {
    double em, ep;
    exp_pm(k.h * L, em, ep);
    level_2s_e(k, inv_mu, eK, em, ep, Idr, dn, up, F);
}

// Level sweep of VEC 2s columns.  `L`, `eK` are the scenario's level tables (eK[j] = exp(-K L[j])).
template <int VEC, class Out>
CRT_HD void column_2s(const Scen2s& s, const double* L, const double* eK, int n_z, const BandIn<VEC>& in,
                      Out& out, double (&absorbed)[VEC]) {
    Coef2s k[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) k[v] = coef_2s(s, in.leaf_r[v], in.leaf_t[v], in.soil_r[v], in.Idr0[v], in.Idf0[v]);
    double gnd[VEC][3];
    for (int j = 0; j < n_z; ++j) {
        const double Lj = L[j], eKj = eK[j];
        double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            level_2s(k[v], s.inv_mu, Lj, eKj, Idr[v], dn[v], up[v], F[v]);
            if (j == 0) { gnd[v][0] = Idr[v]; gnd[v][1] = dn[v]; gnd[v][2] = up[v]; }
            if (j == n_z - 1) absorbed[v] = absorbed_from_ends(Idr[v], gnd[v][0], dn[v], gnd[v][1], up[v], gnd[v][2]);
        }
        out.st(F_IDR, j, Idr);
        out.st(F_DN, j, dn);
        out.st(F_UP, j, up);
        out.st(F_F, j, F);
    }
}

// =================================================================================================
// bl  Beer-Lambert   (ref _solve_bl.py:9-93)
// =================================================================================================
struct ScenBl {
    double K_b, inv_mu;
};

struct CoefBl {
    double Kg, Idr0, Idf0;  // grey-leaf extinction K_b sqrt(1 - omega) (ref :58-62)
};

CRT_HD CoefBl coef_bl(const ScenBl& s, double r, double t, double Idr0, double Idf0) {
    CoefBl k;
    k.Kg = s.K_b * sqrt(1.0 - (t + r));
    k.Idr0 = Idr0;
    k.Idf0 = Idf0;
    return k;
}

// tb = exp(-K_b L) (ref :31), td = tau_df_fn(K_b_fn, L) (ref :35-37): scenario level tables
CRT_HD void level_bl(const ScenBl& s, const CoefBl& k, double L, double tb, double td, double& Idr, double& dn,
                     double& up, double& F) {
    const double tau_g = exp_neg(k.Kg * L);               // ref :65
    Idr = k.Idr0 * tb;                                    // ref :69
    dn = k.Idf0 * td + 0.5 * (k.Idr0 * (tau_g - tb));     // ref :70-79
    up = 0.0;                                             // ref :87
    F = Idr * s.inv_mu + 2.0 * dn;                        // ref :90
}

template <int VEC, class Out>
CRT_HD void column_bl(const ScenBl& s, const double* L, const double* tau_b, const double* tau_d, int n_z,
                      const BandIn<VEC>& in, Out& out, double (&absorbed)[VEC]) {
    CoefBl k[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) k[v] = coef_bl(s, in.leaf_r[v], in.leaf_t[v], in.Idr0[v], in.Idf0[v]);
    double gnd[VEC][2];
    for (int j = 0; j < n_z; ++j) {
        double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            level_bl(s, k[v], L[j], tau_b[j], tau_d[j], Idr[v], dn[v], up[v], F[v]);
            if (j == 0) { gnd[v][0] = Idr[v]; gnd[v][1] = dn[v]; }
            if (j == n_z - 1) absorbed[v] = absorbed_from_ends(Idr[v], gnd[v][0], dn[v], gnd[v][1], 0.0, 0.0);
        }
        out.st(F_IDR, j, Idr);
        out.st(F_DN, j, dn);
        out.st(F_UP, j, up);
        out.st(F_F, j, F);
    }
}

// =================================================================================================
// bf  Bodin & Franklin (2012)   (ref _solve_bf.py:7-154)
// g77 Goudriaan (1977)          (ref _solve_g77.py:7-135)
// =================================================================================================
struct ScenBf {
    double k_b, mu, inv_mu, L_T;
    double eb0;  // exp(-k_b L_T) = A_sl at the ground
};

CRT_HD ScenBf scen_bf(double psi, double K_b, double L_T) {
    ScenBf s;
    s.k_b = K_b;
    s.mu = cos(psi);
    s.inv_mu = 1.0 / s.mu;
    s.L_T = L_T;
    s.eb0 = exp(-K_b * L_T);
    return s;
}

// Folded per-band coefficients shared by bf and g77.  Per level both schemes need e^{-k_d L} and
// e^{-k_d (L_T - L)} = ed0 e^{+k_d L}: one exp_pm.  bf: a1, a2 scale (eb - ed), (eb - ex) (eq. 8, 9);
// g77: a1 = I_dr0 (1 - rho_c), a2 = -I_dr0 (1 - sigma), kg = k' k_b (eq. 5), adf = I_df0 (1 - rho_c).
struct CoefBf {
    double k_d, ed0, adf, Idr0, a1, a2, soil, c1, c2, c3, kg, rho_c;
};

CRT_HD CoefBf coef_bfg_common(const ScenBf& s, double r_l, double t_l, double Idr0) {
    CoefBf k;
    const double sigma = r_l + t_l;
    const double k_prime = sqrt(1.0 - sigma);                                    // ref _solve_bf.py:69
    k.rho_c = ((1.0 - k_prime) / (1.0 + k_prime)) * (2.0 / (1.0 + 1.6 * s.mu));  // Spitters (1986) eq. 1 (ref :78)
    k.k_d = 0.8 * sqrt(1.0 - sigma);                                             // B&F eq. 2 (ref :81)
    k.c1 = k.k_d / k_prime;                                                      // eq. 14/15 factors (ref :118-130)
    k.c2 = k.k_d / sqrt(1.0 - r_l);
    k.c3 = k.k_d / sqrt(1.0 - t_l);
    k.Idr0 = Idr0;
    k.ed0 = exp(-k.k_d * s.L_T);
    k.kg = k_prime * s.k_b;
    k.adf = k.a1 = k.a2 = k.soil = 0.0;
    return k;
}

CRT_HD CoefBf coef_bf(const ScenBf& s, double r_l, double t_l, double soil_r, double Idr0, double Idf0) {
    CoefBf k = coef_bfg_common(s, r_l, t_l, Idr0);
    k.adf = Idf0;                                                // ref :86 (no (1 - rho_c) factor in bf)
    k.a1 = Idr0 * t_l / (k.k_d - s.k_b);                         // eq. 8 (ref :97); k_d = k_b is a pole the reference shares
    k.a2 = Idr0 * r_l / (k.k_d + s.k_b);                         // eq. 9 (ref :101-105)
    const double I_df_g = k.adf * k.ed0;                         // ground-level values for eq. 11 (ref :114)
    const double I_sc_d_g = k.a1 * (s.eb0 - k.ed0);
    k.soil = soil_r * (Idr0 * s.eb0 + I_df_g + I_sc_d_g);
    return k;
}

CRT_HD CoefBf coef_g77(const ScenBf& s, double r_l, double t_l, double soil_r, double Idr0, double Idf0) {
    CoefBf k = coef_bfg_common(s, r_l, t_l, Idr0);
    k.adf = Idf0 * (1.0 - k.rho_c);                              // ref _solve_g77.py:73
    k.a1 = Idr0 * (1.0 - k.rho_c);                               // ref :84
    k.a2 = -Idr0 * (1.0 - (r_l + t_l));                          // ref :84-86
    const double I_df_g = k.adf * k.ed0;
    const double I_sc_g = k.a1 * exp(-k.kg * s.L_T) + k.a2 * s.eb0;
    k.soil = soil_r * (Idr0 * s.eb0 + I_df_g + 0.5 * I_sc_g);    // ref :95
    return k;
}

// f = {I_dr, I_df_d, I_df_u, F, aI_lsl, aI_lsh, aI_l}; eb = exp(-k_b L) from the scenario level table
// level_bfg_e: the level given its exponentials ed = exp(-k_d L), ep = exp(+k_d L), eg = exp(-kg L) (g77 only)
template <bool G77>
CRT_HD void level_bfg_e(const ScenBf& s, const CoefBf& k, double eb, double ed, double ep, double eg, double (&f)[7]);

template <bool G77>
CRT_HD void level_bfg(const ScenBf& s, const CoefBf& k, double L, double eb, double (&f)[7]) {
    double ed, ep;                                  // exp(-k_d L), exp(+k_d L)
    exp_pm(k.k_d * L, ed, ep);
    level_bfg_e<G77>(s, k, eb, ed, ep, G77 ? exp_neg(k.kg * L) : 0.0, f);
}

template <bool G77>
CRT_HD void level_bfg_e(const ScenBf& s, const CoefBf& k, double eb, double ed, double ep, double eg, double (&f)[7]) {
    const double e2 = k.ed0 * ep;                   // exp(-k_d (L_T - L))   (ref _solve_bf.py:114)
    const double I_df = k.adf * ed;                 // ref _solve_bf.py:86 / _solve_g77.py:73
    const double Idr = k.Idr0 * eb;                 // ref :90
    double I_sc_d, I_sc_u;
    if (G77) {
        const double I_sc = k.a1 * eg + k.a2 * eb;                  // eq. 5 (ref _solve_g77.py:84-86)
        I_sc_d = 0.5 * I_sc;                                        // ref :89-90
        I_sc_u = 0.5 * I_sc;
    } else {
        I_sc_d = k.a1 * (eb - ed);                                  // eq. 8
        I_sc_u = k.a2 * (eb - e2 * s.eb0);                          // eq. 9: exp(k_d L - (k_b + k_d) L_T) = e2 eb0
    }
    const double I_sr = k.soil * e2;                                // eq. 11
    const double diff = k.c1 * I_df + k.c2 * I_sc_u + k.c3 * I_sc_d;
    const double sh = (1.0 - eb) * diff;                            // eq. 14
    const double sl = eb * (diff + s.k_b * k.Idr0);                 // eq. 15
    const double dn = I_sc_d + I_df, up = I_sc_u + I_sr;
    f[0] = Idr;
    f[1] = dn;
    f[2] = up;
    f[3] = Idr * s.inv_mu + 2.0 * up + 2.0 * dn;
    f[4] = sl;
    f[5] = sh;
    f[6] = sl + sh;
}

template <bool G77, int VEC, class Out>
CRT_HD void column_bfg(const ScenBf& s, const double* L, const double* eb, int n_z, const BandIn<VEC>& in, Out& out,
                       double (&rho_c)[VEC], double (&absorbed)[VEC]) {
    CoefBf k[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        k[v] = G77 ? coef_g77(s, in.leaf_r[v], in.leaf_t[v], in.soil_r[v], in.Idr0[v], in.Idf0[v])
                   : coef_bf(s, in.leaf_r[v], in.leaf_t[v], in.soil_r[v], in.Idr0[v], in.Idf0[v]);
        rho_c[v] = k[v].rho_c;
    }
    double gnd[VEC][3];
    for (int j = 0; j < n_z; ++j) {
        double o[7][VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double f[7];
            level_bfg<G77>(s, k[v], L[j], eb[j], f);
#pragma unroll
            for (int q = 0; q < 7; ++q) o[q][v] = f[q];
            if (j == 0) { gnd[v][0] = f[0]; gnd[v][1] = f[1]; gnd[v][2] = f[2]; }
            if (j == n_z - 1) absorbed[v] = absorbed_from_ends(f[0], gnd[v][0], f[1], gnd[v][1], f[2], gnd[v][2]);
        }
        out.st(F_IDR, j, o[0]);
        out.st(F_DN, j, o[1]);
        out.st(F_UP, j, o[2]);
        out.st(F_F, j, o[3]);
        out.st(F_X0, j, o[4]);
        out.st(F_X1, j, o[5]);
        out.st(F_X2, j, o[6]);
    }
}

template <int VEC, class Out>
CRT_HD void column_bf(const ScenBf& s, const double* L, const double* eb, int n_z, const BandIn<VEC>& in, Out& out,
                      double (&rho_c)[VEC], double (&absorbed)[VEC]) {
    column_bfg<false, VEC>(s, L, eb, n_z, in, out, rho_c, absorbed);
}

template <int VEC, class Out>
CRT_HD void column_g77(const ScenBf& s, const double* L, const double* eb, int n_z, const BandIn<VEC>& in, Out& out,
                       double (&absorbed)[VEC]) {
    double rho_c[VEC];
    column_bfg<true, VEC>(s, L, eb, n_z, in, out, rho_c, absorbed);
}

// =================================================================================================
// n79  Norman (1979) after Bonan SP 14.3   (ref _solve_n79.py:11-200)
//
// The 2 n_z unknowns alternate (up_j, dn_j) per level j.  The reference's Thomas algorithm (`tdma`,
// ref :167-200) is reproduced row for row; the forward-sweep coefficients e, f of the two rows of
// level j are parked in the four OUTPUT arrays at level j (I_dr <- e_up, F <- e_dn, I_df_u <- f_up,
// I_df_d <- f_dn) and overwritten with the final values during back-substitution, so the solve needs
// no scratch memory beyond the arrays it has to write anyway.  (Only the downward row's pair of every
// CK-th level is parked -- checkpoints; everything between is recomputed in the back sweep.)
// Level tables: everything of a row that does not depend on the band (fill_level_tables<N79>; ref :41-59).
// =================================================================================================
struct ScenN79 {
    double inv_mu;
};

// Layer coefficients shared by the upward row above a layer and the downward row below it:
// aiv = fiv = refld - trand^2/refld,  biv = eiv = trand/refld   (ref :85-88, :98-101, :111-114), plus the
// band-dependent factors of the two beam sources, rho - tau eiv (:108/:129) and tau - rho biv (:92/:118).
// One reciprocal instead of the reference's two divisions (same algebra; <= 1 ulp apart).
template <int VEC>
struct N79Layer {
    double f[VEC], e[VEC], wu[VEC], wd[VEC];
};

// tab = the eight level tables of fill_level_tables<N79>: tbcum, g, td, 1 - td, fsun, 1 - fsun, 1/(fsun dlai),
// 1/((1 - fsun) dlai).
template <int VEC, class Out>
CRT_HD void column_n79(const ScenN79& s, const double* tab, int n_z, const BandIn<VEC>& in, Out& out,
                       double (&absorbed)[VEC]) {
    const double *tbcum = tab, *g = tab + n_z, *td = tab + 2 * n_z, *omtd = tab + 3 * n_z, *fsun = tab + 4 * n_z,
                 *omfs = tab + 5 * n_z, *isl = tab + 6 * n_z, *ish = tab + 7 * n_z;
    // Segment store: the back sweep keeps BOTH rows' pairs of a recomputed level (downward row in slots 0..CK-1,
    // upward row in slots CK..2CK-1), so it never re-derives a layer or an upward row; half the levels per segment.
    const int CK = out.seg_levels() / 2;
    using Lay = N79Layer<VEC>;
    auto layer = [&](int i, Lay& L) {
        const double t = td[i], o = omtd[i];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double rho = in.leaf_r[v], tau = in.leaf_t[v];
            const double refld = o * rho;
            const double trand = o * tau + t;
            const double ir = rcp_nr(refld);
            L.e[v] = trand * ir;
            L.f[v] = refld - trand * L.e[v];
            L.wu[v] = rho - tau * L.e[v];
            L.wd[v] = tau - rho * L.e[v];
        }
    };
    // Equal layer thickness (every profile the reference's LAI generators make, ref ../leaf_area.py:82-88): tau_d is
    // the same in every layer up to the rounding of lai[j] - lai[j+1], so the layer coefficients are formed once per
    // column instead of once per level and sweep.  Criterion |td[j] - td[0]| <= 16 eps td[0]: the synthetic sweep's 100
    // profiles deviate by up to 8.7 eps (device prologue), and 16 eps times the conditioning of the solve (1e2..4e3)
    // stays an order of magnitude inside the 1e-10 parity bar.
    bool uni = true;
    for (int i = 1; i < n_z - 1; ++i) uni = uni && fabs(td[i] - td[0]) <= 3.6e-15 * td[0];
    // Forward-sweep coefficients of the UPWARD row of level j >= 1 (a = -eiv, c = -fiv of the layer below, L) from
    // those of the downward row of level j-1 (e_in, f_in); unit diagonal (ref :101-108 / :122-129, tdma :186, :191).
    // Used identically in both sweeps, so the back-substitution recomputes the forward values.
    auto up_row = [&](int j, const Lay& L, const double (&e_in)[VEC], const double (&f_in)[VEC], double (&eu)[VEC],
                      double (&fu)[VEC]) {
        const double gj = g[j];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double d = in.Idr0[v] * gj * L.wu[v];
            const double r = rcp_nr(1.0 + L.e[v] * e_in[v]);  // one reciprocal for both quotients of tdma
            eu[v] = -L.f[v] * r;
            fu[v] = (d + L.e[v] * f_in[v]) * r;
        }
    };
    auto soil_row = [&](double (&eu)[VEC], double (&fu)[VEC]) {  // level 0 (ref :79-82)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            eu[v] = -in.soil_r[v];
            fu[v] = in.Idr0[v] * tbcum[0] * in.soil_r[v];
        }
    };
    // Downward row of a level (a = -aiv, c = -biv of layer Ld; ref :85-92 / :111-118) from the upward row's pair.
    auto down_row = [&](double gj, const Lay& Ld, const double (&eu)[VEC], const double (&fu)[VEC], double (&e)[VEC],
                        double (&f)[VEC]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double d = in.Idr0[v] * gj * Ld.wd[v];
            const double r = rcp_nr(1.0 + Ld.f[v] * eu[v]);
            e[v] = -Ld.e[v] * r;
            f[v] = (d + Ld.f[v] * fu[v]) * r;
        }
    };
    // One level of the forward recurrence: (e, f) holds the downward row's pair of level j-1 on entry and of level j
    // on exit; La holds layer j-1 on entry and layer j on exit (the layer the next upward row needs; with `uni` it
    // holds THE layer throughout).  The soil level and the interior/top levels are separate functions, so the level
    // loops carry no per-level selects or struct copies (they were 35 of 166 instructions per layer.band).
    auto step_soil = [&](Lay& La, double (&e)[VEC], double (&f)[VEC], double (&eu)[VEC], double (&fu)[VEC]) {
        soil_row(eu, fu);
        if (uni) {
            down_row(g[0], La, eu, fu, e, f);
        } else {  // the soil-adjacent downward row uses layer 1 as shipped (ref :85-92)
            Lay L1;
            layer(1, L1);
            down_row(g[0], L1, eu, fu, e, f);
            layer(0, La);
        }
    };
    auto step_in = [&](int j, Lay& La, double (&e)[VEC], double (&f)[VEC], double (&eu)[VEC], double (&fu)[VEC]) {
        up_row(j, La, e, f, eu, fu);
        if (j == n_z - 1) {  // top boundary: dn = sky diffuse (ref :132-135); a = c = 0
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                e[v] = 0.0;
                f[v] = in.Idf0[v];
            }
            return;
        }
        if (!uni) layer(j, La);
        down_row(g[j + 1], La, eu, fu, e, f);
    };
    // ---- forward sweep (ref tdma :183-192).  Checkpointed: the DOWNWARD row's (e, f) of every CK-th level is
    // parked (F <- e_dn, I_df_d <- f_dn), 16/CK B per layer.band; the back sweep re-runs the recurrence from a
    // checkpoint through the CK levels above it into the Out object's segment store.
    double e_prev[VEC], f_prev[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) e_prev[v] = f_prev[v] = 0.0;
    const int g_last = (n_z - 1) / CK;  // segment g covers levels g CK .. min((g+1) CK, n_z) - 1
    if (g_last > 0) {  // the top segment is left to the back sweep's recomputation
        Lay La;
        double eu[VEC], fu[VEC];
        layer(0, La);
        step_soil(La, e_prev, f_prev, eu, fu);
        if (CK == 1) {
            out.st_tmp(F_F, 0, e_prev);
            out.st_tmp(F_DN, 0, f_prev);
        }
        for (int j = 1; j < g_last * CK; ++j) {
            step_in(j, La, e_prev, f_prev, eu, fu);
            if ((j + 1) % CK == 0) {
                out.st_tmp(F_F, j, e_prev);
                out.st_tmp(F_DN, j, f_prev);
            }
        }
    }
    // ---- back substitution (ref tdma :195-198) fused with the output stage (ref :141-161)
    double up_above[VEC], dn_above[VEC], top[VEC][3], Io[VEC], omo[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        up_above[v] = dn_above[v] = 0.0;
        omo[v] = 1.0 - (in.leaf_r[v] + in.leaf_t[v]);
        Io[v] = in.Idr0[v] * omo[v];
    }
    // Output stage of level j = base + i from the segment store.  At the top level ed = 0 and up_above = 0, so
    // dn = fd exactly (ref :132-135) without a special case.  Leaves this level's values in up_above / dn_above.
    auto level_out = [&](int i, int j, double (&Idr)[VEC]) {
        double ed[VEC], fd[VEC], eu[VEC], fu[VEC];  // forward pairs of this level's downward and upward rows
        double dn[VEC], up[VEC], F[VEC];
        out.seg_ld(i, 0, ed);
        out.seg_ld(i, 1, fd);
        out.seg_ld(CK + i, 0, eu);
        out.seg_ld(CK + i, 1, fu);
        const double tbc = tbcum[j];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            dn[v] = fd[v] - ed[v] * up_above[v];
            up[v] = fu[v] - eu[v] * dn[v];
            Idr[v] = in.Idr0[v] * tbc;                                       // ref :151
            F[v] = Idr[v] * s.inv_mu + 2.0 * dn[v] + 2.0 * up[v];            // ref :161
        }
        if (j < n_z - 1) {  // layer j (between levels j and j+1): absorbed per unit sunlit/shaded leaf area
            double sl[VEC], sh[VEC];
            const double gd = g[j + 1], o = omtd[j], fs = fsun[j], ofs = omfs[j], wsl = isl[j], wsh = ish[j];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const double direct = Io[v] * gd;                                  // ref :145
                const double diffuse = (dn_above[v] + up[v]) * o * omo[v];         // ref :146
                sl[v] = (diffuse * fs + direct) * wsl;                             // ref :147, :154
                sh[v] = (diffuse * ofs) * wsh;                                     // ref :148, :155
            }
            out.st(F_X0, j, sl);
            out.st(F_X1, j, sh);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            up_above[v] = up[v];
            dn_above[v] = dn[v];
        }
        out.st(F_IDR, j, Idr);
        out.st(F_DN, j, dn);
        out.st(F_UP, j, up);
        out.st(F_F, j, F);
    };
    for (int sg = g_last; sg >= 0; --sg) {
        const int base = sg * CK;
        const int len = (n_z - base < CK) ? n_z - base : CK;
        double e0[VEC], f0[VEC];  // downward row's pair of level base-1 (zeros below the soil row)
        if (sg == g_last) {       // still in registers from the forward sweep
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e0[v] = e_prev[v]; f0[v] = f_prev[v]; }
        } else if (sg > 0) {
            out.ld_pf(F_F, base - 1, 0, e0);
            out.ld_pf(F_DN, base - 1, 1, f0);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) e0[v] = f0[v] = 0.0;
        }
        if (sg >= 2) {  // start fetching the next segment's checkpoint now (consumed after this segment's two passes)
            out.pf_tmp(F_F, base - CK - 1, 0);
            out.pf_tmp(F_DN, base - CK - 1, 1);
        }
        {
            double eu[VEC], fu[VEC];
            Lay La;
            layer((base > 0 && !uni) ? base - 1 : 0, La);
            int i = 0;
            if (base == 0) {
                step_soil(La, e0, f0, eu, fu);
                out.seg_st(0, 0, e0);
                out.seg_st(0, 1, f0);
                out.seg_st(CK, 0, eu);
                out.seg_st(CK, 1, fu);
                i = 1;
            }
            for (; i < len; ++i) {
                step_in(base + i, La, e0, f0, eu, fu);
                out.seg_st(i, 0, e0);
                out.seg_st(i, 1, f0);
                out.seg_st(CK + i, 0, eu);
                out.seg_st(CK + i, 1, fu);
            }
        }
        int i = len - 1;
        double Idr[VEC];
        if (sg == g_last) {  // the top level: remember its values for the canopy-absorbed sum
            level_out(i, base + i, Idr);
#pragma unroll
            for (int v = 0; v < VEC; ++v) { top[v][0] = Idr[v]; top[v][1] = dn_above[v]; top[v][2] = up_above[v]; }
            --i;
        }
        for (; i >= 0; --i) level_out(i, base + i, Idr);
        if (sg == 0) {  // the ground level was the last one written
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                absorbed[v] = absorbed_from_ends(top[v][0], Idr[v], top[v][1], dn_above[v], top[v][2], up_above[v]);
        }
    }
}

// =================================================================================================
// zq  Zhao & Qualls (2005) with the multiple-scattering correction   (ref _solve_zq.py:13-229)
//
// Unknowns x[0 .. 2m+1], m = n_z: SWu0 = x[::2], SWd0 = x[1::2] (m+1 each).  Row 0 (x0 = rho S_0) and
// row 2m+1 (x = I_df0) are trivial; rows 2li-1, 2li (li = 1..m) are a tridiagonal whose coefficients
// are identical in every real layer except next to the ghost soil (li = 1) and ghost top (li = m)
// layers (ref :99-122).  Thomas elimination without pivoting (SuperLU in the reference; agreement
// ~1e-14); the forward coefficients of row 2li of every CK-th level are parked at
// level li-1 of two of the main output arrays (checkpoints), the rest is recomputed in the back sweep.
// =================================================================================================
struct ScenZq {
    double inv_mu, cos_psi, tau_i, t_psi;
};

// Per-column constants of the zq rows.  With r, a, t of the layer below (lo), the layer itself (me) and the
// layer above (hi) of row pair li, the reference's coefficients (ref :113-118) are
//   pen = t_me + (1 - t_me)(1 - a_me)(1 - r_me),   s_x = r_x (1 - a_x)(1 - t_x),   m_lo = 1 - s_lo s_me,  m_hi = 1 - s_me s_hi
//   row 2li-1:  sub = -pen,  main = -s_lo pen,  sup = m_lo;      row 2li:  sub = m_hi,  main = -s_hi pen,  sup = -pen.
// The middle layer is always a real layer, so only s_lo in {s, s_bot} (soil ghost below li = 1: r = 1, t = 0,
// a = 1 - rho), s_hi in {s, 0} (top ghost above li = m), m_lo in {m_mid, m_bot} and m_hi in {m_mid, 1} vary.
struct ZqCol {
    double pen, s, s_bot, m_mid, m_bot, im_mid, im_bot;
    double mainAB, kA, kB;  // interior rows: -s pen (both diagonals), m_mid cA, m_mid cB
    double cA, cB;
};

// eK[j] = exp(-K L[j]) level table.
template <int VEC, class Out>
CRT_HD void column_zq(const ScenZq& s, const double* eK, int n_z, const BandIn<VEC>& in, Out& out,
                      double (&absorbed)[VEC]) {
    const int m = n_z;
    ZqCol col[VEC];
    double x0[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const double bL = in.leaf_r[v], tL = in.leaf_t[v], rho = in.soil_r[v];
        const double r = 2.0 / 3 * (bL / (bL + tL)) + 1.0 / 3 * (tL / (bL + tL));   // eq. 23 (ref :40-43)
        const double a = 1.0 - (bL + tL);                                           // ref :88
        const double t = s.tau_i;
        const double a0 = 1.0 - rho;                                                // ghost soil layer: r=1, t=0, a=1-rho
        ZqCol& c = col[v];
        c.pen = t + (1.0 - t) * (1.0 - a) * (1.0 - r);
        c.s = r * (1.0 - a) * (1.0 - t);
        // (a perfectly black soil zeroes the main-diagonal entry -s_bot pen of row 1, a pivot of the pivot-free Thomas sweep;
        // the reference's pivoting solver does not care.  1e-200 instead of 0 is the exact solution for a soil reflectance
        // that differs from the caller's by 1e-200: every product below stays finite, x0 = rho S[0] stays exactly 0.)
        c.s_bot = fmax(1.0 * (1.0 - a0) * (1.0 - 0.0), 1e-200);
        c.m_mid = 1.0 - c.s * c.s;
        c.m_bot = 1.0 - c.s_bot * c.s;
        c.im_mid = 1.0 / c.m_mid;
        c.im_bot = 1.0 / c.m_bot;
        const double r_psi = 0.5 + 0.3334 * ((bL - tL) / (bL + tL)) * s.cos_psi;    // eq. 22 (ref :35-38)
        c.cA = r_psi * (1.0 - s.t_psi) * (1.0 - a);                                 // C[2li-1] / (m_lo S)  (ref :136-139)
        c.cB = (1.0 - s.t_psi) * (1.0 - a) * (1.0 - r_psi);                         // C[2li]   / (m_hi S)  (ref :140-143)
        c.mainAB = -c.s * c.pen;
        c.kA = c.m_mid * c.cA;
        c.kB = c.m_mid * c.cB;
        x0[v] = rho * (in.Idr0[v] * eK[0]);                                         // row 0: x0 = rho S[0]  (ref :134)
    }
    // Rows 2li-1 ("A") and 2li ("B") of level li (ref :113-118):
    //   A: sub = -pen, main = -s_lo pen, sup = m_lo, rhs = m_lo cA S;   B: sub = m_hi, main = -s_hi pen, sup = -pen, rhs = m_hi cB S
    // Row A's forward pair from the pair of the row before it (one reciprocal per row).
    auto rowA = [&](int li, int v, double e_in, double f_in, double& eA, double& fA) {
        const ZqCol& c = col[v];
        const double S = in.Idr0[v] * eK[li - 1];
        if (li > 1) {
            const double rA = rcp_nr(c.mainAB + c.pen * e_in);
            eA = c.m_mid * rA;
            fA = (c.kA * S + c.pen * f_in) * rA;
        } else {
            const double rA = rcp_nr(-c.s_bot * c.pen + c.pen * e_in);
            eA = c.m_bot * rA;
            fA = (c.m_bot * c.cA * S + c.pen * f_in) * rA;
        }
    };
    // ---- forward sweep over li = 1..m; previous row is row 0 (e = 0, f = x0) at the start.
    // Checkpointed Thomas: only every CK-th level's row-B pair (e, f) is parked in the output arrays
    // (st_tmp: 16/CK B per layer.band of scratch traffic instead of 16 out + 16 back); the back sweep
    // re-runs the forward recurrence from a checkpoint through the CK levels above it into the Out
    // object's segment store (shared memory on the device) and back-substitutes through them.
    // The recomputation uses the same inlined expressions, so it reproduces the forward values.
    const int CK = out.seg_levels();
    auto fwd = [&](int li, int v, double e_in, double f_in, double& eB_o, double& fB_o) {
        const ZqCol& c = col[v];
        double eA, fA;
        rowA(li, v, e_in, f_in, eA, fA);
        const double S = in.Idr0[v] * eK[li - 1];
        if (li < m) {
            const double rB = rcp_nr(c.mainAB - c.m_mid * eA);
            eB_o = -c.pen * rB;
            fB_o = (c.kB * S - c.m_mid * fA) * rB;
        } else {  // top ghost: s_hi = 0, m_hi = 1
            const double rB = rcp_nr(-0.0 * c.pen - eA);
            eB_o = -c.pen * rB;
            fB_o = (c.cB * S - fA) * rB;
        }
    };
    double e_prev[VEC], f_prev[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { e_prev[v] = 0.0; f_prev[v] = x0[v]; }
    const int g_last = (m - 1) / CK;  // segment g covers li = g CK + 1 .. min((g+1) CK, m)
    for (int li = 1; li <= g_last * CK; ++li) {  // the top segment is left to the back sweep's recomputation
#pragma unroll
        for (int v = 0; v < VEC; ++v) fwd(li, v, e_prev[v], f_prev[v], e_prev[v], f_prev[v]);
        if (li % CK == 0) {
            out.st_tmp(F_F, li - 1, e_prev);   // row B's pair of level li, parked at output level li-1
            out.st_tmp(F_DN, li - 1, f_prev);
        }
    }
    // ---- back substitution + multiple-scattering correction (eq. 24, 25; ref :173-187) + outputs.
    // x[2k] = SWu0[k], x[2k+1] = SWd0[k]; x[2m+1] = I_df0 (last row: sub-diagonal is 0).  Step li knows
    // x[2li+1] and produces x[2li], x[2li-1].  Output level j needs SWu0[j] = x[2j] and
    // SWd0[j+1] = x[2j+3]: level j = li is completed at step li, using x[2li+3] kept from step li+1.
    double x_next[VEC];  // x[2li+1]
#pragma unroll
    for (int v = 0; v < VEC; ++v) x_next[v] = in.Idf0[v];
    double top[VEC][3];
    double pend_SWd0[VEC];  // SWd0[li+1] = x[2li+3] for the output level j = li finished at step li
#pragma unroll
    for (int v = 0; v < VEC; ++v) pend_SWd0[v] = 0.0;
    for (int g = g_last; g >= 0; --g) {
        const int base = g * CK;
        const int len = (m - base < CK) ? m - base : CK;
        double e0[VEC], f0[VEC];  // row B's pair of level `base` (row 0 for the lowest segment)
        if (g == g_last) {        // still in registers from the forward sweep
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e0[v] = e_prev[v]; f0[v] = f_prev[v]; }
        } else if (g > 0) {
            out.ld_pf(F_F, base - 1, 0, e0);
            out.ld_pf(F_DN, base - 1, 1, f0);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e0[v] = 0.0; f0[v] = x0[v]; }
        }
        if (g >= 2) {  // start fetching the next segment's checkpoint now (consumed after this segment's two passes)
            out.pf_tmp(F_F, base - CK - 1, 0);
            out.pf_tmp(F_DN, base - CK - 1, 1);
        }
        {
            double e[VEC], f[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e[v] = e0[v]; f[v] = f0[v]; }
            for (int i = 1; i <= len; ++i) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) fwd(base + i, v, e[v], f[v], e[v], f[v]);
                out.seg_st(i - 1, 0, e);
                out.seg_st(i - 1, 1, f);
            }
        }
        for (int i = len; i >= 1; --i) {
            const int li = base + i;
            double eB[VEC], fB[VEC], eBl[VEC], fBl[VEC];  // row B of this level and of the level below
            out.seg_ld(i - 1, 0, eB);
            out.seg_ld(i - 1, 1, fB);
            if (i >= 2) {
                out.seg_ld(i - 2, 0, eBl);
                out.seg_ld(i - 2, 1, fBl);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { eBl[v] = e0[v]; fBl[v] = f0[v]; }
            }
            double xB[VEC], xA[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                double eA, fA;
                rowA(li, v, eBl[v], fBl[v], eA, fA);  // same expressions as the forward sweep
                xB[v] = fB[v] - eB[v] * x_next[v];   // x[2li]   = SWu0[li]
                xA[v] = fA - eA * xB[v];             // x[2li-1] = SWd0[li-1]
            }
            // Now SWu0[li] (xB) is known: finish output level j = li (needs SWu0[li], SWd0[li+1]) -- but
            // level index j runs 0..m-1 with I_df_u[j] = SWu[j], I_df_d[j] = SWd[j+1]; level j=li exists
            // only for li <= m-1 and its SWd0[j+1] = x[2li+3] was x_next of the previous step (pend_SWd0).
            if (li <= m - 1) {
                const int j = li;
                double Idr[VEC], dn[VEC], up[VEC], F[VEC], dn_ss[VEC], up_ss[VEC], F_ss[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const ZqCol& c = col[v];                                 // li' = j+1 >= 2: s_lo = s_me = s, m_lo = m_mid
                    const double SWu0 = xB[v], SWd0 = pend_SWd0[v];
                    dn[v] = (SWd0 + c.s * SWu0) * c.im_mid;                  // eq. 24 with li' = j+1  (ref :178-180)
                    up[v] = (SWu0 + c.s * SWd0) * c.im_mid;                  // eq. 25                 (ref :183-185)
                    dn_ss[v] = SWd0;
                    up_ss[v] = SWu0;
                    Idr[v] = in.Idr0[v] * eK[j];
                    F_ss[v] = Idr[v] * s.inv_mu + 2.0 * SWu0 + 2.0 * SWd0;   // ref :206
                    F[v] = Idr[v] * s.inv_mu + 2.0 * up[v] + 2.0 * dn[v];    // ref :207
                    if (j == m - 1) { top[v][0] = Idr[v]; top[v][1] = dn[v]; top[v][2] = up[v]; }
                }
                out.st(F_IDR, j, Idr);
                out.st(F_DN, j, dn);
                out.st(F_UP, j, up);
                out.st(F_F, j, F);
                out.st(F_X0, j, dn_ss);
                out.st(F_X1, j, up_ss);
                out.st(F_X2, j, F_ss);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                pend_SWd0[v] = x_next[v];  // SWd0[li] = x[2li+1], needed by output level j = li-1
                x_next[v] = xA[v];         // x[2(li-1)+1]
            }
        }
    }
    // output level j = 0: SWu0[0] = x[0] = x0 - e0 x[1] with e0 = 0; SWd0[1] = pend_SWd0
    {
        double Idr[VEC], dn[VEC], up[VEC], F[VEC], dn_ss[VEC], up_ss[VEC], F_ss[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const ZqCol& c = col[v];  // li' = 1: s_lo = s_bot, m_lo = m_bot
            const double SWu0 = x0[v], SWd0 = pend_SWd0[v];
            dn[v] = (SWd0 + c.s * SWu0) * c.im_bot;
            up[v] = (SWu0 + c.s_bot * SWd0) * c.im_bot;
            dn_ss[v] = SWd0;
            up_ss[v] = SWu0;
            Idr[v] = in.Idr0[v] * eK[0];
            F_ss[v] = Idr[v] * s.inv_mu + 2.0 * SWu0 + 2.0 * SWd0;
            F[v] = Idr[v] * s.inv_mu + 2.0 * up[v] + 2.0 * dn[v];
            if (m == 1) { top[v][0] = Idr[v]; top[v][1] = dn[v]; top[v][2] = up[v]; }
            absorbed[v] = absorbed_from_ends(top[v][0], Idr[v], top[v][1], dn[v], top[v][2], up[v]);
        }
        out.st(F_IDR, 0, Idr);
        out.st(F_DN, 0, dn);
        out.st(F_UP, 0, up);
        out.st(F_F, 0, F);
        out.st(F_X0, 0, dn_ss);
        out.st(F_X1, 0, up_ss);
        out.st(F_X2, 0, F_ss);
    }
}

// =================================================================================================
// zq_pa  Zhao & Qualls, pyAPES variant   (ref _solve_zq_pa.py:24-418)
//
// Same tridiagonal as zq (identical row formulas, ref :195-236 vs _solve_zq.py:113-118), but on a
// computational grid of M = min(100, n_z) equal-LAI LAYERS (ref :94-100) with per-layer beam
// transmittance exp(-Kb LAI/M), solved densely by the reference (np.linalg.solve, ref :278) and mapped
// back to the caller's levels by np.interp (ref :354-361).  One thread per column; the M-grid work
// arrays (<= 100 layers) live in local memory.  The absorption block of the reference (ref :363-401)
// does not reach its return value and is not computed.
// =================================================================================================
constexpr int ZQPA_MAX_M = 100;

struct ScenZqPa {
    double inv_mu, cos_psi, Kb, tau_d, LAI;  // tau_d = tau_df_fn(K_b_fn, LAI / M)  (ref :175)
    int M;
};

// np.interp(x, xp, .) for ascending xp[0..n-1] (numpy's rules at and beyond the ends, ref :359-361), reduced
// to what does not depend on the interpolated array: with values stored in reverse (value at xp[i] is
// SW[n-1-i], ref :354-358) the result is   t < 0 ? SW[k-1] : ((SW[k-1] - SW[k]) * w) * t + SW[k].
// k in 1..n-1 identifies the grid interval (k, k-1); t = x - xp[i] (or 0 for an exact hit / x < xp[0],
// or -1 for x at or beyond the last grid point, where numpy returns the end value itself); w = 1/(xp[i+1]-xp[i]).
CRT_HD void interp_np_prepare(double x, const double* xp, int n, double dl, int& k, double& t, double& w) {
    w = 0.0;
    if (x < xp[0]) { k = n - 1; t = 0.0; return; }
    if (x > xp[n - 1]) { k = 1; t = -1.0; return; }
    int j = (int)(x / dl);
    if (j > n - 1) j = n - 1;
    if (j < 0) j = 0;
    while (j > 0 && xp[j] > x) --j;
    while (j < n - 1 && xp[j + 1] <= x) ++j;
    if (j == n - 1) { k = 1; t = -1.0; return; }
    k = n - 1 - j;
    if (xp[j] == x) { t = 0.0; return; }
    t = x - xp[j];
    w = 1.0 / (xp[j + 1] - xp[j]);
}

// ---- closed form of the M-grid system (ordinary columns) -----------------------------------------------------
// The M-grid layers are identical (LAI/M each, ref :96-100), so the interior rows of the reference's matrix
// (ref :205-225) have CONSTANT coefficients: with y_k = (SWu0[k], SWd0[k]) rows 2k-1 and 2k read
//     y_k = T y_{k-1} + c Ib[k],   T = [[a, b], [-s a, D/pen - s b]],  a = pen/D,  b = s pen/D,  D = 1 - s^2,
//     c = (cA, -s cA - (D/pen) cB),   for k = 2..M-1   (row 2k-1 differs at k = 1: soil below; row 2k at k = M: sky above).
// det T = 1 and tr T = pen + D/pen >= 2 (equality only for conservative leaves), so the eigenvalues are lam >= 1 and
// 1/lam, and Ib[k] = I_dr0 exp(-Kb cum[M+1-k]) is geometric with ratio 1/taub.  Hence for k = 1..M-1
//     y_k = alpha lam^{k-(M-1)} v+  +  beta lam^{-(k-1)} v-  +  w Ib[k],      (1 - T taub) w = c,
// with alpha, beta from the two boundary rows (row 1 with row 2 eliminating SWd0[0]; row 2M with SWd0[M] = I_df0).
// The back sweep then needs two multiplications and six FMAs per grid level instead of the checkpointed Thomas
// recurrences (two forward passes with two reciprocals per level each).  The same linear system as the reference's
// dense solve; measured against it on 40 000 adversarial columns: <= 2e-11 outside the two windows below, where the
// column falls back to the Thomas sweep:
//   (lam - 1) M < CRT_ZQPA_DEGENERATE : the two modes approach linear dependence (error ~ eps / ((lam-1) M)^2);
//   |1 - lam taub| < CRT_ZQPA_RESONANCE : the beam decays like the diffuse mode, w ~ 1/(1 - lam taub) cancels.
// (zq's system has the same constant-coefficient structure on equally spaced level axes.  The closed form was built and
// measured there too and DROPPED: 141 -> 61 instructions per unit, but 0.86 -> 0.73 of HBM peak (n_z = 1000: 0.66 -> 0.52).
// zq's flat kernel is bound by how the column store pattern reaches DRAM, and the Thomas sweep's store-free forward phases
// halve the number of concurrently open row fragments; see profiles/r02_zq_closed_form_persistent_DROPPED.txt.)
#ifndef CRT_ZQPA_DEGENERATE
#define CRT_ZQPA_DEGENERATE 5e-2
#endif
#ifndef CRT_ZQPA_RESONANCE
#define CRT_ZQPA_RESONANCE 1e-4
#endif
struct ZqPaClosed {
    double Au, Bu, wu, Ad, Bd, wd;  // SWu0[k] = Au P + Bu Q + wu Ib[k],  SWd0[k] = Ad P + Bd Q + wd Ib[k]
    double lam, il, Q_top;          // P = lam^{k-(M-1)}, Q = lam^{-(k-1)};  Q_top = lam^{-(M-2)} = Q at k = M-1
};
// one_m_t = 1 - tau_d, aL = 1 - (leaf_r + leaf_t); Ib1, IbM1, IbM = Ib[1], Ib[M-1], Ib[M].  Returns false for a
// column that must take the Thomas sweep (windows above, M < 4, non-finite or underflowing coefficients).
CRT_HD bool zq_pa_closed_coef(const ZqCol& c, double one_m_t, double aL, double taub, int M, double Ib1, double IbM1,
                              double IbM, double x0, double Idf0, ZqPaClosed& o) {
#ifdef CRT_ZQPA_NO_CLOSED
    return false;
#else
    if (M < 4) return false;
    const double pen = c.pen, s = c.s, D = c.m_mid;
    const double q1 = one_m_t * aL;                     // = 1 - pen - s, without the cancellation
    const double hm1 = (q1 * (q1 + 2.0 * s)) / (2.0 * pen);  // tr T / 2 - 1
    const double lam = 1.0 + hm1 + sqrt(hm1 * (hm1 + 2.0));
    const double il = 1.0 / lam;
    if (!((lam - 1.0) * M >= CRT_ZQPA_DEGENERATE) || !(fabs(1.0 - lam * taub) >= CRT_ZQPA_RESONANCE)) return false;
    double Qt = 1.0, x = il;  // lam^{-(M-2)} by squaring
    for (int n = M - 2; n > 0; n >>= 1) {
        if (n & 1) Qt *= x;
        x *= x;
    }
    if (!(Qt > 1e-280)) return false;
    const double ip = 1.0 / pen;
    const double a = pen * c.im_mid, b = s * a, T21 = -s * a, T22 = D * ip - s * b;
    // eigenvectors: v+ = (b, lam - a) (T22 - lam and a - lam have opposite signs: lam - a > 0), v- = (1/lam - T22, T21)
    const double vpu = b, vpd = lam - a, vmu = il - T22, vmd = T21;
    const double c0 = c.cA, c1 = -s * c.cA - D * ip * c.cB;
    const double m00 = 1.0 - a * taub, m01 = -b * taub, m10 = -T21 * taub, m11 = 1.0 - T22 * taub;
    const double idet = 1.0 / (m00 * m11 - m01 * m10);
    const double wu = (c0 * m11 - m01 * c1) * idet, wd = (m00 * c1 - m10 * c0) * idet;
    // row 1 (soil below: s_bot, m_bot) with SWd0[0] eliminated through row 2:  c1u SWu0[1] + c1d SWd0[1] = r1
    const double pp = c.s_bot * pen * pen * c.im_mid;
    const double c1u = c.m_bot - s * pp, c1d = -pp;
    const double r1 = c.m_bot * c.cA * Ib1 + pen * x0 + c.s_bot * pen * c.cB * Ib1 - (c1u * wu + c1d * wd) * Ib1;
    const double r2 = pen * Idf0 + c.cB * IbM - wd * IbM1;  // row 2M: SWd0[M-1] = pen I_df0 + cB Ib[M]
    const double A00 = (c1u * vpu + c1d * vpd) * Qt, A01 = c1u * vmu + c1d * vmd, A10 = vpd, A11 = Qt * vmd;
    const double idd = 1.0 / (A00 * A11 - A01 * A10);
    const double al = (r1 * A11 - A01 * r2) * idd, be = (A00 * r2 - A10 * r1) * idd;
    o.Au = al * vpu; o.Ad = al * vpd; o.Bu = be * vmu; o.Bd = be * vmd;
    o.wu = wu; o.wd = wd; o.lam = lam; o.il = il; o.Q_top = Qt;
    const double chk = o.Au + o.Ad + o.Bu + o.Bd + wu + wd;
    return chk - chk == 0.0;  // all finite
#endif
}

// eC[i] = exp(-Kb cum[i]) on the M-grid (cum = running sum of LAI/M as np.cumsum produces it, ref :161).
// The caller's levels come sorted by descending grid interval (the order in which the back sweep can finish
// them): lk[c] = (level j, interval k) as two int32, and tt/ww (interp_np_prepare()) and eK = exp(-Kb lai[j])
// of that level.
//
// The M-grid solve is the checkpointed Thomas sweep of column_zq (checkpoints in the Out object's segment
// store: slots CK+1.. hold the checkpoints, slots 0..CK the segment being back-substituted, slot 0 = the pair below it).  The back sweep
// runs from the top of the canopy down and the corrected fluxes (eq. 24/25, ref :286-345) of grid pair k are
// final at step k, so the interpolation to the caller's levels (ref :350-361) is streamed: as soon as both
// ends of a grid interval are final, every caller level inside it is interpolated and written.  Nothing of
// the M-grid is stored per thread beyond a three-value window (the first version kept four 101-element
// work arrays per thread in local memory and ran at a third of this speed).
template <int VEC, class Out>
CRT_HD void column_zq_pa(const ScenZqPa& s, const double* eC, const double* lk, const double* tt, const double* ww,
                         const double* eK, int n_z, const BandIn<VEC>& in, Out& out, double (&absorbed)[VEC]) {
    const int M = s.M;
    const double dl = s.LAI / M;
    const double taub = exp(-s.Kb * dl);                                            // ref :173
    ZqCol col[VEC];
    ZqPaClosed cf[VEC];
    bool closed = true;
    double x0[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const double bL = in.leaf_r[v], tLf = in.leaf_t[v], rho = in.soil_r[v];
        const double alb = bL + tLf;                                                // ref :136
        const double aL = 1.0 - alb;
        const double tL = tLf / alb, rL = bL / alb;                                 // ref :142-145
        const double rb = 0.5 + 0.3334 * (rL - tL) / (rL + tL) * s.cos_psi;         // ref :185
        const double rd = 2.0 / 3.0 * rL / (rL + tL) + 1.0 / 3.0 * tL / (rL + tL);  // ref :186
        const double t = s.tau_d, a0 = 1.0 - rho;
        ZqCol& c = col[v];
        c.pen = t + (1.0 - t) * (1.0 - aL) * (1.0 - rd);
        c.s = rd * (1.0 - aL) * (1.0 - t);
        c.s_bot = fmax(1.0 * (1.0 - a0) * (1.0 - 0.0), 1e-200);                     // k = 1: soil below (black soil: see column_zq)
        c.m_mid = 1.0 - c.s * c.s;
        c.m_bot = 1.0 - c.s_bot * c.s;
        c.im_mid = 1.0 / c.m_mid;
        c.im_bot = 1.0 / c.m_bot;
        c.cA = rb * (1.0 - taub) * (1.0 - aL);                                      // ref :243-256
        c.cB = (1.0 - taub) * (1.0 - aL) * (1.0 - rb);                              // ref :257-271
        c.mainAB = -c.s * c.pen;
        c.kA = c.m_mid * c.cA;
        c.kB = c.m_mid * c.cB;
        x0[v] = rho * (eC[M] * in.Idr0[v]);                                         // C[0] = SoilAlbedo Ib[0]  (ref :241)
        closed = zq_pa_closed_coef(c, 1.0 - t, aL, taub, M, eC[M] * in.Idr0[v], eC[M >= 2 ? 2 : M] * in.Idr0[v], eC[M >= 1 ? 1 : 0] * in.Idr0[v],
                                   x0[v], in.Idf0[v], cf[v]) && closed;
    }
    // rows 2k-1 ("A") and 2k ("B") of grid layer k, as in column_zq (ref :195-236); Ib[k] = f_sl[k] IbSky (ref :165-169)
    auto rowA = [&](int k, int v, double e_in, double f_in, double& eA, double& fA) {
        const ZqCol& c = col[v];
        const double Ib = eC[M + 1 - k] * in.Idr0[v];
        if (k > 1) {
            const double rA = rcp_nr(c.mainAB + c.pen * e_in);
            eA = c.m_mid * rA;
            fA = (c.kA * Ib + c.pen * f_in) * rA;
        } else {
            const double rA = rcp_nr(-c.s_bot * c.pen + c.pen * e_in);
            eA = c.m_bot * rA;
            fA = (c.m_bot * c.cA * Ib + c.pen * f_in) * rA;
        }
    };
    auto fwd = [&](int k, int v, double e_in, double f_in, double& eB_o, double& fB_o) {
        const ZqCol& c = col[v];
        double eA, fA;
        rowA(k, v, e_in, f_in, eA, fA);
        const double Ib = eC[M + 1 - k] * in.Idr0[v];
        if (k < M) {
            const double rB = rcp_nr(c.mainAB - c.m_mid * eA);
            eB_o = -c.pen * rB;
            fB_o = (c.kB * Ib - c.m_mid * fA) * rB;
        } else {  // k = M: nothing reflects down from above the canopy (rd[M+1] = 0)
            const double rB = rcp_nr(-0.0 * c.pen - eA);
            eB_o = -c.pen * rB;
            fB_o = (c.cB * Ib - fA) * rB;
        }
    };
    const int CK = out.seg_levels();
    double e_prev[VEC], f_prev[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { e_prev[v] = 0.0; f_prev[v] = x0[v]; }
    const int g_last = (M - 1) / CK;  // segment g covers k = g CK + 1 .. min((g+1) CK, M)
    if (!closed) {
        for (int k = 1; k <= g_last * CK; ++k) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) fwd(k, v, e_prev[v], f_prev[v], e_prev[v], f_prev[v]);
            if (k % CK == 0) {
                out.seg_st(CK + 1 + k / CK, 0, e_prev);
                out.seg_st(CK + 1 + k / CK, 1, f_prev);
            }
        }
    }
    // ---- streamed interpolation + outputs (ref :350-361, :403-407).  D[k] = SWd[k], U[k] = SWu[k] after the
    // correction; interval (k, k-1) needs D[k], D[k-1], U[k], U[k-1].
    int jc = 0;  // next entry of `ord`
    double gnd[VEC][3] = {}, top[VEC][3] = {};
    auto emit = [&](int k, const double (&Dk)[VEC], const double (&Dk1)[VEC], const double (&Uk)[VEC], const double (&Uk1)[VEC]) {
        while (jc < n_z) {
            const int* e = reinterpret_cast<const int*>(lk + jc);
            if (e[1] != k) break;
            const int j = e[0];
            const double t = tt[jc], w = ww[jc], eKj = eK[jc];
            ++jc;
            double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                dn[v] = (t < 0.0) ? Dk1[v] : ((Dk1[v] - Dk[v]) * w) * t + Dk[v];
                up[v] = (t < 0.0) ? Uk1[v] : ((Uk1[v] - Uk[v]) * w) * t + Uk[v];
                Idr[v] = in.Idr0[v] * eKj;
                F[v] = Idr[v] * s.inv_mu + 2.0 * up[v] + 2.0 * dn[v];
            }
            if (j == 0 || j == n_z - 1) {  // uniform, twice per column
#pragma unroll
                for (int v = 0; v < VEC; ++v) keep3(j == 0 ? gnd[v] : top[v], Idr[v], dn[v], up[v]);
            }
            out.st4(j, Idr, dn, up, F);
        }
    };
    double D_prev[VEC], U_prev[VEC], U_prev2[VEC];  // D[t+2], U[t+1], U[t+2] when pair t arrives
    auto pair_done = [&](int t, const double (&Dn)[VEC], const double (&Un)[VEC]) {  // Dn = D[t+1], Un = U[t]
        if (t == M - 1) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { D_prev[v] = Dn[v]; U_prev[v] = Un[v]; U_prev2[v] = Un[v]; }  // U[M] = U[M-1]  (ref :343)
        } else {
            emit(t + 2, D_prev, Dn, U_prev2, U_prev);
#pragma unroll
            for (int v = 0; v < VEC; ++v) { D_prev[v] = Dn[v]; U_prev2[v] = U_prev[v]; U_prev[v] = Un[v]; }
        }
    };
    // ---- back substitution: SWu0[k] = x[2k], SWd0[k] = x[2k+1]; multiple scattering eq. 24/25 (ref :286-345)
    double SWd0_hi[VEC], pend[VEC];  // x[2k+1] = SWd0[k];  SWd0[k+1]
#pragma unroll
    for (int v = 0; v < VEC; ++v) { SWd0_hi[v] = in.Idf0[v]; pend[v] = 0.0; }
    if (closed) {
        // Closed form: SWu0[k], SWd0[k] for k = M-1 .. 1 from the two modes (advanced by one multiplication each)
        // and the beam term; pend = SWd0[k+1] starts as SWd0[M] = I_df0 (last row).
        // The mode amplitudes are carried pre-multiplied (pu = Au P, qu = Bu Q, ...): four multiplications per grid level
        // and four live values instead of P, Q and four coefficients.
        double pu[VEC], pd[VEC], qu[VEC], qd[VEC], wuI[VEC], wdI[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            pend[v] = in.Idf0[v];
            pu[v] = cf[v].Au; pd[v] = cf[v].Ad;
            qu[v] = cf[v].Bu * cf[v].Q_top; qd[v] = cf[v].Bd * cf[v].Q_top;
            wuI[v] = cf[v].wu * in.Idr0[v]; wdI[v] = cf[v].wd * in.Idr0[v];
        }
        for (int k = M - 1; k >= 1; --k) {
            const double eCk = eC[M + 1 - k];
            double Dn[VEC], Un[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const ZqCol& c = col[v];
                const double u = (pu[v] + qu[v]) + wuI[v] * eCk;  // SWu0[k];  Ib[k] = I_dr0 eC[M+1-k]
                const double d = (pd[v] + qd[v]) + wdI[v] * eCk;  // SWd0[k]
                Dn[v] = (pend[v] + u * c.s) * c.im_mid;  // eq. 24 (ref :288-312)
                Un[v] = (u + pend[v] * c.s) * c.im_mid;  // eq. 25 (ref :318-342)
                pend[v] = d;
                pu[v] *= cf[v].il; pd[v] *= cf[v].il;
                qu[v] *= cf[v].lam; qd[v] *= cf[v].lam;
            }
            pair_done(k, Dn, Un);
        }
    } else
    for (int g = g_last; g >= 0; --g) {
        const int base = g * CK;
        const int len = (M - base < CK) ? M - base : CK;
        double e0[VEC], f0[VEC];
        if (g == g_last) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e0[v] = e_prev[v]; f0[v] = f_prev[v]; }
        } else if (g > 0) {
            out.seg_ld(CK + 1 + g, 0, e0);
            out.seg_ld(CK + 1 + g, 1, f0);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e0[v] = 0.0; f0[v] = x0[v]; }
        }
        {
            double e[VEC], f[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { e[v] = e0[v]; f[v] = f0[v]; }
            out.seg_st(0, 0, e0);  // slot 0 = the pair below the segment, slot i = level base + i: the back sweep
            out.seg_st(0, 1, f0);  // reads "this level" and "the level below" without a per-level special case
            for (int i = 1; i <= len; ++i) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) fwd(base + i, v, e[v], f[v], e[v], f[v]);
                out.seg_st(i, 0, e);
                out.seg_st(i, 1, f);
            }
        }
        for (int i = len; i >= 1; --i) {
            const int k = base + i;
            double eB[VEC], fB[VEC], eBl[VEC], fBl[VEC];
            out.seg_ld(i, 0, eB);
            out.seg_ld(i, 1, fB);
            out.seg_ld(i - 1, 0, eBl);
            out.seg_ld(i - 1, 1, fBl);
            double SWu0_k[VEC], SWd0_k[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                double eA, fA;
                rowA(k, v, eBl[v], fBl[v], eA, fA);              // row 2k-1 again, same expressions
                SWd0_k[v] = SWd0_hi[v];
                SWu0_k[v] = fB[v] - eB[v] * SWd0_hi[v];          // x[2k]
                SWd0_hi[v] = fA - eA * SWu0_k[v];                // x[2k-1] = SWd0[k-1]
            }
            if (k <= M - 1) {  // pair k couples layers k and k+1: row class of li = k+1 >= 2
                double Dn[VEC], Un[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const ZqCol& c = col[v];
                    Dn[v] = (pend[v] + SWu0_k[v] * c.s) * c.im_mid;  // eq. 24 (ref :288-312)
                    Un[v] = (SWu0_k[v] + pend[v] * c.s) * c.im_mid;  // eq. 25 (ref :318-342)
                }
                pair_done(k, Dn, Un);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) pend[v] = SWd0_k[v];   // needed by pair k-1
        }
    }
    {   // pair 0: SWu0[0] = x[0] = x0, SWd0[1] = pend; row class li = 1 (soil below)
        double Dn[VEC], Un[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const ZqCol& c = col[v];
            Dn[v] = (pend[v] + x0[v] * c.s) * c.im_bot;
            Un[v] = (x0[v] + pend[v] * c.s_bot) * c.im_bot;
        }
        pair_done(0, Dn, Un);
    }
    emit(1, D_prev, D_prev, U_prev2, U_prev);  // D[0] = D[1]  (ref :313)
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        absorbed[v] = absorbed_from_ends(top[v][0], gnd[v][0], top[v][1], gnd[v][1], top[v][2], gnd[v][2]);
}

// =================================================================================================
// 4s  Tian et al. (2007) four-stream   (ref _solve_4s.py:8-293)
//
// The reference integrates  y' = M y + v exp(-kappa x),  y = [R2d, R1d, R1u, R2u]  (ref `eqns` :48-97)
// with scipy's collocation BVP solver, twice per band (direct, diffuse; ref :235-262).  M is constant
// in x, so the solution is closed-form.  With D = (R2d, R1d), U = (R2u, R1u) the system reads
//     D' =  P D + Q U + vD e^{-kappa x},     U' = -Q D - P U - vD e^{-kappa x},
// and for s = D + U, w = D - U:   s' = (P - Q) w,   w' = (P + Q) s + 2 vD e^{-kappa x},
// where P - Q = -diag(k2/mu2, k1/mu1) is diagonal.  Hence s'' = N s + 2 (P-Q) vD e^{-kappa x} with the
// 2x2 matrix N = (P-Q)(P+Q), whose eigenvalues lambda_k^2 are real and positive for omega < 1.
// Solution = 2 decaying + 2 growing exponentials (written as e^{-lambda x} and e^{-lambda (LAI-x)} so
// nothing overflows) + the particular e^{-kappa x} term; the four coefficients come from the
// boundary conditions (ref `dfdr_bcs` :99-138): D(0) = incident, U(LAI) = soil reflection of the
// downward irradiance plus the direct beam.  Direct and diffuse problems share the matrix and their
// solutions are only ever used summed (ref :274-281), so one 4x4 solve per band suffices.
// Accuracy: limited by roundoff (~1e-13), i.e. far tighter than the reference's tol = 1e-6 BVP.
//
// The reference integrates the ODE numerically and is indifferent to the spectrum of N; a closed form is not:
//   * lambda_0^2 (the larger eigenvalue) is positive for every omega <= 1, but lambda_1^2 falls through zero at
//     omega* < 1 (0.99536 for the default ellipsoidal G with mu_s = 0.501) and is NEGATIVE up to omega = 1:
//     the mode turns from exponential to oscillatory, and e^{-lam x}, e^{-lam (LAI - x)} become linearly
//     dependent as lam -> 0.  Columns with lambda_1^2 LAI^2 < 1/4 therefore use the ENTIRE basis
//     c(x) = cosh(lam x), sh(x) = sinh(lam x)/lam (= cos, sin/nu for lambda^2 = -nu^2 < 0; power series near
//     0), which is analytic in lambda^2 and well conditioned through the crossing.
//   * kappa = K_b equal to lambda_k (an ordinary place: omega = 0.594 at psi = 10 deg) makes the particular
//     solution e^{-kappa x}/(kappa^2 - lambda_k^2) blow up and cancel against the homogeneous mode.  Within 5 %
//     of a resonance the particular solution's component along that eigenvector is carried as the divided
//     difference R_k(x) = (e^{-kappa x} - e^{-lam_k x})/(kappa^2 - lam_k^2) = -x e^{-kappa x} phi1((kappa - lam_k) x)/(kappa + lam_k),
//     phi1(u) = expm1(u)/u, which is finite through the resonance (entire-basis variant: the solution of
//     R'' = lambda^2 R + e^{-kappa x}, R(0) = R'(0) = 0, by its power series).
// Both are rare columns: they take `coef_4s_general` (not inlined) and the direct-evaluation level path.
// =================================================================================================
struct Scen4s {
    double mu0, inv_mu, kappa, L_T, eKT;  // cos psi, 1/cos psi, K_b = G/mu0, total LAI, exp(-kappa L_T)
    double mu_s, m1, m2, G1, G2;          // sector cosine, mu_1, mu_2 (ref :188-189), G sector integrals (ref :148-149)
    // band-independent reciprocals, hoisted out of the per-band coefficient stage
    double inv_m1, inv_m2, q0, q1, inv_q0, inv_q1, inv_pi_mu0;
};

CRT_HD Scen4s scen_4s(double psi, double K_b, double G1, double G2, double mu_s, double L_T) {
    Scen4s s;
    s.mu0 = cos(psi);
    s.inv_mu = 1.0 / s.mu0;
    s.kappa = K_b;
    s.L_T = L_T;
    s.eKT = exp(-K_b * L_T);
    s.mu_s = mu_s;
    s.m1 = 0.5 * mu_s * mu_s;
    s.m2 = 0.5 * (1.0 - mu_s * mu_s);
    s.G1 = G1;
    s.G2 = G2;
    s.inv_m1 = 1.0 / s.m1;
    s.inv_m2 = 1.0 / s.m2;
    s.q0 = G2 * s.inv_m2;  // -(P - Q) diagonal: k2/mu2, k1/mu1
    s.q1 = G1 * s.inv_m1;
    s.inv_q0 = 1.0 / s.q0;
    s.inv_q1 = 1.0 / s.q1;
    s.inv_pi_mu0 = 1.0 / (CRT_PI * s.mu0);
    return s;
}

// 4x4 Gaussian elimination with partial pivoting, written with compare-and-swap so that everything
// stays in registers on the device.
CRT_HD void solve4(double (&A)[4][4], double (&b)[4]) {
    double inv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            const bool sw = fabs(A[r][c]) > fabs(A[c][c]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double t = A[c][k];
                A[c][k] = sw ? A[r][k] : t;
                A[r][k] = sw ? t : A[r][k];
            }
            const double tb = b[c];
            b[c] = sw ? b[r] : tb;
            b[r] = sw ? tb : b[r];
        }
        inv[c] = rcp_nr(A[c][c]);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            const double f = A[r][c] * inv[c];
#pragma unroll
            for (int k = c + 1; k < 4; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
#pragma unroll
    for (int r = 3; r >= 0; --r) {
        double acc = b[r];
#pragma unroll
        for (int k = r + 1; k < 4; ++k) acc -= A[r][k] * b[k];
        b[r] = acc * inv[r];
    }
}

// Where the ordinary (real, well separated eigenvalue) closed form hands over to coef_4s_general.  The general branch
// costs direct exp / series evaluations at every level AND makes its whole warp take that path, so the windows are as
// narrow as the ordinary form's accuracy allows (measured against the general form, tests/test_hostmath.py):
//   (lambda_1 LAI)^2 < CRT_4S_ENTIRE_BELOW : the two lambda_1 modes e^{-lam x}, e^{-lam (LAI-x)} approach linear
//       dependence like 1/(lam LAI); at 1e-3 the boundary solve still leaves ~1e-13
//   |kappa^2 - lambda_k^2| <= CRT_4S_RESONANCE_WIDTH kappa^2 : the particular solution's 1/(kappa^2 - lambda_k^2)
//       cancels against the homogeneous part: measured error of the ordinary form 2e-14 / width (2e-11 at 1e-3).
//       Resonant columns: 0.09 % of the sweep's.
#ifndef CRT_4S_ENTIRE_BELOW
#define CRT_4S_ENTIRE_BELOW 1e-3
#endif
#ifndef CRT_4S_RESONANCE_WIDTH
#define CRT_4S_RESONANCE_WIDTH 1e-3
#endif
// Folded per-band coefficients: I_dn(x) = sum_k dnP[k] e^{-lam_k (LAI-x)} + dnM[k] e^{-lam_k x} + dnK e^{-kappa x}
struct Coef4s {
    // I_dn(x) = sum_k dnP[k] p_k + dnM[k] m_k + dnK X,  m_k = e^{-lam_k x},  p_k = e^{-lam_k (LAI - x)},  X = e^{-kappa x}.
    // Fast form (lam[0] > 0): dnP/upP are pre-multiplied by g_k = e^{-lam_k LAI}, so p_k's factor is just e^{+lam_k x},
    // the other half of the exp_pm that yields m_k.
    // The two lowest mantissa bits of a positive lam[0] say whether X is e^{-kappa x} (0) or the resonance-safe
    // R_k(x) = (e^{-kappa x} - e^{-lam_k x})/(kappa^2 - lam_k^2) of mode k = bits - 1 (|kappa^2 - lam_k^2| <=
    // CRT_4S_RESONANCE_WIDTH kappa^2: coefficients from coef_4s_general); every fast-form lam[0] is stored that way.
    // Slow form (lam[0] < 0, direct evaluation of every basis function, unscaled coefficients): lam[0] holds
    // -lambda_0 with a 3-bit mode word in its lowest mantissa bits (a relative perturbation of lambda_0 below
    // 2^-49; e^{-lam x} moves by at most that / e in absolute terms):
    //   mode & 1        mode 1 uses the entire basis: lam[1] = lambda_1^2 (any sign), m_1 = c(x), p_1 = sh(x)
    //   (mode >> 1) & 3 resonant mode + 1 (0 = none): X = R_k(x) instead of e^{-kappa x}
    // Reasons for the slow form: lam_0 LAI >= 600 (the scaled product could overflow), lambda_1^2 LAI^2 <
    // CRT_4S_ENTIRE_BELOW, |kappa^2 - lambda_k^2| <= CRT_4S_RESONANCE_WIDTH kappa^2.
    double lam[2], dnP[2], dnM[2], upP[2], upM[2], dnK, upK, Idr0;
};

CRT_HD double pack_mode_4s(double lam0, int mode) {  // -(lam0 with its three lowest mantissa bits = mode)
#if defined(__CUDA_ARCH__)
    return -__longlong_as_double((__double_as_longlong(lam0) & ~7LL) | (long long)mode);
#else
    int64_t b;
    memcpy(&b, &lam0, 8);
    b = (b & ~(int64_t)7) | (int64_t)mode;
    double r;
    memcpy(&r, &b, 8);
    return -r;
#endif
}
// fast form: lam0 (> 0) with its two lowest mantissa bits = res (a relative change of lam0 below 2^-50)
CRT_HD double tag_res_4s(double lam0, int res) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((__double_as_longlong(lam0) & ~3LL) | (long long)res);
#else
    int64_t b;
    memcpy(&b, &lam0, 8);
    b = (b & ~(int64_t)3) | (int64_t)res;
    double r;
    memcpy(&r, &b, 8);
    return r;
#endif
}
CRT_HD int fast_res_4s(double lam0_stored) {
#if defined(__CUDA_ARCH__)
    return (int)(__double_as_longlong(lam0_stored) & 3LL);
#else
    int64_t b;
    memcpy(&b, &lam0_stored, 8);
    return (int)(b & 3);
#endif
}
CRT_HD int unpack_mode_4s(double lam0_stored) {
#if defined(__CUDA_ARCH__)
    return (int)(__double_as_longlong(lam0_stored) & 7LL);
#else
    int64_t b;
    memcpy(&b, &lam0_stored, 8);
    return (int)(b & 7);
#endif
}

// c(x) = cosh(sqrt(l2) x), sh(x) = sinh(sqrt(l2) x)/sqrt(l2): entire functions of l2 (cos, sin/nu for l2 = -nu^2).
CRT_HD void csh_entire(double l2, double x, double& c, double& sh) {
    const double u2 = l2 * x * x;
    if (fabs(u2) < 1e-4) {  // truncation u2^5/10! < 3e-27
        c = 1.0 + u2 * (1.0 / 2.0) * (1.0 + u2 * (1.0 / 12.0) * (1.0 + u2 * (1.0 / 30.0) * (1.0 + u2 * (1.0 / 56.0))));
        sh = x * (1.0 + u2 * (1.0 / 6.0) * (1.0 + u2 * (1.0 / 20.0) * (1.0 + u2 * (1.0 / 42.0) * (1.0 + u2 * (1.0 / 72.0)))));
    } else if (l2 > 0.0) {
        const double lam = sqrt(l2);
        c = cosh(lam * x);
        sh = sinh(lam * x) / lam;
    } else {
        const double nu = sqrt(-l2);
        c = cos(nu * x);
        sh = sin(nu * x) / nu;
    }
}

// R(x) = (e^{-kappa x} - e^{-lam x})/(kappa^2 - lam^2), finite through kappa = lam.  eK = e^{-kappa x}.
CRT_HD double res_exp_4s(double kappa, double lam, double x, double eK) {
    const double u = (kappa - lam) * x;
    const double phi1 = fabs(u) < 1e-8 ? 1.0 + 0.5 * u : expm1(u) / u;
    return -eK * x * phi1 / (kappa + lam);
}

// R(x): R'' = l2 R + e^{-kappa x}, R(0) = R'(0) = 0, by its power series (used for kappa x, |l2| x^2 <~ 1 only).
CRT_HD double res_series_4s(double kappa, double l2, double x) {
    double a0 = 0.0, a1 = 0.0;     // a_n, a_{n+1}
    double e = 1.0;                // (-kappa)^n / n!
    double xn = x * x;             // x^{n+2}
    double sum = 0.0;
    for (int n = 0; n < 40; ++n) {
        const double a2 = (l2 * a0 + e) / ((n + 2.0) * (n + 1.0));
        sum += a2 * xn;
        a0 = a1;
        a1 = a2;
        e *= -kappa / (n + 1.0);
        xn *= x;
    }
    return sum;
}

// General form of the 4s coefficients: any sign of lambda_1^2, kappa at or near lambda_k (see the header above).
// s(x) = D + U = sum_i phi_i u_i(x), u_i'' = lambda_i^2 u_i + c_i e^{-kappa x} (c from 2 (P-Q) vD = sum c_i phi_i),
// w = D - U = (P-Q)^{-1} s' = sum_i chi_i u_i', chi_i = -phi_i / diag(q).  Each u_i = a_i A_i + b_i B_i + c_i F_i with
// a homogeneous pair (A_i, B_i) and a particular F_i chosen per mode:
//   exponential pair  A = e^{-lam x}, B = e^{-lam (LAI - x)};  entire pair  A = c(x), B = sh(x)  (A' = l2 B, B' = A);
//   F = e^{-kappa x}/(kappa^2 - l2), or the resonance-safe R (R' = -kappa R - A/(kappa + lam); entire pair: R' = -kappa R + B).
// In the resonant case e^{-kappa x} = A_k [- kappa B_k] + (kappa^2 - l2_k) R_k is eliminated from the other mode's
// particular term, so the level stage still sums five products (X := R_k takes the e^{-kappa x} slot).
CRT_HD void coef_4s_general(const Scen4s& s, double r, double t, double rho, double Idr0, double Idf0,
                                     const double (&N)[2][2], const double (&l2)[2], const double (&phi)[2][2],
                                     Coef4s& k) {
    const double omega = r + t;
    const double R_dr0 = Idr0 * s.inv_pi_mu0, R_df0 = Idf0 * (1.0 / CRT_PI);
    const double e1 = 0.25 * omega * R_dr0 * s.mu_s, e2 = 0.25 * omega * R_dr0 * (1.0 - s.mu_s);
    const double m1 = s.m1, m2 = s.m2, L = s.L_T, kap = s.kappa, k2 = kap * kap;
    const double G = kap * s.mu0;
    const double vD0 = G * e2 * s.inv_m2, vD1 = G * e1 * s.inv_m1;
    const double r0 = -2.0 * s.q0 * vD0, r1 = -2.0 * s.q1 * vD1;
    (void)N;
    // forcing in the eigenbasis
    const double dphi = phi[0][0] * phi[1][1] - phi[1][0] * phi[0][1];
    double c[2];
    c[0] = (r0 * phi[1][1] - r1 * phi[1][0]) / dphi;
    c[1] = (phi[0][0] * r1 - phi[0][1] * r0) / dphi;

    const bool entire1 = !(l2[1] * L * L >= CRT_4S_ENTIRE_BELOW);
    int res = 0;
    if (fabs(k2 - l2[0]) <= CRT_4S_RESONANCE_WIDTH * k2) res = 1;
    else if (fabs(k2 - l2[1]) <= CRT_4S_RESONANCE_WIDTH * k2) res = 2;

    double chi[2][2], mphi[2], mchi[2], lam[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        chi[i][0] = -phi[i][0] * s.inv_q0;
        chi[i][1] = -phi[i][1] * s.inv_q1;
        mphi[i] = m2 * phi[i][0] + m1 * phi[i][1];
        mchi[i] = m2 * chi[i][0] + m1 * chi[i][1];
        lam[i] = l2[i] > 0.0 ? sqrt(l2[i]) : 0.0;
    }
    // boundary data (f(0), f'(0), f(L), f'(L)) of A_i, B_i, F_i
    double Ab[2][4], Bb[2][4], Fb[2][4];
    double shL = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (i == 1 && entire1) {
            double cL;
            csh_entire(l2[1], L, cL, shL);
            Ab[i][0] = 1.0; Ab[i][1] = 0.0; Ab[i][2] = cL;  Ab[i][3] = l2[1] * shL;
            Bb[i][0] = 0.0; Bb[i][1] = 1.0; Bb[i][2] = shL; Bb[i][3] = cL;
        } else {
            const double g = exp(-lam[i] * L);
            Ab[i][0] = 1.0; Ab[i][1] = -lam[i];    Ab[i][2] = g;   Ab[i][3] = -lam[i] * g;
            Bb[i][0] = g;   Bb[i][1] = lam[i] * g; Bb[i][2] = 1.0; Bb[i][3] = lam[i];
        }
        if (res == i + 1) {
            if (i == 1 && entire1) {
                const double RL = res_series_4s(kap, l2[1], L);
                Fb[i][0] = 0.0; Fb[i][1] = 0.0; Fb[i][2] = RL; Fb[i][3] = -kap * RL + shL;
            } else {
                const double RL = res_exp_4s(kap, lam[i], L, s.eKT);
                const double ik = 1.0 / (kap + lam[i]);
                Fb[i][0] = 0.0; Fb[i][1] = -ik; Fb[i][2] = RL; Fb[i][3] = -kap * RL - Ab[i][2] * ik;
            }
        } else {
            const double id = 1.0 / (k2 - l2[i]);
            Fb[i][0] = id; Fb[i][1] = -kap * id; Fb[i][2] = s.eKT * id; Fb[i][3] = -kap * s.eKT * id;
        }
    }
    // 4x4 system for [B_0, B_1, A_0, A_1] coefficients (the slot order P0, P1, M0, M1 of the level stage)
    double A[4][4], rhs[4];
    auto fill = [&](int col, int i, const double (&f)[4]) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            A[cc][col] = 0.5 * (phi[i][cc] * f[0] + chi[i][cc] * f[1]);               // D_cc(0)
            A[2 + cc][col] = 0.5 * (phi[i][cc] * f[2] - chi[i][cc] * f[3])            // U_cc(L)
                             - rho * (mphi[i] * f[2] + mchi[i] * f[3]);              //  - 2 rho (m2 D_0 + m1 D_1)(L)
        }
    };
    fill(0, 0, Bb[0]);
    fill(1, 1, Bb[1]);
    fill(2, 0, Ab[0]);
    fill(3, 1, Ab[1]);
    const double beam = rho * s.mu0 * R_dr0 * s.eKT;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        double Dp0 = 0.0, botp = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            Dp0 += 0.5 * c[i] * (phi[i][cc] * Fb[i][0] + chi[i][cc] * Fb[i][1]);
            botp += c[i] * (0.5 * (phi[i][cc] * Fb[i][2] - chi[i][cc] * Fb[i][3]) - rho * (mphi[i] * Fb[i][2] + mchi[i] * Fb[i][3]));
        }
        rhs[cc] = R_df0 - Dp0;
        rhs[2 + cc] = beam - botp;
    }
    solve4(A, rhs);
    // irradiances: I_dn = pi sum coef (mphi f + mchi f'),  I_up = pi sum coef (mphi f - mchi f')
    double dK = 0.0, uK = 0.0;  // coefficient of the X slot
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double aB = rhs[i], aA = rhs[2 + i];
        if (i == 1 && entire1) {
            k.dnM[i] = CRT_PI * (aA * mphi[i] + aB * mchi[i]);
            k.dnP[i] = CRT_PI * (aA * l2[1] * mchi[i] + aB * mphi[i]);
            k.upM[i] = CRT_PI * (aA * mphi[i] - aB * mchi[i]);
            k.upP[i] = CRT_PI * (-aA * l2[1] * mchi[i] + aB * mphi[i]);
        } else {
            k.dnM[i] = CRT_PI * aA * (mphi[i] - lam[i] * mchi[i]);
            k.dnP[i] = CRT_PI * aB * (mphi[i] + lam[i] * mchi[i]);
            k.upM[i] = CRT_PI * aA * (mphi[i] + lam[i] * mchi[i]);
            k.upP[i] = CRT_PI * aB * (mphi[i] - lam[i] * mchi[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double dn_i = CRT_PI * c[i] * (mphi[i] - kap * mchi[i]);  // coefficient of F_i in I_dn (F' = -kappa F + ...)
        const double up_i = CRT_PI * c[i] * (mphi[i] + kap * mchi[i]);
        if (res == i + 1) {
            dK += dn_i;
            uK += up_i;
            if (i == 1 && entire1) {  // R' = -kappa R + B
                k.dnP[i] += CRT_PI * c[i] * mchi[i];
                k.upP[i] -= CRT_PI * c[i] * mchi[i];
            } else {                  // R' = -kappa R - A/(kappa + lam)
                const double ik = 1.0 / (kap + lam[i]);
                k.dnM[i] -= CRT_PI * c[i] * mchi[i] * ik;
                k.upM[i] += CRT_PI * c[i] * mchi[i] * ik;
            }
        } else {
            const double id = 1.0 / (k2 - l2[i]);
            if (res == 0) {
                dK += dn_i * id;
                uK += up_i * id;
            } else {  // the other mode is resonant: e^{-kappa x} = A_k [- kappa B_k] + (kappa^2 - l2_k) R_k
                const int kk = res - 1;
                const double dk = dn_i * id, uk = up_i * id;
                k.dnM[kk] += dk;
                k.upM[kk] += uk;
                if (kk == 1 && entire1) {
                    k.dnP[kk] -= kap * dk;
                    k.upP[kk] -= kap * uk;
                }
                dK += dk * (k2 - l2[kk]);
                uK += uk * (k2 - l2[kk]);
            }
        }
    }
    k.dnK = dK;
    k.upK = uK;
    k.Idr0 = Idr0;
    k.lam[0] = pack_mode_4s(lam[0], (entire1 ? 1 : 0) | (res << 1));
    k.lam[1] = entire1 ? l2[1] : lam[1];
}

// Eigen-system of the 4-stream operator for one band: N (s'' = N s + ...), its eigenvalues l2 (descending) and the
// eigenvectors phi, normalised to unit max-norm.  One function for the ordinary and the rare path: same numbers.
struct Eig4s {
    double N00, N01, N10, N11, l2[2], phi[2][2];
};
// N and its eigenvalues.  The rarity test (ordinary_4s) is evaluated on l2 by TWO kernels -- the sweep kernel, which
// skips a rare column, and fixup_4s_kernel, which redoes it -- and they must agree on every column, so this part is
// written with explicitly rounded operations: no context-dependent FMA contraction, the same bits everywhere.
#if defined(__CUDA_ARCH__)
#define CRT_MUL(a, b) __dmul_rn((a), (b))
#define CRT_ADD(a, b) __dadd_rn((a), (b))
#define CRT_SUB(a, b) __dsub_rn((a), (b))
#else
#define CRT_MUL(a, b) ((a) * (b))
#define CRT_ADD(a, b) ((a) + (b))
#define CRT_SUB(a, b) ((a) - (b))
#endif
CRT_HD void l2_4s(const Scen4s& s, double omega, Eig4s& E, double& dif, double& disc) {
    const double h = CRT_MUL(0.5, omega);
    const double al = CRT_MUL(CRT_MUL(h, CRT_SUB(1.0, s.mu_s)), s.G2);   // alpha_p = alpha_m (ref :190-191), P = 1
    const double be = CRT_MUL(CRT_MUL(h, CRT_SUB(1.0, s.mu_s)), s.G1);   // beta  (ref :192-193)
    const double ga = CRT_MUL(CRT_MUL(h, s.mu_s), s.G1);                 // gamma (ref :194-195)
    // index 0 <-> sector 2 (|mu| in [mu_s, 1]), index 1 <-> sector 1;  q = -(P - Q) diagonal
    const double T00 = CRT_MUL(CRT_SUB(CRT_MUL(2.0, al), s.G2), s.inv_m2), T01 = CRT_MUL(CRT_MUL(2.0, be), s.inv_m2);
    const double T10 = CRT_MUL(CRT_MUL(2.0, be), s.inv_m1), T11 = CRT_MUL(CRT_SUB(CRT_MUL(2.0, ga), s.G1), s.inv_m1);
    E.N00 = -CRT_MUL(s.q0, T00); E.N01 = -CRT_MUL(s.q0, T01); E.N10 = -CRT_MUL(s.q1, T10); E.N11 = -CRT_MUL(s.q1, T11);
    // eigenvalues of N (real: N01 N10 >= 0)
    const double tr = CRT_ADD(E.N00, E.N11);
    dif = CRT_SUB(E.N00, E.N11);
    disc = sqrt(CRT_ADD(CRT_MUL(dif, dif), CRT_MUL(4.0, CRT_MUL(E.N01, E.N10))));
    const double det = CRT_SUB(CRT_MUL(E.N00, E.N11), CRT_MUL(E.N01, E.N10));
    E.l2[0] = CRT_MUL(0.5, CRT_ADD(tr, disc));
    E.l2[1] = det / E.l2[0];
}
CRT_HD void eig_4s(const Scen4s& s, double omega, Eig4s& E) {
    double dif, disc;
    l2_4s(s, omega, E, dif, disc);
    const double N00 = E.N00, N01 = E.N01, N10 = E.N10, N11 = E.N11;
    if (dif >= 0.0) {
        E.phi[0][0] = E.l2[0] - N11; E.phi[0][1] = N10;
        E.phi[1][0] = N01;           E.phi[1][1] = 0.5 * (-dif - disc);   // l2[1] - N00 without cancellation
    } else {
        E.phi[0][0] = N01;           E.phi[0][1] = 0.5 * (-dif + disc);   // l2[0] - N00
        E.phi[1][0] = 0.5 * (dif - disc); E.phi[1][1] = N10;              // l2[1] - N11
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double nrm = rcp_nr(fmax(fabs(E.phi[i][0]), fabs(E.phi[i][1])));
        E.phi[i][0] *= nrm;
        E.phi[i][1] *= nrm;
    }
}
// rare columns: vanishing / negative lambda_1^2, kappa near lambda_k  (also catches NaN inputs)
CRT_HD bool ordinary_4s(const Scen4s& s, const double (&l2)[2]) {
    const double k2g = CRT_MUL(s.kappa, s.kappa);
    const double w = CRT_MUL(CRT_4S_RESONANCE_WIDTH, k2g);
    return CRT_MUL(CRT_MUL(l2[1], s.L_T), s.L_T) >= CRT_4S_ENTIRE_BELOW && fabs(CRT_SUB(k2g, l2[0])) > w && fabs(CRT_SUB(k2g, l2[1])) > w;
}

// The rare columns, OUT OF LINE and self-contained (scalars in, coefficients out through memory; it rebuilds the
// eigen-system itself): nothing of the ordinary path is live across this call or passed to it by reference, so the
// ordinary path's coefficients stay in registers.  (With the general form called on coef_4s's own l2 / phi / k the
// coefficient stage of every column carried a 960-byte stack frame: 4s 0.83 -> 0.56 of HBM peak.)
CRT_HD_NOINLINE void coef_4s_rare(const Scen4s* sp, double r, double t, double rho, double Idr0, double Idf0, Coef4s* out) {
    const Scen4s& s = *sp;
    Eig4s E;
    eig_4s(s, r + t, E);
    const double Nm[2][2] = {{E.N00, E.N01}, {E.N10, E.N11}};
    Coef4s kg;
    coef_4s_general(s, r, t, rho, Idr0, Idf0, Nm, E.l2, E.phi, kg);
    // Resonance only (both modes exponential, nothing can overflow): keep the column on the fast level path --
    // scaled growing modes, exponentials by recurrence -- with X = R_k(x) flagged in lam[0]'s low bits.
    const int mode = unpack_mode_4s(kg.lam[0]);
    const double l0 = sqrt(E.l2[0]);
    if ((mode & 1) == 0 && (mode >> 1) != 0 && l0 * s.L_T < 600.0) {
        kg.lam[0] = tag_res_4s(l0, mode >> 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double gi = exp_neg(kg.lam[i] * s.L_T);
            kg.dnP[i] *= gi;
            kg.upP[i] *= gi;
        }
    }
    *out = kg;
}

// Ordinary columns only.  For a rare column it returns lam[0] = NaN and the CALLER invokes coef_4s_rare -- from a
// place where little is live (the call sitting inside this function cost every column ~200 bytes of spills).
CRT_HD bool coef_4s_is_rare(const Coef4s& k) { return k.lam[0] != k.lam[0]; }
CRT_HD Coef4s coef_4s(const Scen4s& s, double r, double t, double rho, double Idr0, double Idf0) {
    const double omega = r + t;                                   // ref :180
    const double R_dr0 = Idr0 * s.inv_pi_mu0;                     // I / (pi mu): irradiance -> radiance (ref :169)
    const double R_df0 = Idf0 * (1.0 / CRT_PI);                   // ref :170
    const double e1 = 0.25 * omega * R_dr0 * s.mu_s;              // eps_1 (ref :196-197)
    const double e2 = 0.25 * omega * R_dr0 * (1.0 - s.mu_s);      // eps_2 (ref :198-199)
    const double m1 = s.m1, m2 = s.m2;
    const double q0 = s.q0, q1 = s.q1;                            // -(P - Q) diagonal
    const double G = s.kappa * s.mu0;
    const double vD0 = G * e2 * s.inv_m2, vD1 = G * e1 * s.inv_m1;  // forcing of rows 1, 2 (ref :80-87)
    Eig4s E;
    eig_4s(s, omega, E);
    if (!ordinary_4s(s, E.l2)) {  // the caller redoes this column with coef_4s_rare (see coef_4s_is_rare)
        Coef4s kr;
        kr.lam[0] = kr.lam[1] = kr.dnK = kr.upK = nan("");
        kr.Idr0 = Idr0;
#pragma unroll
        for (int i = 0; i < 2; ++i) kr.dnP[i] = kr.dnM[i] = kr.upP[i] = kr.upM[i] = 0.0;
        return kr;
    }
    const double N00 = E.N00, N01 = E.N01, N10 = E.N10, N11 = E.N11;
    const double (&l2)[2] = E.l2;
    const double (&phi)[2][2] = E.phi;
    Coef4s k;
    double psi_[2][2], g[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        k.lam[i] = i == 0 ? tag_res_4s(sqrt(l2[0]), 0) : sqrt(l2[i]);
        psi_[i][0] = -k.lam[i] * phi[i][0] * s.inv_q0;   // (P-Q)^{-1} phi lambda
        psi_[i][1] = -k.lam[i] * phi[i][1] * s.inv_q1;
        g[i] = exp_neg(k.lam[i] * s.L_T);
    }
    // particular solution of the direct problem: (kappa^2 I - N) s_p = 2 (P-Q) vD
    const double kap = s.kappa, k2 = kap * kap;
    const double B00 = k2 - N00, B01 = -N01, B10 = -N10, B11 = k2 - N11;
    const double r0 = -2.0 * q0 * vD0, r1 = -2.0 * q1 * vD1;
    const double idB = rcp_nr(B00 * B11 - B01 * B10);
    const double sp0 = (r0 * B11 - B01 * r1) * idB, sp1 = (B00 * r1 - B10 * r0) * idB;
    const double wp0 = kap * sp0 * s.inv_q0, wp1 = kap * sp1 * s.inv_q1;  // (P-Q)^{-1} (-kappa s_p)
    const double Dp0 = 0.5 * (sp0 + wp0), Dp1 = 0.5 * (sp1 + wp1);
    const double Up0 = 0.5 * (sp0 - wp0), Up1 = 0.5 * (sp1 - wp1);

    // boundary conditions -> A c = rhs, c = [a~_1, a~_2, b_1, b_2]
    double A[4][4], rhs[4];
    double mP[2], mM[2];  // (m2, m1) . (phi +- psi)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double p0 = phi[i][0] + psi_[i][0], p1 = phi[i][1] + psi_[i][1];
        const double n0 = phi[i][0] - psi_[i][0], n1 = phi[i][1] - psi_[i][1];
        mP[i] = m2 * p0 + m1 * p1;
        mM[i] = m2 * n0 + m1 * n1;
        // top, x = 0:  D = incident            (ref :113-119, :137)
        A[0][i] = 0.5 * p0 * g[i];   A[0][2 + i] = 0.5 * n0;
        A[1][i] = 0.5 * p1 * g[i];   A[1][2 + i] = 0.5 * n1;
        // bottom, x = LAI:  U_i - 2 rho (m2 D_0 + m1 D_1) = rho mu0 R_dr0 e^{-kappa LAI}   (ref :129-137)
        A[2][i] = 0.5 * n0 - rho * mP[i];   A[2][2 + i] = (0.5 * p0 - rho * mM[i]) * g[i];
        A[3][i] = 0.5 * n1 - rho * mP[i];   A[3][2 + i] = (0.5 * p1 - rho * mM[i]) * g[i];
    }
    const double mDp = m2 * Dp0 + m1 * Dp1;
    const double bot = s.eKT * (rho * s.mu0 * R_dr0 + 2.0 * rho * mDp);
    rhs[0] = R_df0 - Dp0;
    rhs[1] = R_df0 - Dp1;
    rhs[2] = bot - s.eKT * Up0;
    rhs[3] = bot - s.eKT * Up1;
    solve4(A, rhs);
#pragma unroll
    for (int i = 0; i < 2; ++i) {  // irradiances: 2 pi (mu_1 R1 + mu_2 R2)   (ref :246-249, :277-281)
        k.dnP[i] = CRT_PI * rhs[i] * mP[i];
        k.dnM[i] = CRT_PI * rhs[2 + i] * mM[i];
        k.upP[i] = CRT_PI * rhs[i] * mM[i];
        k.upM[i] = CRT_PI * rhs[2 + i] * mP[i];
    }
    k.dnK = 2.0 * CRT_PI * mDp;
    k.upK = 2.0 * CRT_PI * (m2 * Up0 + m1 * Up1);
    k.Idr0 = Idr0;
    if (k.lam[0] * s.L_T < 600.0) {  // lam[0] is the larger eigenvalue
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            k.dnP[i] *= g[i];
            k.upP[i] *= g[i];
        }
    } else {
        k.lam[0] = pack_mode_4s(k.lam[0], 0);
    }
    return k;
}

// One level of one 4s column from its basis functions: m_k = e^{-lam_k x}, p_k = (scaled) e^{-lam_k (LAI - x)} (or the
// entire pair), X = e^{-kappa x} (or the resonance-safe R_k); eK = e^{-kappa x} for the direct beam.
CRT_HD void level_4s_x(const Scen4s& s, const Coef4s& k, double eK, double X, double m0, double p0, double m1, double p1,
                       double& Idr, double& dn, double& up, double& F) {
    dn = k.dnK * X + (k.dnP[0] * p0 + k.dnM[0] * m0) + (k.dnP[1] * p1 + k.dnM[1] * m1);
    up = k.upK * X + (k.upP[0] * p0 + k.upM[0] * m0) + (k.upP[1] * p1 + k.upM[1] * m1);
    Idr = k.Idr0 * eK;                        // ref :284
    F = Idr * s.inv_mu + 2.0 * up + 2.0 * dn; // ref :290
}
CRT_HD void level_4s_e(const Scen4s& s, const Coef4s& k, double eK, double m0, double p0, double m1, double p1,
                       double& Idr, double& dn, double& up, double& F) {
    level_4s_x(s, k, eK, eK, m0, p0, m1, p1, Idr, dn, up, F);
}

// Slow form: every basis function of level x evaluated directly (mode word: see Coef4s).  Out of line, scalars by
// value: the scenario / coefficient structs of the caller must not have their address taken (they would live in
// local memory for the whole kernel).
CRT_HD_NOINLINE void basis_4s_slow(double L_T, double kappa, double lam0_stored, double lam1, double x, double eK, double* b5) {
    const int mode = unpack_mode_4s(lam0_stored);
    const double l0 = -lam0_stored, xr = L_T - x;
    double X = eK, m1, p1;
    const double m0 = exp(-l0 * x), p0 = exp(-l0 * xr);
    if (mode & 1) {
        csh_entire(lam1, x, m1, p1);
    } else {
        m1 = exp(-lam1 * x);
        p1 = exp(-lam1 * xr);
    }
    const int res = mode >> 1;
    if (res == 1) X = res_exp_4s(kappa, l0, x, eK);
    else if (res == 2) X = (mode & 1) ? res_series_4s(kappa, lam1, x) : res_exp_4s(kappa, lam1, x, eK);
    b5[0] = X; b5[1] = m0; b5[2] = p0; b5[3] = m1; b5[4] = p1;
}

// One level of one 4s column: x = cumulative LAI of the level, eK = exp(-kappa x)   (ref :246-290)
CRT_HD void level_4s(const Scen4s& s, const Coef4s& k, double x, double eK, double& Idr, double& dn, double& up,
                     double& F) {
    double m0, p0, m1, p1;
    if (k.lam[0] > 0.0) {
        exp_pm(k.lam[0] * x, m0, p0);
        exp_pm(k.lam[1] * x, m1, p1);
        const int res = fast_res_4s(k.lam[0]);
        const double X = res ? res_exp_4s(s.kappa, k.lam[res - 1], x, eK) : eK;
        level_4s_x(s, k, eK, X, m0, p0, m1, p1, Idr, dn, up, F);
    } else {
        double b5[5];
        basis_4s_slow(s.L_T, s.kappa, k.lam[0], k.lam[1], x, eK, b5);
        level_4s_x(s, k, eK, b5[0], b5[1], b5[2], b5[3], b5[4], Idr, dn, up, F);
    }
}

// One level of an ORDINARY 4s column (real, well separated eigenvalues; not resonant): what the row-sweep kernel
// evaluates -- its rare columns are redone by fixup_4s_kernel, so none of the general machinery is compiled into it.
CRT_HD void level_4s_plain(const Scen4s& s, const Coef4s& k, double x, double eK, double& Idr, double& dn, double& up,
                           double& F) {
    double m0, p0, m1, p1;
    if (k.lam[0] > 0.0) {
        exp_pm(k.lam[0] * x, m0, p0);
        exp_pm(k.lam[1] * x, m1, p1);
    } else {  // lam_0 LAI >= 600: unscaled coefficients, direct exponentials
        const double l0 = -k.lam[0], xr = s.L_T - x;
        m0 = exp(-l0 * x);
        p0 = exp(-l0 * xr);
        m1 = exp(-k.lam[1] * x);
        p1 = exp(-k.lam[1] * xr);
    }
    level_4s_e(s, k, eK, m0, p0, m1, p1, Idr, dn, up, F);
}

// L[j], eK[j] = exp(-kappa L[j]) level tables.
// 4s is FP64-bound (two exp_pm per level), so on equally spaced levels -- every profile the reference's
// generators make (ref ../leaf_area.py:82-88) -- the four exponentials are advanced by constant factors
// e^{+-lam_k dL} and re-anchored with exact exp_pm every LV4 levels (drift <= LV4 ulp).  Spacing is checked
// per group against the level table; irregular groups and flagged (huge lam LAI) columns take exp_pm.
constexpr int LV4 = 8;
template <int VEC, class Out>
CRT_HD void column_4s(const Scen4s& s, const double* L, const double* eK, int n_z, const BandIn<VEC>& in, Out& out,
                      double (&absorbed)[VEC]) {
    Coef4s k[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) k[v] = coef_4s(s, in.leaf_r[v], in.leaf_t[v], in.soil_r[v], in.Idr0[v], in.Idf0[v]);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (coef_4s_is_rare(k[v])) {
            Coef4s kg;  // through copies: neither k[] nor s may have their address taken on the common path
            const Scen4s sg = s;
            coef_4s_rare(&sg, in.leaf_r[v], in.leaf_t[v], in.soil_r[v], in.Idr0[v], in.Idf0[v], &kg);
            k[v] = kg;
        }
    }
    double gnd[VEC][3];
    const double tol = 8.0 * 2.220446049250313e-16 * s.L_T;
    for (int j0 = 0; j0 < n_z; j0 += LV4) {
        const int j1 = (j0 + LV4 < n_z) ? j0 + LV4 : n_z;
        bool uniform = (j1 - j0) >= 3;
        const double dL = (j1 - j0) >= 2 ? L[j0] - L[j0 + 1] : 0.0;
        for (int j = j0 + 1; j + 1 < j1; ++j) uniform = uniform && fabs((L[j] - L[j + 1]) - dL) <= tol;
        double m0[VEC], p0[VEC], m1[VEC], p1[VEC], qi0[VEC], qd0[VEC], qi1[VEC], qd1[VEC];
        bool rec[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            rec[v] = uniform && k[v].lam[0] > 0.0;
            if (rec[v]) {
                exp_pm(k[v].lam[0] * L[j0], m0[v], p0[v]);
                exp_pm(k[v].lam[1] * L[j0], m1[v], p1[v]);
                exp_pm(k[v].lam[0] * dL, qd0[v], qi0[v]);  // x shrinks by dL per level: m grows, p decays
                exp_pm(k[v].lam[1] * dL, qd1[v], qi1[v]);
            }
        }
        for (int j = j0; j < j1; ++j) {
            double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (rec[v]) {
                    const int res = fast_res_4s(k[v].lam[0]);
                    const double X = res ? res_exp_4s(s.kappa, k[v].lam[res - 1], L[j], eK[j]) : eK[j];
                    level_4s_x(s, k[v], eK[j], X, m0[v], p0[v], m1[v], p1[v], Idr[v], dn[v], up[v], F[v]);
                    m0[v] *= qi0[v];
                    p0[v] *= qd0[v];
                    m1[v] *= qi1[v];
                    p1[v] *= qd1[v];
                } else {
                    level_4s(s, k[v], L[j], eK[j], Idr[v], dn[v], up[v], F[v]);
                }
                if (j == 0) { gnd[v][0] = Idr[v]; gnd[v][1] = dn[v]; gnd[v][2] = up[v]; }
                if (j == n_z - 1) absorbed[v] = absorbed_from_ends(Idr[v], gnd[v][0], dn[v], gnd[v][1], up[v], gnd[v][2]);
            }
            out.st(F_IDR, j, Idr);
            out.st(F_DN, j, dn);
            out.st(F_UP, j, up);
            out.st(F_F, j, F);
        }
    }
}

}  // namespace crt
