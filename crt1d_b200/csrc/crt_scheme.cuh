// crt_scheme.cuh -- scheme dispatch shared by the CUDA kernels and the host-compiled test harness:
// which band-independent level tables a scheme needs (the scenario prologue the reference computes
// before its band loop) and how one group of VEC columns is solved from them.
#pragma once

#include "../../include/crt1d_b200.h"
#include "crt_core.cuh"

namespace crt {

// Number of n_z-long level tables each scheme keeps (shared memory on the device).
CRT_HD int n_level_tables(int scheme) {
    switch (scheme) {
        case CRT1D_SCHEME_2S: return 2;   // L, exp(-K L)
        case CRT1D_SCHEME_4S: return 2;   // L, exp(-K L)
        case CRT1D_SCHEME_BL: return 3;   // L, tau_b, tau_d
        case CRT1D_SCHEME_BF: return 2;   // L, exp(-k_b L)
        case CRT1D_SCHEME_G77: return 2;  // L, exp(-k_b L)
        case CRT1D_SCHEME_N79: return 8;  // tbcum, g, td, 1-td, fsun, 1-fsun, 1/(fsun dlai), 1/((1-fsun) dlai)
        case CRT1D_SCHEME_ZQ: return 1;   // exp(-K L)
        case CRT1D_SCHEME_ZQ_PA: return 12;  // lai, exp(-Kb lai), cum[0..M], exp(-Kb cum[0..M]) (M <= n_z; 3 slots), kk, tt, ww; in emission order: (level, kk), tt, ww, exp(-Kb lai)
        default: return 0;
    }
}

// Level-table entry j of scenario s (what each reference solver evaluates once, outside its band loop).
// `tab` holds n_level_tables(scheme) consecutive arrays of n_z doubles.
template <int SCHEME>
CRT_HD void fill_level_tables(const crt1d_batch& in, int64_t s, int j, double* tab) {
    const int n_z = in.n_z;
    const double K_b = in.K_b[s];
    const double* lai = in.lai_lib + (int64_t)in.lai_idx[s] * n_z;
    const double Lj = lai[j];
    if constexpr (SCHEME == CRT1D_SCHEME_2S || SCHEME == CRT1D_SCHEME_4S || SCHEME == CRT1D_SCHEME_BF ||
                  SCHEME == CRT1D_SCHEME_G77) {
        tab[j] = Lj;
        tab[n_z + j] = exp(-K_b * Lj);  // _solve_2s.py:125,150; _solve_4s.py:284; _solve_bf.py:90-93; _solve_g77.py:77-80
    } else if constexpr (SCHEME == CRT1D_SCHEME_BL) {
        tab[j] = Lj;
        tab[n_z + j] = exp(-K_b * Lj);  // tau_b, _solve_bl.py:31
        tab[2 * n_z + j] = in.tau_d_lev[(int64_t)in.lai_idx[s] * n_z + j];  // tau_d, _solve_bl.py:35-37 (prologue quadrature)
    } else if constexpr (SCHEME == CRT1D_SCHEME_ZQ) {
        tab[j] = exp(-K_b * Lj);  // S / I_dr0, _solve_zq.py:131
    } else if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        tab[j] = Lj;
        tab[n_z + j] = exp(-K_b * Lj);  // f_slo, _solve_zq_pa.py:352
        if (j == 0) {  // running sum of LAI/M exactly as np.cumsum accumulates it (_solve_zq_pa.py:161)
            const int M = n_z < ZQPA_MAX_M ? n_z : ZQPA_MAX_M;
            const double dl = lai[0] / M;
            double* cum = tab + 2 * n_z;
            double c = 0.0;
            cum[0] = 0.0;
            for (int i = 1; i <= M; ++i) {  // serial by definition; the exponentials of cum[] are taken in the second pass,
                c += dl;                    // one per thread (100 dependent exp calls here were ~30 000 cycles of a
                cum[i] = c;                 // one-thread prologue in every CTA)
            }
        }
    } else if constexpr (SCHEME == CRT1D_SCHEME_N79) {
        // Everything of a row that does not depend on the band (ref _solve_n79.py:41-59, :85-129, :145-155):
        //   [0] tbcum[j] = exp(-K_b L[j])
        //   [1] g[j]     = tbcum[j] (1 - tb[j-1]): the beam source of the upward row of level j and of the downward
        //                  row of level j-1 (j >= 1); g[0] = tbcum[1] (1 - tb[1]), the soil-adjacent row as shipped (:85-92)
        //   [2] td[j], [3] 1 - td[j]   layer j (between levels j and j+1): tau_d(dlai[j]) from the prologue
        //   [4] fsun[j], [5] 1 - fsun[j], [6] 1 / (fsun dlai), [7] 1 / ((1 - fsun) dlai)
        auto tb_of = [&](int i) { return exp(-K_b * (lai[i] - lai[i + 1])); };  // tb[i], :45
        tab[j] = exp(-K_b * Lj);
        if (j == 0) {
            tab[n_z] = exp(-K_b * lai[1]) * (1.0 - tb_of(1));
        } else {
            tab[n_z + j] = tab[j] * (1.0 - tb_of(j - 1));
        }
        if (j < n_z - 1) {
            const double dl = Lj - lai[j + 1];                                    // :41
            // :53 (prologue quadrature / 9sky).  A zero-thickness layer has tau_d = 1 and is the identity; the layer
            // algebra (:85-88, divisions by (1 - td) rho) reaches that limit smoothly -- any td in [1 - 1e-15, 1)
            // gives the same fluxes to 1e-14 -- but not AT td == 1 (0 * inf).  The reference's own quad returns
            // 1 - 2^-53 for L = 0, an exact prologue (the device Gauss-Legendre rule) returns 1: use the former.
            // (Only the exact value is replaced: the "9sky" rule legitimately returns td > 1 for thin layers.)
            const double td_in = in.tau_d_lev[(int64_t)in.lai_idx[s] * n_z + j];
            const double tdj = td_in == 1.0 ? 1.0 - 0x1p-53 : td_in;
            const double fs = exp(-K_b * ((Lj + lai[j + 1]) / 2.0));              // fracsun, :57-58
            tab[2 * n_z + j] = tdj;
            tab[3 * n_z + j] = 1.0 - tdj;
            tab[4 * n_z + j] = fs;
            tab[5 * n_z + j] = 1.0 - fs;
            tab[6 * n_z + j] = 1.0 / (fs * dl);
            tab[7 * n_z + j] = 1.0 / ((1.0 - fs) * dl);
        } else {
            for (int q = 2; q < 8; ++q) tab[q * n_z + j] = 0.0;
        }
    }
}

// Second and third passes over the levels, each after a barrier (zq_pa only): the interpolation tables of
// the caller's levels on the M-grid (needs cum[] from the first pass) and their processing order.
template <int SCHEME>
CRT_HD void fill_level_tables_2(const crt1d_batch& in, int64_t s, int j, double* tab) {
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        const int n_z = in.n_z;
        const int M = n_z < ZQPA_MAX_M ? n_z : ZQPA_MAX_M;
        const double* cum = tab + 2 * n_z;
        {  // exp(-Kb cum[i]), i = 0..M (M <= n_z: the thread of the last level takes the extra entry when M = n_z)
            const double K_b = in.K_b[s];
            double* eC = tab + 2 * n_z + (M + 1);
            if (j <= M) eC[j] = j == 0 ? 1.0 : exp(-K_b * cum[j]);
            if (j == n_z - 1 && M == n_z) eC[M] = exp(-K_b * cum[M]);
        }
        int k;
        double t, w;
        interp_np_prepare(tab[j], cum, M + 1, tab[0] / M, k, t, w);
        tab[5 * n_z + j] = (double)k;
        tab[6 * n_z + j] = t;
        tab[7 * n_z + j] = w;
    }
}
template <int SCHEME>
CRT_HD void fill_level_tables_3(const crt1d_batch& in, int64_t s, int j, double* tab) {
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        const int n_z = in.n_z;
        const double* kk = tab + 5 * n_z;
        int rank = 0;  // levels finished before level j: larger kk first, ties from the top level down
        for (int i = 0; i < n_z; ++i) rank += (kk[i] > kk[j]) || (kk[i] == kk[j] && i > j);
        // emission-ordered copies, so the back sweep walks all four tables with one cursor:
        // (level, kk) packed as two int32 in one slot; t, w and exp(-Kb lai) of that level
        int* lk = reinterpret_cast<int*>(tab + 8 * n_z + rank);
        lk[0] = j;
        lk[1] = (int)kk[j];
        tab[9 * n_z + rank] = tab[6 * n_z + j];
        tab[10 * n_z + rank] = tab[7 * n_z + j];
        tab[11 * n_z + rank] = tab[n_z + j];
    }
}
// Solve VEC adjacent columns of scenario s.  `tab` = the level tables above.
template <int SCHEME, int VEC, class Out>
CRT_HD void solve_column_group(const crt1d_batch& in, int64_t s, const double* tab, const BandIn<VEC>& b, Out& out,
                               double (&rho_c)[VEC], double (&absorbed)[VEC]) {
    const int n_z = in.n_z;
    const double psi = in.psi[s], K_b = in.K_b[s];
    const double L_T = in.lai_lib[(int64_t)in.lai_idx[s] * n_z];
#pragma unroll
    for (int v = 0; v < VEC; ++v) rho_c[v] = 0.0;
    if constexpr (SCHEME == CRT1D_SCHEME_2S) {
        const Scen2s sc = scen_2s(psi, K_b, in.mu_bar[s], in.mla_deg, L_T);
        column_2s<VEC>(sc, tab, tab + n_z, n_z, b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_4S) {
        const double mu_s = in.mu_s > 0.0 ? in.mu_s : 0.501;
        const Scen4s sc = scen_4s(psi, K_b, in.G_int[2 * s], in.G_int[2 * s + 1], mu_s, L_T);
        column_4s<VEC>(sc, tab, tab + n_z, n_z, b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_BL) {
        ScenBl sc;
        sc.K_b = K_b;
        sc.inv_mu = 1.0 / cos(psi);
        column_bl<VEC>(sc, tab, tab + n_z, tab + 2 * n_z, n_z, b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_BF) {
        const ScenBf sc = scen_bf(psi, K_b, L_T);
        column_bf<VEC>(sc, tab, tab + n_z, n_z, b, out, rho_c, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_G77) {
        const ScenBf sc = scen_bf(psi, K_b, L_T);
        column_g77<VEC>(sc, tab, tab + n_z, n_z, b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_N79) {
        ScenN79 sc;
        sc.inv_mu = 1.0 / cos(psi);
        column_n79<VEC>(sc, tab, n_z, b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        ScenZqPa sc;
        sc.cos_psi = cos(psi);
        sc.inv_mu = 1.0 / sc.cos_psi;
        sc.Kb = K_b;
        sc.tau_d = in.tau_i[s];
        sc.LAI = L_T;
        sc.M = n_z < ZQPA_MAX_M ? n_z : ZQPA_MAX_M;
        column_zq_pa<VEC>(sc, tab + 2 * n_z + sc.M + 1, tab + 8 * n_z, tab + 9 * n_z, tab + 10 * n_z, tab + 11 * n_z, n_z,
                          b, out, absorbed);
    } else if constexpr (SCHEME == CRT1D_SCHEME_ZQ) {
        ScenZq sc;
        sc.cos_psi = cos(psi);
        sc.inv_mu = 1.0 / sc.cos_psi;
        sc.tau_i = in.tau_i[s];
        sc.t_psi = in.tau_psi[s];
        column_zq<VEC>(sc, tab, n_z, b, out, absorbed);
    }
}

// Load the band inputs of VEC adjacent bands starting at b0 (b0 + VEC <= n_wl).
template <int VEC>
CRT_HD BandIn<VEC> load_bands(const crt1d_batch& in, int64_t s, int b0) {
    BandIn<VEC> b;
    const int64_t n_wl = in.n_wl;
    const double* lr = in.leaf_r_lib + (int64_t)in.leaf_idx[s] * n_wl + b0;
    const double* lt = in.leaf_t_lib + (int64_t)in.leaf_idx[s] * n_wl + b0;
    const double* sr = in.soil_r_lib ? in.soil_r_lib + (int64_t)in.soil_idx[s] * n_wl + b0 : nullptr;
    const double* dr = in.I_dr0_lib + (int64_t)in.sky_idx[s] * n_wl + b0;
    const double* df = in.I_df0_lib + (int64_t)in.sky_idx[s] * n_wl + b0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        b.leaf_r[v] = lr[v];
        b.leaf_t[v] = lt[v];
        b.soil_r[v] = sr ? sr[v] : 0.0;
        b.Idr0[v] = dr[v];
        b.Idf0[v] = df[v];
    }
    return b;
}

// Rows per scenario of extra-output slot x (n79's absorbed profiles live on the n_z - 1 layers).
CRT_HD int extra_rows(int scheme, int n_z) { return scheme == CRT1D_SCHEME_N79 ? n_z - 1 : n_z; }

}  // namespace crt
