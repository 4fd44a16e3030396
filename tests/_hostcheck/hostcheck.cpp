// hostcheck.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by crt1d_b200/, bench.py or smoke()).
// Compiles the per-column arithmetic that the CUDA kernels execute (crt1d_b200/csrc/crt_core.cuh,
// crt_scheme.cuh, crt_leafangle.cuh are __host__ __device__) for the host, so the `-m "not gpu"` test
// tier can check the kernel math against the oracle in a container without a GPU.  It is not a
// fallback: the product has no code path that reaches this file.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../crt1d_b200/csrc/crt_leafangle.cuh"
#include "../../crt1d_b200/csrc/crt_scheme.cuh"
#include "../../crt1d_b200/csrc/crt_spectra.cuh"

namespace {

template <int VEC>
struct HostOut {
    double* p[crt::N_FIELDS];
    int64_t stride;
    void st(int f, int j, const double (&x)[VEC]) const {
        if (!p[f]) return;
        for (int v = 0; v < VEC; ++v) p[f][(int64_t)j * stride + v] = x[v];
    }
    void st4(int j, const double (&a)[VEC], const double (&b)[VEC], const double (&c)[VEC], const double (&d)[VEC]) const {
        st(crt::F_IDR, j, a);
        st(crt::F_DN, j, b);
        st(crt::F_UP, j, c);
        st(crt::F_F, j, d);
    }
    void st_tmp(int f, int j, const double (&x)[VEC]) const { st(f, j, x); }
    // segment store of the checkpointed Thomas sweeps (shared memory on the device)
    static constexpr int CK = 8;
    mutable double seg_buf[CK + 100 / CK + 4][2][VEC];  // + zq_pa's checkpoint slots
    int seg_levels() const { return CK; }
    void seg_st(int slot, int k, const double (&x)[VEC]) const {
        for (int v = 0; v < VEC; ++v) seg_buf[slot][k][v] = x[v];
    }
    void seg_ld(int slot, int k, double (&x)[VEC]) const {
        for (int v = 0; v < VEC; ++v) x[v] = seg_buf[slot][k][v];
    }
    void ld_tmp(int f, int j, double (&x)[VEC]) const {
        for (int v = 0; v < VEC; ++v) x[v] = p[f][(int64_t)j * stride + v];
    }
    void pf_tmp(int, int, int) const {}  // checkpoint prefetch: a cp.async on the device, nothing to do here
    void ld_pf(int f, int j, int, double (&x)[VEC]) const { ld_tmp(f, j, x); }
};

template <int SCHEME, int VEC>
void run(const crt1d_batch& in, const crt1d_out& out) {
    const int n_z = in.n_z, n_wl = in.n_wl;
    std::vector<double> tab((size_t)crt::n_level_tables(SCHEME) * n_z);
    const int64_t prof = (int64_t)n_z * n_wl, xprof = (int64_t)crt::extra_rows(SCHEME, n_z) * n_wl;
    for (int64_t s = 0; s < in.n_scen; ++s) {
        for (int j = 0; j < n_z; ++j) crt::fill_level_tables<SCHEME>(in, s, j, tab.data());
        for (int j = 0; j < n_z; ++j) crt::fill_level_tables_2<SCHEME>(in, s, j, tab.data());
        for (int j = 0; j < n_z; ++j) crt::fill_level_tables_3<SCHEME>(in, s, j, tab.data());
        double acc[4] = {0, 0, 0, 0};
        for (int b0 = 0; b0 < n_wl; b0 += VEC) {
            const crt::BandIn<VEC> b = crt::load_bands<VEC>(in, s, b0);
            HostOut<VEC> o;
            o.stride = n_wl;
            o.p[crt::F_IDR] = out.I_dr ? out.I_dr + s * prof + b0 : nullptr;
            o.p[crt::F_DN] = out.I_df_d ? out.I_df_d + s * prof + b0 : nullptr;
            o.p[crt::F_UP] = out.I_df_u ? out.I_df_u + s * prof + b0 : nullptr;
            o.p[crt::F_F] = out.F ? out.F + s * prof + b0 : nullptr;
            o.p[crt::F_X0] = out.x0 ? out.x0 + s * xprof + b0 : nullptr;
            o.p[crt::F_X1] = out.x1 ? out.x1 + s * xprof + b0 : nullptr;
            o.p[crt::F_X2] = out.x2 ? out.x2 + s * xprof + b0 : nullptr;
            double rho_c[VEC], ab[VEC];
            crt::solve_column_group<SCHEME, VEC>(in, s, tab.data(), b, o, rho_c, ab);
            if (SCHEME == CRT1D_SCHEME_BF && out.rho_c)
                for (int v = 0; v < VEC; ++v) out.rho_c[s * n_wl + b0 + v] = rho_c[v];
            if (out.absorbed)
                for (int k = 0; k < out.n_bw; ++k)
                    for (int v = 0; v < VEC; ++v) acc[k] += out.band_w[(int64_t)k * n_wl + b0 + v] * ab[v];
        }
        if (out.absorbed)
            for (int k = 0; k < out.n_bw; ++k) out.absorbed[s * out.n_bw + k] = acc[k];
    }
}

template <int SCHEME>
void run_vec(const crt1d_batch& in, const crt1d_out& out, int vec) {
    if (vec == 2) run<SCHEME, 2>(in, out);
    else run<SCHEME, 1>(in, out);
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int hostcheck_solve(int scheme, const crt1d_batch* in, const crt1d_out* out, int vec) {
    if (vec == 2 && in->n_wl % 2 != 0) return -1;
    switch (scheme) {
        case CRT1D_SCHEME_2S: run_vec<CRT1D_SCHEME_2S>(*in, *out, vec); break;
        case CRT1D_SCHEME_4S: run_vec<CRT1D_SCHEME_4S>(*in, *out, vec); break;
        case CRT1D_SCHEME_BF: run_vec<CRT1D_SCHEME_BF>(*in, *out, vec); break;
        case CRT1D_SCHEME_BL: run_vec<CRT1D_SCHEME_BL>(*in, *out, vec); break;
        case CRT1D_SCHEME_G77: run_vec<CRT1D_SCHEME_G77>(*in, *out, vec); break;
        case CRT1D_SCHEME_N79: run_vec<CRT1D_SCHEME_N79>(*in, *out, vec); break;
        case CRT1D_SCHEME_ZQ: run_vec<CRT1D_SCHEME_ZQ>(*in, *out, vec); break;
        case CRT1D_SCHEME_ZQ_PA: run_vec<CRT1D_SCHEME_ZQ_PA>(*in, *out, vec); break;
        default: return -1;
    }
    return 0;
}

__attribute__((visibility("default"))) double hostcheck_leaf_G(int family, double param, double psi) {
    return crt::leaf_G(family, param, psi);
}

static crt::QuadRule make_rule(int n_quad, int panels) {
    crt::QuadRule r;
    memset(&r, 0, sizeof(r));
    r.n = n_quad;
    r.panels = panels;
    if (n_quad > 0) crt::gauss_legendre(n_quad, r.x, r.w);
    return r;
}

__attribute__((visibility("default"))) double hostcheck_tau_d(int family, double param, int n_quad, int panels, double L) {
    const crt::QuadRule r = make_rule(n_quad, panels);
    return crt::tau_d_quadrature(family, param, r, L);
}

__attribute__((visibility("default"))) double hostcheck_leaf_integral(int family, double param, double mu_s, int n_quad, int which) {
    const crt::QuadRule r = make_rule(n_quad, 1);
    return crt::leaf_integral(family, param, mu_s, r, which);
}
}

extern "C" __attribute__((visibility("default"))) void hostcheck_exp_pm(int n, const double* x, double* em, double* ep) {
    for (int i = 0; i < n; ++i) crt::exp_pm(x[i], em[i], ep[i]);
}

extern "C" __attribute__((visibility("default"))) void hostcheck_smear_tuv(int n_rows, int n_x, const double* x, const double* y,
                                                                           int n_bins, const double* bins, double* out) {
    for (int r = 0; r < n_rows; ++r)
        for (int i = 0; i < n_bins; ++i) out[(int64_t)r * n_bins + i] = crt::smear_tuv_bin(x, y + (int64_t)r * n_x, n_x, bins[i], bins[i + 1]);
}
