/* crt1d_b200 -- C ABI of the B200-native canopy radiative-transfer solvers.
 *
 * Drop-in boundary for the solver hot path of zmoon/crt1d.  The reference has no FFI (it is pure
 * Python); the interface each entry point replaces is the reference's solver-plugin call
 *     sol = scheme["solver"](**{k: p[k] for k in scheme["args"]}, **extra)       (crt1d/model.py:305-310)
 * i.e. one `solve_<id>(*, psi, I_dr0_all, I_df0_all, lai, leaf_t, leaf_r, soil_r, K_b_fn, G_fn, ...)`
 * per scheme (signatures: crt1d/solvers/_solve_<id>.py, listed per function below), generalised to a
 * batch of S independent scenarios.  Python callables (`K_b_fn`, `G_fn`) cannot cross a C ABI: the caller
 * passes what the reference computes FROM them in each solver's band-independent prologue (K_b, G, the
 * quadratures), or asks the library to evaluate a closed parametric leaf-angle family on the device.
 *
 * Conventions
 *   - All arrays are IEEE float64, C order.  Profiles are [S][n_z][n_wl], band fastest, exactly the
 *     reference's (n_z, n_wl) layout (crt1d/variables.yml:175-208) with a leading scenario axis.
 *   - Level index 0 = ground (lai = total LAI), n_z-1 = canopy top (lai = 0)  (crt1d/model.py:240-246).
 *   - `crt1d_solve*` take DEVICE pointers and enqueue work on `stream` (a cudaStream_t, may be NULL for
 *     the legacy default stream); they never allocate, free or synchronise.  `crt1d_solve_host` takes HOST
 *     pointers, does H2D + solve + D2H itself and returns when the results are in the caller's buffers.
 *   - Return value: CRT1D_OK (0) or a negative error code (`crt1d_solve_host` may also return the positive
 *     finding CRT1D_NONFINITE); `crt1d_last_error()` gives the message of the last failure on the calling
 *     thread.  No entry point aborts, throws, or falls back to the CPU.
 *   - Re-entrant; no global mutable state except a per-thread error string and (host path only) a
 *     per-thread device workspace that is grown on demand and released by `crt1d_release_workspace`.
 */
#ifndef CRT1D_B200_H
#define CRT1D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRT1D_ABI_VERSION 4

/* exported-symbol marker (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define CRT1D_API __attribute__((visibility("default")))
#else
#define CRT1D_API
#endif

/* error codes */
#define CRT1D_OK 0
#define CRT1D_ERR_INVALID_ARG (-1)   /* bad scheme id / family id / size */
#define CRT1D_ERR_NULL_POINTER (-2)  /* a required pointer is NULL */
#define CRT1D_ERR_UNSUPPORTED (-3)   /* valid request this build cannot serve (e.g. n_z too large for shared memory) */
#define CRT1D_ERR_CUDA (-4)          /* a CUDA runtime call failed; see crt1d_last_error() */
#define CRT1D_ERR_NO_DEVICE (-5)     /* no CUDA device visible */
#define CRT1D_ERR_NO_MEMORY (-6)     /* device workspace allocation failed (host path) */
/* positive = results delivered, with a finding (host path only; the asynchronous device path reports it in
 * crt1d_out.status): at least one scenario holds a non-finite value (NaN / Inf inputs, r + t = 0 in 2s, ...).
 * The reference returns such arrays silently (numpy RuntimeWarning at most). */
#define CRT1D_NONFINITE 1

/* crt1d_out.status bits, per scenario */
#define CRT1D_STATUS_NONFINITE 1 /* a non-finite value at the ground or top level of some band column */

/* scheme ids -- the reference's scheme names (crt1d/solvers/__init__.py:17-30) */
#define CRT1D_SCHEME_2S 0   /* Dickinson-Sellers two-stream         crt1d/solvers/_solve_2s.py:11-163  */
#define CRT1D_SCHEME_4S 1   /* Tian et al. four-stream              crt1d/solvers/_solve_4s.py:8-293   */
#define CRT1D_SCHEME_BF 2   /* Bodin & Franklin                     crt1d/solvers/_solve_bf.py:7-154   */
#define CRT1D_SCHEME_BL 3   /* Beer-Lambert                         crt1d/solvers/_solve_bl.py:9-93    */
#define CRT1D_SCHEME_G77 4  /* Goudriaan (1977)                     crt1d/solvers/_solve_g77.py:7-135  */
#define CRT1D_SCHEME_N79 5  /* Norman (1979)                        crt1d/solvers/_solve_n79.py:11-200 */
#define CRT1D_SCHEME_ZQ 6   /* Zhao & Qualls                        crt1d/solvers/_solve_zq.py:13-229  */
#define CRT1D_SCHEME_ZQ_PA 7 /* Zhao & Qualls, pyAPES variant       crt1d/solvers/_solve_zq_pa.py:24-418 */
#define CRT1D_N_SCHEMES 8

/* leaf-angle families (crt1d/leaf_angle.py:118-202) */
#define CRT1D_G_SPHERICAL 0
#define CRT1D_G_HORIZONTAL 1
#define CRT1D_G_VERTICAL 2
#define CRT1D_G_ELLIPSOIDAL_APPROX 3        /* param = x  (default case: crt1d/cases.py:29-30) */
#define CRT1D_G_ELLIPSOIDAL 4               /* param = x  */
#define CRT1D_G_ELLIPSOIDAL_APPROX_BONAN 5  /* param = chi_l */

/* One batch of S scenarios over shared libraries of profiles and spectra.
 * A scenario is one call of a reference solver.  Unused fields for a scheme may be NULL / 0. */
typedef struct crt1d_batch {
    int64_t n_scen; /* S */
    int32_t n_z;    /* interface levels per profile ("nlayers" in crt1d.Model) */
    int32_t n_wl;   /* wavelength bands */
    int32_t n_lai;  /* rows in lai_lib / tau_d_lev */
    int32_t n_leaf; /* rows in leaf_r_lib, leaf_t_lib */
    int32_t n_soil; /* rows in soil_r_lib */
    int32_t n_sky;  /* rows in I_dr0_lib, I_df0_lib */

    /* per scenario, [S] */
    const double* psi;    /* solar zenith angle, radians                  (arg `psi`)                      */
    const double* K_b;    /* K_b_fn(psi) = G_fn(psi)/cos(psi)             (all schemes, e.g. _solve_2s.py:26) */
    const double* G;      /* G_fn(psi)                                    (4s: _solve_4s.py:145)            */
    const double* mu_bar; /* int cos/G sin  over the hemisphere           (2s: _solve_2s.py:32)             */
    const double* G_int;  /* [S][2] sector integrals of G over [0,mu_s],[mu_s,1]  (4s: _solve_4s.py:148-149) */
    const double* tau_i;  /* tau_df_fn(K_b_fn, mean dlai)                 (zq: _solve_zq.py:50-51);
                             zq_pa: tau_df_fn(K_b_fn, LAI / min(100, n_z))  (_solve_zq_pa.py:175)            */
    const double* tau_psi;/* tau_b_fn(K_b_fn, psi, mean dlai)             (zq: _solve_zq.py:52)             */
    const int32_t* lai_idx;  /* row of lai_lib      */
    const int32_t* leaf_idx; /* row of leaf_*_lib   */
    const int32_t* soil_idx; /* row of soil_r_lib   */
    const int32_t* sky_idx;  /* row of I_d*0_lib    */

    /* libraries */
    const double* lai_lib;    /* [n_lai][n_z]   cumulative LAI            (arg `lai`)                        */
    const double* tau_d_lev;  /* [n_lai][n_z]   bl: tau_df_fn(K_b_fn, lai[j]) (_solve_bl.py:35-37);
                                                n79: tau_df_fn(K_b_fn, dlai[j]), j < n_z-1 (_solve_n79.py:53) */
    const double* leaf_r_lib; /* [n_leaf][n_wl] (arg `leaf_r`) */
    const double* leaf_t_lib; /* [n_leaf][n_wl] (arg `leaf_t`) */
    const double* soil_r_lib; /* [n_soil][n_wl] (arg `soil_r`) */
    const double* I_dr0_lib;  /* [n_sky][n_wl]  (arg `I_dr0_all`, in-band W m-2) */
    const double* I_df0_lib;  /* [n_sky][n_wl]  (arg `I_df0_all`) */

    /* scalars */
    double mla_deg; /* mean leaf angle, degrees (2s arg `mla`, _solve_2s.py:28) */
    double mu_s;    /* 4s sector-dividing cosine (option `mu_s`, _solve_4s.py:9); 0 selects the default 0.501 */
} crt1d_batch;

/* Where results go.  Profiles are [S][n_z][n_wl].  For the closed-form schemes (2s, 4s, bl, bf, g77)
 * any profile pointer may be NULL: that field is then not written (reduced-diagnostic mode).
 * n79 and zq use I_dr / I_df_d / I_df_u / F as elimination scratch and require all four. */
typedef struct crt1d_out {
    double* I_dr;   /* direct beam irradiance            (RET_KEYS_ALL_SCHEMES, crt1d/solvers/__init__.py:34) */
    double* I_df_d; /* downward diffuse irradiance */
    double* I_df_u; /* upward diffuse irradiance   */
    double* F;      /* actinic flux                */
    /* scheme extras, same names as the reference's return dicts
     *   zq : x0 = I_df_d_ss, x1 = I_df_u_ss, x2 = F_ss     [S][n_z][n_wl]    (_solve_zq.py:221-229)
     *   bf : x0 = aI_lsl,    x1 = aI_lsh,    x2 = aI_l     [S][n_z][n_wl]    (_solve_bf.py:145-154)
     *   g77: x0 = aI_lsl,    x1 = aI_lsh,    x2 = aI_l     [S][n_z][n_wl]    (_solve_g77.py:127-135)
     *   n79: x0 = aI_lsl,    x1 = aI_lsh                   [S][n_z-1][n_wl]  (_solve_n79.py:157-164)  */
    double* x0;
    double* x1;
    double* x2;
    double* rho_c; /* bf: canopy reflectance per band [S][n_wl] (the reference returns the last band's scalar) */

    /* fused epilogue: canopy-integrated absorbed irradiance per spectral band group
     * (Sum over layers of model.py:606-609 `aI`, weighted Sum over wl as diagnostics.py:71-81) */
    const double* band_w; /* [n_bw][n_wl] weights, e.g. rows PAR and NIR from spectra._x_frac_in_bounds */
    int32_t n_bw;         /* 0..4 */
    double* absorbed;     /* [S][n_bw], or NULL */

    /* optional float32 STORAGE of the profiles (BASELINE north_star: "optional float32 path at <= 1e-5"):
     * 0 = float64 (the reference's layout, default); 1 = I_dr, I_df_d, I_df_u, F, x0, x1, x2 point to
     * float arrays of the same shapes.  All arithmetic stays float64 (fp64 coefficient stage AND level stage,
     * so the result is the float64 value rounded once: rel. err <= 6e-8); rho_c / absorbed stay float64.
     * Halves the HBM traffic of the write-bound schemes.  Not available for n79 and zq
     * (CRT1D_ERR_UNSUPPORTED): they park float64 elimination checkpoints in the profile arrays. */
    int32_t profile_f32;

    /* optional per-scenario status word [S] (int32; NULL = not wanted): the library zeroes it on the launch
     * stream and the kernels OR in CRT1D_STATUS_* bits.  Checked where every column is evaluated anyway: the
     * ground and top levels (all fields; a NaN/Inf coefficient reaches every level of its column). */
    int32_t* status;
} crt1d_out;

/* ---- library info ---------------------------------------------------------------------------- */
CRT1D_API int crt1d_abi_version(void);
CRT1D_API const char* crt1d_strerror(int code);
CRT1D_API const char* crt1d_last_error(void);
CRT1D_API int crt1d_device_count(int* n_devices); /* CRT1D_ERR_NO_DEVICE if none */

/* ---- solvers: device pointers, asynchronous on `stream` ------------------------------------- */
CRT1D_API int crt1d_solve(int scheme, const crt1d_batch* in, const crt1d_out* out, void* stream);
/* one named entry per reference solver function (thin aliases of crt1d_solve) */
CRT1D_API int crt1d_solve_2s(const crt1d_batch* in, const crt1d_out* out, void* stream);  /* replaces solve_2s  _solve_2s.py:11  */
CRT1D_API int crt1d_solve_4s(const crt1d_batch* in, const crt1d_out* out, void* stream);  /* replaces solve_4s  _solve_4s.py:8   */
CRT1D_API int crt1d_solve_bf(const crt1d_batch* in, const crt1d_out* out, void* stream);  /* replaces solve_bf  _solve_bf.py:7   */
CRT1D_API int crt1d_solve_bl(const crt1d_batch* in, const crt1d_out* out, void* stream);  /* replaces solve_bl  _solve_bl.py:9   */
CRT1D_API int crt1d_solve_g77(const crt1d_batch* in, const crt1d_out* out, void* stream); /* replaces solve_g77 _solve_g77.py:7  */
CRT1D_API int crt1d_solve_n79(const crt1d_batch* in, const crt1d_out* out, void* stream); /* replaces solve_n79 _solve_n79.py:11 */
CRT1D_API int crt1d_solve_zq(const crt1d_batch* in, const crt1d_out* out, void* stream);  /* replaces solve_zq  _solve_zq.py:13  */
CRT1D_API int crt1d_solve_zq_pa(const crt1d_batch* in, const crt1d_out* out, void* stream); /* replaces solve_zq_pa _solve_zq_pa.py:24 */

/* ---- solver: host pointers, synchronous (H2D + kernels + D2H inside) -------------------------
 * Any host memory works and any batch size: the scenarios go through two 128 MB device slots on two streams
 * (kernel of chunk k+1 overlaps the D2H of chunk k), so the call is PCIe-bound and not limited by HBM capacity.
 * Page-locked output buffers (cudaHostAlloc / cudaHostRegister) receive the DMA directly; pageable buffers are
 * filled from a page-locked staging ring by a few copy threads (first-touch page faults are their limit).
 * Returns CRT1D_NONFINITE (> 0) when results were delivered but some scenario holds NaN/Inf (out->status, if set,
 * says which).  Index arrays are not range-checked on this side of the ABI: every *_idx[s] must be a valid row
 * of its library. */
CRT1D_API int crt1d_solve_host(int scheme, const crt1d_batch* in_host, const crt1d_out* out_host, int device);
CRT1D_API int crt1d_release_workspace(void); /* frees the calling thread's cached device workspace, streams and staging ring */
/* Scenarios per `crt1d_solve` launch, at most `max_scen`, that fill whole waves of resident CTAs for the kernel the
 * library picks for (scheme, n_z, n_wl) on `device` (-1 = the current one): a caller that cuts a sweep into chunks
 * (the reference would loop `Model.run`, crt1d/model.py:298-320) should cut it into chunks of this size -- a launch
 * of 4.1 waves costs 5.  Returns the size (> 0) or a negative error code. */
CRT1D_API int64_t crt1d_preferred_batch(int scheme, int32_t n_z, int32_t n_wl, int64_t max_scen, int device);
/* test / tuning hook: re-read the CRT1D_B200_* kernel-selection environment variables (they are read once, when
 * the library is loaded; no getenv on the launch path). */
CRT1D_API int crt1d_reload_tuning(void);

/* ---- layer absorption  (replaces _calc_absorption, crt1d/model.py:573-647) -------------------
 * Inputs: profiles [S][n_z][n_wl], K_b [S], lai/leaf libraries + indices as in crt1d_batch.
 * Outputs (each [S][n_z-1][n_wl], any may be NULL): aI, aI_df, aI_dr, aI_sh, aI_sl, aI_df_sl, aI_df_sh.
 * Device pointers, asynchronous. */
typedef struct crt1d_absorption_out {
    double* aI;
    double* aI_df;
    double* aI_dr;
    double* aI_sh;
    double* aI_sl;
    double* aI_df_sl;
    double* aI_df_sh;
} crt1d_absorption_out;
CRT1D_API int crt1d_calc_absorption(const crt1d_batch* in, const double* I_dr, const double* I_df_d,
                          const double* I_df_u, const crt1d_absorption_out* out, void* stream);

/* ---- canopy energy balance per scenario and spectral band group ------------------------------
 * (replaces diagnostics.compare_ebal, crt1d/diagnostics.py:476-530, with the band() weights of
 * diagnostics.py:56-81).  Reads only the ground and top rows of the three profiles [S][n_z][n_wl].
 * ebal[S][n_bw][4] = { incoming  = Sum_wl w (I_dr + I_df_d)[top],
 *                      outgoing  = Sum_wl w  I_df_u[top]            (reflected),
 *                      soil_abs  = Sum_wl w ((I_dr + I_df_d)[0] - I_df_u[0]),
 *                      canopy_abs = Sum_wl w (I_df_d[top] - I_df_u[top] + I_dr[top] - I_dr[0] - (I_df_d[0] - I_df_u[0])) }
 * band_w: [n_bw][n_wl] weights (PAR/NIR/solar fractions, or photon-flux weights w / e_wl_umol); n_bw in 1..4.
 * Device pointers, asynchronous. */
CRT1D_API int crt1d_energy_balance(int64_t n_scen, int32_t n_z, int32_t n_wl, const double* I_dr,
                                   const double* I_df_d, const double* I_df_u, const double* band_w,
                                   int32_t n_bw, double* ebal, void* stream);

/* ---- leaf-angle kernels (device pointers, asynchronous) -------------------------------------
 * G(psi) and K_b = G/cos(psi) for n angles         (replaces leaf_angle.G_*, model.py:291)       */
CRT1D_API int crt1d_leaf_G(int family, double param, int64_t n, const double* psi, double* G, double* K_b,
                 void* stream);
/* tau_d(L) = 2 int_0^{pi/2} exp(-K_b(psi) L) sin cos dpsi  for n LAI values
 * (replaces common.tau_df_fn, crt1d/solvers/common.py:30-87).  n_quad > 0: Gauss-Legendre with n_quad
 * nodes (<= 128); n_quad == 0: the reference's 9-sector "9sky" rule.                              */
CRT1D_API int crt1d_tau_d(int family, double param, int n_quad, int64_t n, const double* L, double* tau_d,
                void* stream);
/* scalars that depend on the leaf-angle family only: out[0] = mu_bar (2s), out[1], out[2] = G sector
 * integrals for `mu_s` (4s); Gauss-Legendre with n_quad nodes.  `out` is a DEVICE pointer to 3 doubles. */
CRT1D_API int crt1d_leaf_integrals(int family, double param, double mu_s, int n_quad, double* out, void* stream);

/* ---- spectral binning (device pointers, asynchronous) ---------------------------------------------
 * out[r][i] = trapezoidally integrated average of y_r(x) over [bins[i], bins[i+1]], i < n_bins, for n_rows
 * spectra y[n_rows][n_x] on one ascending grid x[n_x]; bins[n_bins + 1] ascending
 * (replaces spectra.smear_tuv / _smear_tuv_1, crt1d/spectra.py:221-300).            */
CRT1D_API int crt1d_smear_tuv(int64_t n_rows, int32_t n_x, const double* x, const double* y, int32_t n_bins,
                              const double* bins, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRT1D_B200_H */
