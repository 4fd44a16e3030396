/* c_abi_example.c -- a plain-C caller of the crt1d_b200 C ABI (no Python, no torch, no CUDA headers).
 * Reads one scenario from a small binary file written by the test, calls crt1d_solve_host for the 2s
 * scheme with HOST buffers and writes the four profiles back.  Built and run by
 * tests/test_gpu_parity.py::test_plain_c_caller.
 *
 * file layout (little-endian): int32 n_z, n_wl; double psi, K_b, mu_bar, mla;
 *   double lai[n_z], leaf_r[n_wl], leaf_t[n_wl], soil_r[n_wl], I_dr0[n_wl], I_df0[n_wl]            */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/crt1d_b200.h"

static double* rd(FILE* f, size_t n) {
    double* p = (double*)malloc(n * sizeof(double));
    if (!p || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    int32_t dims[2];
    double sc[4];
    if (fread(dims, sizeof(int32_t), 2, f) != 2 || fread(sc, sizeof(double), 4, f) != 4) return 2;
    const int n_z = dims[0], n_wl = dims[1];
    double *lai = rd(f, n_z), *lr = rd(f, n_wl), *lt = rd(f, n_wl), *sr = rd(f, n_wl), *dr = rd(f, n_wl), *df = rd(f, n_wl);
    fclose(f);

    if (crt1d_abi_version() != CRT1D_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }
    int32_t zero = 0;
    crt1d_batch b;
    memset(&b, 0, sizeof b);
    b.n_scen = 1; b.n_z = n_z; b.n_wl = n_wl; b.n_lai = b.n_leaf = b.n_soil = b.n_sky = 1;
    b.psi = &sc[0]; b.K_b = &sc[1]; b.mu_bar = &sc[2]; b.mla_deg = sc[3];
    b.lai_idx = b.leaf_idx = b.soil_idx = b.sky_idx = &zero;
    b.lai_lib = lai; b.leaf_r_lib = lr; b.leaf_t_lib = lt; b.soil_r_lib = sr; b.I_dr0_lib = dr; b.I_df0_lib = df;

    const size_t n = (size_t)n_z * n_wl;
    double* res = (double*)malloc(4 * n * sizeof(double));
    crt1d_out o;
    memset(&o, 0, sizeof o);
    o.I_dr = res; o.I_df_d = res + n; o.I_df_u = res + 2 * n; o.F = res + 3 * n;

    int rc = crt1d_solve_host(CRT1D_SCHEME_2S, &b, &o, 0);
    if (rc != CRT1D_OK) { fprintf(stderr, "crt1d_solve_host: %d %s: %s\n", rc, crt1d_strerror(rc), crt1d_last_error()); return 4; }
    /* an invalid call must come back as an error code, not a crash */
    b.mu_bar = NULL;
    if (crt1d_solve_host(CRT1D_SCHEME_2S, &b, &o, 0) != CRT1D_ERR_NULL_POINTER) return 5;
    crt1d_release_workspace();

    f = fopen(argv[2], "wb");
    if (!f || fwrite(res, sizeof(double), 4 * n, f) != 4 * n) return 6;
    fclose(f);
    printf("ok %d x %d\n", n_z, n_wl);
    return 0;
}
