// crt_internal.h -- C++ interface between the C ABI (crt_abi.cu) and the kernel launchers (crt_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "../../include/crt1d_b200.h"
#include "crt_leafangle.cuh"

namespace crt {

size_t solve_shared_bytes(int scheme, int n_z);
int64_t preferred_batch(int scheme, int n_z, int n_wl, int64_t max_scen, int n_sm);
void reload_tuning();  // re-read the CRT1D_B200_* tuning variables (read once at load otherwise)
cudaError_t launch_solve(int scheme, const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream);
cudaError_t launch_absorption(const crt1d_batch& in, const double* I_dr, const double* I_df_d, const double* I_df_u,
                              const crt1d_absorption_out& out, bool vec2, cudaStream_t stream);
cudaError_t launch_energy_balance(int64_t n_scen, int n_z, int n_wl, const double* I_dr, const double* I_df_d,
                                  const double* I_df_u, const double* band_w, int n_bw, double* ebal, cudaStream_t stream);
cudaError_t launch_leaf_G(int family, double param, int64_t n, const double* psi, double* G, double* K_b,
                          cudaStream_t stream);
cudaError_t launch_tau_d(int family, double param, const QuadRule& rule, int64_t n, const double* L, double* tau_d,
                         cudaStream_t stream);
cudaError_t launch_leaf_integrals(int family, double param, double mu_s, const QuadRule& rule, double* out,
                                  cudaStream_t stream);
cudaError_t launch_smear_tuv(int64_t n_rows, int n_x, const double* x, const double* y, int n_bins, const double* bins,
                             double* out, cudaStream_t stream);

}  // namespace crt
