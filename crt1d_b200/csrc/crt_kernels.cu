// crt_kernels.cu -- sm_100a kernels of the crt1d solver hot path and their launchers.
//
// Decomposition (all schemes): one CTA = one scenario x one tile of BLOCK*VEC adjacent wavelength
// bands; one thread = VEC adjacent bands ("columns"), all n_z levels.  Bands are the fastest axis of
// the reference's (n_z, n_wl) output layout, so at every level a warp writes VEC*32 consecutive
// doubles per field (256 or 512 contiguous bytes, 16-byte vector stores when VEC = 2): the output
// stream -- 32 B per layer.band for 2s/4s and the dominant HBM traffic -- is fully coalesced and
// write-once (st.global.cs, evict-first: nothing is re-read, keep it out of L2's way).
// The scenario's band-independent level tables (what each reference solver evaluates before its band
// loop) are built cooperatively in shared memory once per CTA.  No tensor cores: nothing here is a
// dense contraction; the kernels are bound by HBM writes and the FP64 pipe.
#include <cuda_runtime.h>

#include "crt_internal.h"
#include "crt_scheme.cuh"

namespace crt {

constexpr int BLOCK = 128;

// ---------------------------------------------------------------------------------------------
// global-memory column accessor
// ---------------------------------------------------------------------------------------------
template <int VEC>
struct GlobalOut {
    double* p[N_FIELDS];  // pre-offset to (scenario, level 0, first band of this thread); nullptr = skip
    int64_t stride;       // doubles between consecutive levels (= n_wl)

    // final results: written once, never re-read by this kernel -> streaming (evict-first) stores
    __device__ __forceinline__ void st(int f, int j, const double (&x)[VEC]) const {
        double* q = p[f];
        if (q == nullptr) return;
        q += (int64_t)j * stride;
        if constexpr (VEC == 2) {
            __stcs(reinterpret_cast<double2*>(q), make_double2(x[0], x[1]));
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) __stcs(q + v, x[v]);
        }
    }
    // elimination scratch parked in the output arrays: re-read by the same thread during
    // back-substitution -> default (write-back, L2-resident) stores
    __device__ __forceinline__ void st_tmp(int f, int j, const double (&x)[VEC]) const {
        double* q = p[f] + (int64_t)j * stride;
        if constexpr (VEC == 2) {
            *reinterpret_cast<double2*>(q) = make_double2(x[0], x[1]);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) q[v] = x[v];
        }
    }
    __device__ __forceinline__ void ld_tmp(int f, int j, double (&x)[VEC]) const {
        const double* q = p[f] + (int64_t)j * stride;
        if constexpr (VEC == 2) {
            const double2 t = __ldcs(reinterpret_cast<const double2*>(q));  // last use: evict-first
            x[0] = t.x;
            x[1] = t.y;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[v] = __ldcs(q + v);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// the solver kernel
// ---------------------------------------------------------------------------------------------
template <int SCHEME, int VEC>
__global__ void __launch_bounds__(BLOCK) solve_kernel(const crt1d_batch in, const crt1d_out out, int tiles_per_scen,
                                                      int tiles_per_cta) {
    extern __shared__ double tab[];
    __shared__ double red[BLOCK / 32][4];

    const int ctas_per_scen = (tiles_per_scen + tiles_per_cta - 1) / tiles_per_cta;
    const int64_t s = blockIdx.x / ctas_per_scen;
    const int t0 = (blockIdx.x % ctas_per_scen) * tiles_per_cta;
    const int t1 = min(t0 + tiles_per_cta, tiles_per_scen);
    const int n_z = in.n_z, n_wl = in.n_wl;

    for (int j = threadIdx.x; j < n_z; j += BLOCK) fill_level_tables<SCHEME>(in, s, j, tab);
    __syncthreads();

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t prof = (int64_t)n_z * n_wl;                       // doubles per scenario in a profile
    const int64_t xprof = (int64_t)extra_rows(SCHEME, n_z) * n_wl;  // ... in an extra-output slot

    for (int t = t0; t < t1; ++t) {
        const int b0 = (t * BLOCK + threadIdx.x) * VEC;
        if (b0 >= n_wl) continue;
        const BandIn<VEC> b = load_bands<VEC>(in, s, b0);
        GlobalOut<VEC> o;
        o.stride = n_wl;
        o.p[F_IDR] = out.I_dr ? out.I_dr + s * prof + b0 : nullptr;
        o.p[F_DN] = out.I_df_d ? out.I_df_d + s * prof + b0 : nullptr;
        o.p[F_UP] = out.I_df_u ? out.I_df_u + s * prof + b0 : nullptr;
        o.p[F_F] = out.F ? out.F + s * prof + b0 : nullptr;
        o.p[F_X0] = out.x0 ? out.x0 + s * xprof + b0 : nullptr;
        o.p[F_X1] = out.x1 ? out.x1 + s * xprof + b0 : nullptr;
        o.p[F_X2] = out.x2 ? out.x2 + s * xprof + b0 : nullptr;
        double rho_c[VEC], ab[VEC];
        solve_column_group<SCHEME, VEC>(in, s, tab, b, o, rho_c, ab);
        if constexpr (SCHEME == CRT1D_SCHEME_BF) {
            if (out.rho_c) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) out.rho_c[s * n_wl + b0 + v] = rho_c[v];
            }
        }
        if (out.absorbed) {
            for (int k = 0; k < out.n_bw; ++k) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[k] += out.band_w[(int64_t)k * n_wl + b0 + v] * ab[v];
            }
        }
    }

    if (out.absorbed) {  // fixed-order block reduction (deterministic); launcher guarantees ctas_per_scen == 1
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = acc[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < out.n_bw) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; ++w) v += red[w][threadIdx.x];
            out.absorbed[s * out.n_bw + threadIdx.x] = v;
        }
    }
}

template <int SCHEME, int VEC>
static cudaError_t launch_one(const crt1d_batch& in, const crt1d_out& out, cudaStream_t stream) {
    const int cols = BLOCK * VEC;
    const int tiles_per_scen = (in.n_wl + cols - 1) / cols;
    const int tiles_per_cta = out.absorbed ? tiles_per_scen : 1;
    const int ctas_per_scen = (tiles_per_scen + tiles_per_cta - 1) / tiles_per_cta;
    const int64_t grid = in.n_scen * ctas_per_scen;
    if (grid <= 0 || grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)n_level_tables(SCHEME) * in.n_z * sizeof(double);
    auto kern = solve_kernel<SCHEME, VEC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid, BLOCK, smem, stream>>>(in, out, tiles_per_scen, tiles_per_cta);
    return cudaGetLastError();
}

template <int SCHEME>
static cudaError_t launch_vec(const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    return vec2 ? launch_one<SCHEME, 2>(in, out, stream) : launch_one<SCHEME, 1>(in, out, stream);
}

size_t solve_shared_bytes(int scheme, int n_z) { return (size_t)n_level_tables(scheme) * n_z * sizeof(double); }

cudaError_t launch_solve(int scheme, const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    switch (scheme) {
        case CRT1D_SCHEME_2S: return launch_vec<CRT1D_SCHEME_2S>(in, out, vec2, stream);
        case CRT1D_SCHEME_4S: return launch_vec<CRT1D_SCHEME_4S>(in, out, vec2, stream);
        case CRT1D_SCHEME_BF: return launch_vec<CRT1D_SCHEME_BF>(in, out, vec2, stream);
        case CRT1D_SCHEME_BL: return launch_vec<CRT1D_SCHEME_BL>(in, out, vec2, stream);
        case CRT1D_SCHEME_G77: return launch_vec<CRT1D_SCHEME_G77>(in, out, vec2, stream);
        case CRT1D_SCHEME_N79: return launch_vec<CRT1D_SCHEME_N79>(in, out, vec2, stream);
        case CRT1D_SCHEME_ZQ: return launch_vec<CRT1D_SCHEME_ZQ>(in, out, vec2, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------
// layer absorption  (ref ../model.py:573-647)
// thread = VEC adjacent bands, walks up the levels carrying the level below in registers, so every
// profile value is read exactly once; per-layer scalars (1 - tau_b, f_sl) sit in shared memory.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_vec(const double* q, double (&x)[VEC]) {
    if constexpr (VEC == 2) {
        const double2 t = __ldcs(reinterpret_cast<const double2*>(q));
        x[0] = t.x;
        x[1] = t.y;
    } else {
        x[0] = __ldcs(q);
    }
}
template <int VEC>
__device__ __forceinline__ void st_vec(double* base, int64_t off, const double (&x)[VEC]) {
    if (base == nullptr) return;
    if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<double2*>(base + off), make_double2(x[0], x[1]));
    } else {
        __stcs(base + off, x[0]);
    }
}

template <int VEC>
__global__ void __launch_bounds__(BLOCK) absorption_kernel(const crt1d_batch in, const double* __restrict__ I_dr,
                                                           const double* __restrict__ I_df_d,
                                                           const double* __restrict__ I_df_u,
                                                           const crt1d_absorption_out out, int tiles_per_scen) {
    extern __shared__ double tab[];  // [0, n_z-1): 1 - exp(-K_b dlai);  [n_z, 2n_z-1): f_sl
    const int64_t s = blockIdx.x / tiles_per_scen;
    const int t = blockIdx.x % tiles_per_scen;
    const int n_z = in.n_z, n_wl = in.n_wl;
    const double K_b = in.K_b[s];
    const double* lai = in.lai_lib + (int64_t)in.lai_idx[s] * n_z;
    for (int i = threadIdx.x; i < n_z - 1; i += BLOCK) {
        tab[i] = 1.0 - exp(-K_b * (lai[i] - lai[i + 1]));           // ref :617-619
        tab[n_z + i] = exp(-K_b * ((lai[i] + lai[i + 1]) / 2.0));   // ref :601-602
    }
    __syncthreads();
    const int b0 = (t * BLOCK + threadIdx.x) * VEC;
    if (b0 >= n_wl) return;
    double leaf_a[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int64_t k = (int64_t)in.leaf_idx[s] * n_wl + b0 + v;
        leaf_a[v] = 1.0 - (in.leaf_r_lib[k] + in.leaf_t_lib[k]);    // ref :585
    }
    const int64_t pin = s * (int64_t)n_z * n_wl + b0;
    const int64_t pout = s * (int64_t)(n_z - 1) * n_wl + b0;
    double dr0[VEC], dn0[VEC], up0[VEC], dr1[VEC], dn1[VEC], up1[VEC];
    ld_vec<VEC>(I_dr + pin, dr0);
    ld_vec<VEC>(I_df_d + pin, dn0);
    ld_vec<VEC>(I_df_u + pin, up0);
    for (int i = 0; i < n_z - 1; ++i) {
        const int64_t o1 = pin + (int64_t)(i + 1) * n_wl;
        ld_vec<VEC>(I_dr + o1, dr1);
        ld_vec<VEC>(I_df_d + o1, dn1);
        ld_vec<VEC>(I_df_u + o1, up1);
        const double omtb = tab[i], fsl = tab[n_z + i], fsh = 1.0 - fsl;
        double a[VEC], a_dr[VEC], a_df[VEC], a_df_sl[VEC], a_df_sh[VEC], a_sl[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            a[v] = dr1[v] - dr0[v] + dn1[v] - dn0[v] + up0[v] - up1[v];  // ref :606-609
            a_dr[v] = dr1[v] * omtb * leaf_a[v];                         // ref :617-621
            a_df[v] = a[v] - a_dr[v];                                    // ref :628
            a_df_sl[v] = a_df[v] * fsl;                                  // ref :631-632
            a_df_sh[v] = a_df[v] * fsh;
            a_sl[v] = a_df_sl[v] + a_dr[v];                              // ref :633
            dr0[v] = dr1[v];
            dn0[v] = dn1[v];
            up0[v] = up1[v];
        }
        const int64_t o = pout + (int64_t)i * n_wl;
        st_vec<VEC>(out.aI, o, a);
        st_vec<VEC>(out.aI_df, o, a_df);
        st_vec<VEC>(out.aI_dr, o, a_dr);
        st_vec<VEC>(out.aI_sh, o, a_df_sh);  // ref :634: a_sh = a_df_sh
        st_vec<VEC>(out.aI_sl, o, a_sl);
        st_vec<VEC>(out.aI_df_sl, o, a_df_sl);
        st_vec<VEC>(out.aI_df_sh, o, a_df_sh);
    }
}

cudaError_t launch_absorption(const crt1d_batch& in, const double* I_dr, const double* I_df_d, const double* I_df_u,
                              const crt1d_absorption_out& out, bool vec2, cudaStream_t stream) {
    const int vec = vec2 ? 2 : 1;
    const int tiles = (in.n_wl + BLOCK * vec - 1) / (BLOCK * vec);
    const int64_t grid = in.n_scen * tiles;
    if (grid <= 0 || grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    const size_t smem = 2 * (size_t)in.n_z * sizeof(double);
    if (vec2) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(absorption_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        absorption_kernel<2><<<(unsigned)grid, BLOCK, smem, stream>>>(in, I_dr, I_df_d, I_df_u, out, tiles);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(absorption_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        absorption_kernel<1><<<(unsigned)grid, BLOCK, smem, stream>>>(in, I_dr, I_df_d, I_df_u, out, tiles);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// leaf-angle kernels  (ref ../leaf_angle.py:118-202, common.py:11-95)
// ---------------------------------------------------------------------------------------------
__global__ void leaf_G_kernel(int family, double param, int64_t n, const double* __restrict__ psi, double* G,
                              double* K_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double g = leaf_G(family, param, psi[i]);
    if (G) G[i] = g;
    if (K_b) K_b[i] = g / cos(psi[i]);  // ref ../model.py:291
}

cudaError_t launch_leaf_G(int family, double param, int64_t n, const double* psi, double* G, double* K_b,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    leaf_G_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(family, param, n, psi, G, K_b);
    return cudaGetLastError();
}

// tau_d(L) = 2 int_0^{pi/2} exp(-K_b(psi) L) sin(psi) cos(psi) dpsi   (ref common.py:30-37)
__global__ void tau_d_kernel(int family, double param, QuadRule rule, int64_t n, const double* __restrict__ L,
                             double* tau_d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tau_d[i] = tau_d_quadrature(family, param, rule, L[i]);
}

cudaError_t launch_tau_d(int family, double param, const QuadRule& rule, int64_t n, const double* L, double* tau_d,
                         cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    tau_d_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(family, param, rule, n, L, tau_d);
    return cudaGetLastError();
}

__global__ void leaf_integrals_kernel(int family, double param, double mu_s, QuadRule rule, double* out) {
    if (threadIdx.x < 3 && blockIdx.x == 0) out[threadIdx.x] = leaf_integral(family, param, mu_s, rule, threadIdx.x);
}

cudaError_t launch_leaf_integrals(int family, double param, double mu_s, const QuadRule& rule, double* out,
                                  cudaStream_t stream) {
    leaf_integrals_kernel<<<1, 32, 0, stream>>>(family, param, mu_s, rule, out);
    return cudaGetLastError();
}

}  // namespace crt
