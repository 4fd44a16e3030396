// crt_abi.cu -- the extern "C" boundary declared in include/crt1d_b200.h: argument validation, error
// reporting, the device-pointer entry points, and the host-pointer path (H2D + kernels + D2H).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "crt_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(CRT1D_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

const char* scheme_name(int scheme) {
    static const char* names[CRT1D_N_SCHEMES] = {"2s", "4s", "bf", "bl", "g77", "n79", "zq", "zq_pa"};
    return (scheme >= 0 && scheme < CRT1D_N_SCHEMES) ? names[scheme] : "?";
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Shape / pointer validation shared by the device and host paths.  Never dereferences data pointers.
int validate(int scheme, const crt1d_batch* in, const crt1d_out* out) {
    if (scheme < 0 || scheme >= CRT1D_N_SCHEMES) return fail(CRT1D_ERR_INVALID_ARG, "unknown scheme id " + std::to_string(scheme));
    if (in == nullptr || out == nullptr) return fail(CRT1D_ERR_NULL_POINTER, "crt1d_batch / crt1d_out pointer is NULL");
    const std::string sn = scheme_name(scheme);
    if (in->n_scen < 0 || in->n_wl < 1 || in->n_z < 2)
        return fail(CRT1D_ERR_INVALID_ARG, sn + ": need n_scen >= 0, n_wl >= 1, n_z >= 2");
    if (scheme == CRT1D_SCHEME_N79 && in->n_z < 3)
        return fail(CRT1D_ERR_INVALID_ARG, "n79: needs n_z >= 3 (the reference indexes td[1], _solve_n79.py:85)");
    if (in->n_lai < 1 || in->n_leaf < 1 || in->n_sky < 1)
        return fail(CRT1D_ERR_INVALID_ARG, sn + ": library row counts must be >= 1");
    if (in->n_scen == 0) return CRT1D_OK;
#define NEED(field)                                                                                     \
    if (in->field == nullptr) return fail(CRT1D_ERR_NULL_POINTER, sn + ": crt1d_batch." #field " is required")
    NEED(psi);
    NEED(K_b);
    NEED(lai_idx);
    NEED(leaf_idx);
    NEED(sky_idx);
    NEED(lai_lib);
    NEED(leaf_r_lib);
    NEED(leaf_t_lib);
    NEED(I_dr0_lib);
    NEED(I_df0_lib);
    if (scheme != CRT1D_SCHEME_BL) {
        NEED(soil_r_lib);
        NEED(soil_idx);
        if (in->n_soil < 1) return fail(CRT1D_ERR_INVALID_ARG, sn + ": n_soil must be >= 1");
    }
    if (scheme == CRT1D_SCHEME_2S) NEED(mu_bar);
    if (scheme == CRT1D_SCHEME_4S) NEED(G_int);
    if (scheme == CRT1D_SCHEME_ZQ) {
        NEED(tau_i);
        NEED(tau_psi);
    }
    if (scheme == CRT1D_SCHEME_ZQ_PA) NEED(tau_i);
    if (scheme == CRT1D_SCHEME_BL || scheme == CRT1D_SCHEME_N79) NEED(tau_d_lev);
#undef NEED
    if (scheme == CRT1D_SCHEME_N79 || scheme == CRT1D_SCHEME_ZQ) {
        if (!out->I_dr || !out->I_df_d || !out->I_df_u || !out->F)
            return fail(CRT1D_ERR_NULL_POINTER, sn + ": I_dr, I_df_d, I_df_u and F are all required (used as elimination scratch)");
    }
    if (out->profile_f32 != 0 && out->profile_f32 != 1) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_out.profile_f32 must be 0 or 1");
    if (out->profile_f32 && (scheme == CRT1D_SCHEME_N79 || scheme == CRT1D_SCHEME_ZQ))
        return fail(CRT1D_ERR_UNSUPPORTED, sn + ": float32 profile storage is not available (float64 elimination checkpoints live in the profile arrays)");
    if (out->n_bw < 0 || out->n_bw > 4) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_out.n_bw must be in 0..4");
    if (out->absorbed != nullptr && (out->band_w == nullptr || out->n_bw == 0))
        return fail(CRT1D_ERR_NULL_POINTER, "crt1d_out.absorbed needs band_w and n_bw >= 1");
    if (crt::solve_shared_bytes(scheme, in->n_z) > 227u * 1024u)
        return fail(CRT1D_ERR_UNSUPPORTED, sn + ": n_z = " + std::to_string(in->n_z) + " needs more than 227 KB of shared-memory level tables");
    return CRT1D_OK;
}

bool can_vec2(const crt1d_batch* in, const crt1d_out* out) {
    if (in->n_wl % 2 != 0) return false;
    const void* ptrs[] = {out->I_dr, out->I_df_d, out->I_df_u, out->F, out->x0, out->x1, out->x2};
    for (const void* p : ptrs)
        if (p != nullptr && !aligned16(p)) return false;  // (8-byte alignment would do for float pairs; keep one rule)
    return true;
}

// ---- per-thread device workspace of the host path ------------------------------------------------
struct Workspace {
    int device = -1;
    void* ptr = nullptr;
    size_t bytes = 0;
};
thread_local Workspace g_ws;

int ws_reserve(int device, size_t bytes) {
    if (g_ws.ptr != nullptr && g_ws.device == device && g_ws.bytes >= bytes) return CRT1D_OK;
    if (g_ws.ptr != nullptr) {
        cudaFree(g_ws.ptr);
        g_ws = Workspace();
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CRT1D_ERR_NO_MEMORY, "cudaMalloc of " + std::to_string(bytes) + " B workspace failed: " + cudaGetErrorString(e));
    }
    g_ws.device = device;
    g_ws.ptr = p;
    g_ws.bytes = bytes;
    return CRT1D_OK;
}

// bump allocator over the workspace, 256-byte aligned slices
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    static size_t pad(size_t n) { return (n + 255) & ~size_t(255); }
    template <class T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += pad(count * sizeof(T));
        return p;
    }
};

}  // namespace

extern "C" {

int crt1d_abi_version(void) { return CRT1D_ABI_VERSION; }

const char* crt1d_strerror(int code) {
    switch (code) {
        case CRT1D_OK: return "ok";
        case CRT1D_ERR_INVALID_ARG: return "invalid argument";
        case CRT1D_ERR_NULL_POINTER: return "required pointer is NULL";
        case CRT1D_ERR_UNSUPPORTED: return "unsupported configuration";
        case CRT1D_ERR_CUDA: return "CUDA runtime error";
        case CRT1D_ERR_NO_DEVICE: return "no CUDA device";
        case CRT1D_ERR_NO_MEMORY: return "device memory allocation failed";
        default: return "unknown error code";
    }
}

const char* crt1d_last_error(void) { return g_last_error.c_str(); }

int crt1d_device_count(int* n_devices) {
    if (n_devices == nullptr) return fail(CRT1D_ERR_NULL_POINTER, "n_devices is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        *n_devices = 0;
        return fail(CRT1D_ERR_NO_DEVICE, std::string("no CUDA device visible: ") + cudaGetErrorString(e));
    }
    *n_devices = n;
    return CRT1D_OK;
}

int crt1d_solve(int scheme, const crt1d_batch* in, const crt1d_out* out, void* stream) {
    int rc = validate(scheme, in, out);
    if (rc != CRT1D_OK) return rc;
    if (in->n_scen == 0) return CRT1D_OK;
    cudaError_t e = crt::launch_solve(scheme, *in, *out, can_vec2(in, out), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, (std::string("launch of ") + scheme_name(scheme) + " kernel").c_str());
    return CRT1D_OK;
}

int crt1d_solve_2s(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_2S, in, out, stream); }
int crt1d_solve_4s(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_4S, in, out, stream); }
int crt1d_solve_bf(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_BF, in, out, stream); }
int crt1d_solve_bl(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_BL, in, out, stream); }
int crt1d_solve_g77(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_G77, in, out, stream); }
int crt1d_solve_n79(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_N79, in, out, stream); }
int crt1d_solve_zq(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_ZQ, in, out, stream); }
int crt1d_solve_zq_pa(const crt1d_batch* in, const crt1d_out* out, void* stream) { return crt1d_solve(CRT1D_SCHEME_ZQ_PA, in, out, stream); }

int crt1d_calc_absorption(const crt1d_batch* in, const double* I_dr, const double* I_df_d, const double* I_df_u,
                          const crt1d_absorption_out* out, void* stream) {
    if (in == nullptr || out == nullptr || I_dr == nullptr || I_df_d == nullptr || I_df_u == nullptr)
        return fail(CRT1D_ERR_NULL_POINTER, "crt1d_calc_absorption: NULL argument");
    if (in->n_scen < 0 || in->n_wl < 1 || in->n_z < 2) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_calc_absorption: bad sizes");
    if (in->n_scen == 0) return CRT1D_OK;
    if (!in->K_b || !in->lai_idx || !in->lai_lib || !in->leaf_idx || !in->leaf_r_lib || !in->leaf_t_lib)
        return fail(CRT1D_ERR_NULL_POINTER, "crt1d_calc_absorption: K_b, lai_idx, lai_lib, leaf_idx, leaf_r_lib, leaf_t_lib are required");
    if (2u * (size_t)in->n_z * sizeof(double) > 227u * 1024u) return fail(CRT1D_ERR_UNSUPPORTED, "crt1d_calc_absorption: n_z too large");
    bool vec2 = (in->n_wl % 2 == 0) && aligned16(I_dr) && aligned16(I_df_d) && aligned16(I_df_u);
    const void* outs[] = {out->aI, out->aI_df, out->aI_dr, out->aI_sh, out->aI_sl, out->aI_df_sl, out->aI_df_sh};
    for (const void* p : outs)
        if (p != nullptr && !aligned16(p)) vec2 = false;
    cudaError_t e = crt::launch_absorption(*in, I_dr, I_df_d, I_df_u, *out, vec2, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of absorption kernel");
    return CRT1D_OK;
}

int crt1d_energy_balance(int64_t n_scen, int32_t n_z, int32_t n_wl, const double* I_dr, const double* I_df_d,
                         const double* I_df_u, const double* band_w, int32_t n_bw, double* ebal, void* stream) {
    if (n_scen < 0 || n_z < 2 || n_wl < 1) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_energy_balance: bad sizes");
    if (n_bw < 1 || n_bw > 4) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_energy_balance: n_bw must be in 1..4");
    if (n_scen == 0) return CRT1D_OK;
    if (!I_dr || !I_df_d || !I_df_u || !band_w || !ebal) return fail(CRT1D_ERR_NULL_POINTER, "crt1d_energy_balance: NULL argument");
    cudaError_t e = crt::launch_energy_balance(n_scen, n_z, n_wl, I_dr, I_df_d, I_df_u, band_w, n_bw, ebal,
                                               static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of energy_balance kernel");
    return CRT1D_OK;
}

int crt1d_leaf_G(int family, double param, int64_t n, const double* psi, double* G, double* K_b, void* stream) {
    if (family < 0 || family > CRT1D_G_ELLIPSOIDAL_APPROX_BONAN) return fail(CRT1D_ERR_INVALID_ARG, "unknown leaf-angle family");
    if (n < 0) return fail(CRT1D_ERR_INVALID_ARG, "n < 0");
    if (n > 0 && psi == nullptr) return fail(CRT1D_ERR_NULL_POINTER, "psi is NULL");
    cudaError_t e = crt::launch_leaf_G(family, param, n, psi, G, K_b, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of leaf_G kernel");
    return CRT1D_OK;
}

static int make_rule(int n_quad, crt::QuadRule* rule) {
    if (n_quad < 0 || n_quad > 128) return fail(CRT1D_ERR_INVALID_ARG, "n_quad must be in 0..128");
    memset(rule, 0, sizeof(*rule));
    rule->n = n_quad;
    rule->panels = 12;  // 12 graded panels x n_quad nodes: ~1e-16 abs error down to L = 0.004
    if (n_quad > 0) crt::gauss_legendre(n_quad, rule->x, rule->w);
    return CRT1D_OK;
}

int crt1d_tau_d(int family, double param, int n_quad, int64_t n, const double* L, double* tau_d, void* stream) {
    if (family < 0 || family > CRT1D_G_ELLIPSOIDAL_APPROX_BONAN) return fail(CRT1D_ERR_INVALID_ARG, "unknown leaf-angle family");
    if (n < 0) return fail(CRT1D_ERR_INVALID_ARG, "n < 0");
    if (n > 0 && (L == nullptr || tau_d == nullptr)) return fail(CRT1D_ERR_NULL_POINTER, "L / tau_d is NULL");
    crt::QuadRule rule;
    int rc = make_rule(n_quad, &rule);
    if (rc != CRT1D_OK) return rc;
    cudaError_t e = crt::launch_tau_d(family, param, rule, n, L, tau_d, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of tau_d kernel");
    return CRT1D_OK;
}

int crt1d_smear_tuv(int64_t n_rows, int32_t n_x, const double* x, const double* y, int32_t n_bins, const double* bins,
                    double* out, void* stream) {
    if (n_rows < 0 || n_x < 2 || n_bins < 0) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_smear_tuv: need n_rows >= 0, n_x >= 2, n_bins >= 0");
    if (n_rows * (int64_t)n_bins > 2147483647LL * 128) return fail(CRT1D_ERR_INVALID_ARG, "crt1d_smear_tuv: too many (row, bin) pairs");
    if (n_rows > 0 && n_bins > 0 && (x == nullptr || y == nullptr || bins == nullptr || out == nullptr))
        return fail(CRT1D_ERR_NULL_POINTER, "crt1d_smear_tuv: x / y / bins / out is NULL");
    cudaError_t e = crt::launch_smear_tuv(n_rows, n_x, x, y, n_bins, bins, out, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of smear_tuv kernel");
    return CRT1D_OK;
}

int crt1d_leaf_integrals(int family, double param, double mu_s, int n_quad, double* out, void* stream) {
    if (family < 0 || family > CRT1D_G_ELLIPSOIDAL_APPROX_BONAN) return fail(CRT1D_ERR_INVALID_ARG, "unknown leaf-angle family");
    if (out == nullptr) return fail(CRT1D_ERR_NULL_POINTER, "out is NULL");
    if (n_quad < 1) return fail(CRT1D_ERR_INVALID_ARG, "n_quad must be in 1..128");
    if (!(mu_s > 0.0 && mu_s < 1.0)) return fail(CRT1D_ERR_INVALID_ARG, "mu_s must be in (0, 1)");
    crt::QuadRule rule;
    int rc = make_rule(n_quad, &rule);
    if (rc != CRT1D_OK) return rc;
    cudaError_t e = crt::launch_leaf_integrals(family, param, mu_s, rule, out, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "launch of leaf_integrals kernel");
    return CRT1D_OK;
}

int crt1d_release_workspace(void) {
    if (g_ws.ptr != nullptr) {
        cudaSetDevice(g_ws.device);
        cudaFree(g_ws.ptr);
    }
    g_ws = Workspace();
    return CRT1D_OK;
}

// Host-pointer path: what a non-CUDA caller (the reference's Python, via ctypes) binds.
int crt1d_solve_host(int scheme, const crt1d_batch* in, const crt1d_out* out, int device) {
    int rc = validate(scheme, in, out);
    if (rc != CRT1D_OK) return rc;
    if (in->n_scen == 0) return CRT1D_OK;
    int n_dev = 0;
    rc = crt1d_device_count(&n_dev);
    if (rc != CRT1D_OK) return rc;
    if (device < 0 || device >= n_dev) return fail(CRT1D_ERR_INVALID_ARG, "device index out of range");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

    const size_t S = (size_t)in->n_scen, nz = (size_t)in->n_z, nw = (size_t)in->n_wl;
    const size_t prof = S * nz * nw;
    const size_t xrows = (scheme == CRT1D_SCHEME_N79) ? nz - 1 : nz;
    const size_t xprof = S * xrows * nw;

    struct Copy {
        void* dst;
        const void* src;
        size_t bytes;
    };
    std::vector<Copy> h2d, d2h;
    // pass 1: size; pass 2: carve.  (two passes keep the carving code in one place)
    crt1d_batch din = *in;
    crt1d_out dout = *out;
    size_t need = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            rc = ws_reserve(device, need);
            if (rc != CRT1D_OK) return rc;
        }
        Carver cv(pass == 1 ? g_ws.ptr : nullptr);
        auto in_d = [&](const double* src, size_t count) -> const double* {
            if (src == nullptr) return nullptr;
            double* d = cv.take<double>(count);
            if (pass == 1) h2d.push_back({d, src, count * sizeof(double)});
            return d;
        };
        auto in_i = [&](const int32_t* src, size_t count) -> const int32_t* {
            if (src == nullptr) return nullptr;
            int32_t* d = cv.take<int32_t>(count);
            if (pass == 1) h2d.push_back({d, src, count * sizeof(int32_t)});
            return d;
        };
        auto out_d = [&](double* dst, size_t count) -> double* {
            if (dst == nullptr) return nullptr;
            double* d = cv.take<double>(count);
            if (pass == 1) d2h.push_back({dst, d, count * sizeof(double)});
            return d;
        };
        const size_t esz = out->profile_f32 ? sizeof(float) : sizeof(double);
        auto out_p = [&](double* dst, size_t count) -> double* {  // profile in the requested storage type
            if (dst == nullptr) return nullptr;
            double* d = reinterpret_cast<double*>(cv.take<char>(count * esz));
            if (pass == 1) d2h.push_back({dst, d, count * esz});
            return d;
        };
        din.psi = in_d(in->psi, S);
        din.K_b = in_d(in->K_b, S);
        din.G = in_d(in->G, S);
        din.mu_bar = in_d(in->mu_bar, S);
        din.G_int = in_d(in->G_int, 2 * S);
        din.tau_i = in_d(in->tau_i, S);
        din.tau_psi = in_d(in->tau_psi, S);
        din.lai_idx = in_i(in->lai_idx, S);
        din.leaf_idx = in_i(in->leaf_idx, S);
        din.soil_idx = in_i(in->soil_idx, S);
        din.sky_idx = in_i(in->sky_idx, S);
        din.lai_lib = in_d(in->lai_lib, (size_t)in->n_lai * nz);
        din.tau_d_lev = in_d(in->tau_d_lev, (size_t)in->n_lai * nz);
        din.leaf_r_lib = in_d(in->leaf_r_lib, (size_t)in->n_leaf * nw);
        din.leaf_t_lib = in_d(in->leaf_t_lib, (size_t)in->n_leaf * nw);
        din.soil_r_lib = in_d(in->soil_r_lib, (size_t)in->n_soil * nw);
        din.I_dr0_lib = in_d(in->I_dr0_lib, (size_t)in->n_sky * nw);
        din.I_df0_lib = in_d(in->I_df0_lib, (size_t)in->n_sky * nw);
        dout.band_w = in_d(out->band_w, (size_t)out->n_bw * nw);
        dout.I_dr = out_p(out->I_dr, prof);
        dout.I_df_d = out_p(out->I_df_d, prof);
        dout.I_df_u = out_p(out->I_df_u, prof);
        dout.F = out_p(out->F, prof);
        dout.x0 = out_p(out->x0, xprof);
        dout.x1 = out_p(out->x1, xprof);
        dout.x2 = out_p(out->x2, xprof);
        dout.rho_c = out_d(out->rho_c, S * nw);
        dout.absorbed = out_d(out->absorbed, S * (size_t)out->n_bw);
        need = cv.off;
    }

    cudaStream_t stream = nullptr;  // legacy default stream: ordered with the synchronous copies below
    for (const Copy& c : h2d) {
        e = cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    }
    e = crt::launch_solve(scheme, din, dout, can_vec2(&din, &dout), stream);
    if (e != cudaSuccess) return cuda_fail(e, (std::string("launch of ") + scheme_name(scheme) + " kernel").c_str());
    for (const Copy& c : d2h) {
        e = cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return cuda_fail(e, "D2H copy");
    }
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "kernel execution / synchronise");
    return CRT1D_OK;
}

}  // extern "C"
