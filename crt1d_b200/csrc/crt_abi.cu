// crt_abi.cu -- the extern "C" boundary declared in include/crt1d_b200.h: argument validation, error
// reporting, the device-pointer entry points, and the host-pointer path (H2D + kernels + D2H).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
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

// ---- host-pointer path: per-thread context (device workspace, streams, events, pinned staging ring) -----------
// NVTX ranges (header-only nvtx3; no-ops unless a profiler is attached) mark the phases of the host path.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Worker threads that move staged output from the page-locked ring into the caller's PAGEABLE arrays.  A D2H copy
// into pageable memory is bound by first-touch page faults of the destination (measured 4.7 GB/s for one thread,
// vs 50+ GB/s of PCIe 5 into page-locked memory), so the faults are spread over several threads while the DMA
// engine keeps filling the ring.  Process-wide, created on first use, never destroyed (no teardown order issues
// at exit); idle workers sleep on a condition variable.
class CopyPool {
public:
    struct Job {
        cudaEvent_t ready;      // recorded after the D2H copy into `src`
        void* dst;
        const void* src;
        size_t bytes;
        std::atomic<int>* busy; // staging slot flag, cleared when the memcpy is done
        std::atomic<int>* error;
        int device;
    };
    static CopyPool& get() {
        static CopyPool* pool = new CopyPool();
        return *pool;
    }
    int size() const { return (int)workers_.size(); }
    void submit(const Job& j) {
        {
            std::lock_guard<std::mutex> lock(mu_);
            q_.push_back(j);
        }
        cv_.notify_one();
    }
    // wait until `flag` is 0 (a staging slot is free again)
    void wait_clear(std::atomic<int>& flag) {
        std::unique_lock<std::mutex> lock(mu_);
        done_cv_.wait(lock, [&] { return flag.load(std::memory_order_acquire) == 0; });
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw >= 4 ? (int)(hw / 2) : 1;
        if (const char* e = getenv("CRT1D_B200_COPY_THREADS")) n = atoi(e);
        if (n < 1) n = 1;
        if (n > 16) n = 16;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
        for (auto& t : workers_) t.detach();
    }
    void run() {
        int dev = -1;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return !q_.empty(); });
                j = q_.front();
                q_.pop_front();
            }
            if (j.device != dev) {
                cudaSetDevice(j.device);
                dev = j.device;
            }
            if (cudaEventSynchronize(j.ready) != cudaSuccess) {
                cudaGetLastError();
                j.error->store(1);
            } else {
                memcpy(j.dst, j.src, j.bytes);
            }
            {
                std::lock_guard<std::mutex> lock(mu_);
                j.busy->store(0, std::memory_order_release);
            }
            done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Job> q_;
    std::vector<std::thread> workers_;
};

constexpr size_t kStagePiece = 8u << 20;     // bytes per staging slot
constexpr size_t kChunkTarget = 128u << 20;  // output bytes per device chunk slot (two slots)
constexpr size_t kStageMin = 4u << 20;       // pageable outputs below this go through plain cudaMemcpy

struct HostCtx {
    int device = -1;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    char* stage = nullptr;
    int n_stage = 0;
    std::vector<cudaEvent_t> ev_stage;
    std::unique_ptr<std::atomic<int>[]> stage_busy;
    std::atomic<int> error{0};

    void release() {
        if (device >= 0) cudaSetDevice(device);
        if (ws) cudaFree(ws);
        if (stage) cudaFreeHost(stage);
        for (cudaEvent_t e : ev_stage) cudaEventDestroy(e);
        for (int i = 0; i < 2; ++i) {
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
            if (ev_free[i]) cudaEventDestroy(ev_free[i]);
            ev_done[i] = ev_free[i] = nullptr;
        }
        if (s_compute) cudaStreamDestroy(s_compute);
        if (s_copy) cudaStreamDestroy(s_copy);
        ws = nullptr; ws_bytes = 0; stage = nullptr; n_stage = 0; ev_stage.clear(); stage_busy.reset();
        s_compute = s_copy = nullptr; device = -1;
        cudaGetLastError();
    }
    int prepare(int dev) {
        if (device != dev && device >= 0) release();
        if (s_compute == nullptr) {
            cudaError_t e = cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking);
            for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
                e = cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) return cuda_fail(e, "stream / event creation");
            device = dev;
        }
        return CRT1D_OK;
    }
    int reserve_ws(size_t bytes) {
        if (ws != nullptr && ws_bytes >= bytes) return CRT1D_OK;
        if (ws != nullptr) {
            cudaFree(ws);
            ws = nullptr;
            ws_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ws, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ws = nullptr;
            return fail(CRT1D_ERR_NO_MEMORY, "cudaMalloc of " + std::to_string(bytes) + " B workspace failed: " + cudaGetErrorString(e));
        }
        ws_bytes = bytes;
        return CRT1D_OK;
    }
    int reserve_stage(int slots) {
        if (n_stage >= slots) return CRT1D_OK;
        if (stage) {
            cudaFreeHost(stage);
            stage = nullptr;
        }
        for (cudaEvent_t e : ev_stage) cudaEventDestroy(e);
        ev_stage.clear();
        n_stage = 0;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&stage), (size_t)slots * kStagePiece, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            stage = nullptr;
            return fail(CRT1D_ERR_NO_MEMORY, std::string("cudaHostAlloc of the staging ring failed: ") + cudaGetErrorString(e));
        }
        ev_stage.resize(slots);
        for (int i = 0; i < slots; ++i) {
            e = cudaEventCreateWithFlags(&ev_stage[i], cudaEventDisableTiming | cudaEventBlockingSync);
            if (e != cudaSuccess) return cuda_fail(e, "event creation");
        }
        stage_busy.reset(new std::atomic<int>[slots]);
        for (int i = 0; i < slots; ++i) stage_busy[i].store(0);
        n_stage = slots;
        return CRT1D_OK;
    }
};
thread_local HostCtx g_ctx;

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

size_t pad256(size_t n) { return (n + 255) & ~size_t(255); }

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
        case CRT1D_NONFINITE: return "non-finite values in the results";
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
    g_ctx.release();
    return CRT1D_OK;
}

int64_t crt1d_preferred_batch(int scheme, int32_t n_z, int32_t n_wl, int64_t max_scen, int device) {
    if (scheme < 0 || scheme >= CRT1D_N_SCHEMES || n_z < 1 || n_wl < 1 || max_scen < 1) {
        fail(CRT1D_ERR_INVALID_ARG, "crt1d_preferred_batch: bad scheme id or size");
        return CRT1D_ERR_INVALID_ARG;
    }
    int n_sm = 0;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(CRT1D_ERR_NO_DEVICE, "crt1d_preferred_batch: no CUDA device");
    cudaError_t e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return cuda_fail(e, "crt1d_preferred_batch: cudaDeviceGetAttribute");
    return crt::preferred_batch(scheme, n_z, n_wl, max_scen, n_sm);
}

int crt1d_reload_tuning(void) {
    crt::reload_tuning();
    return CRT1D_OK;
}

// Host-pointer path: what a non-CUDA caller (the reference's Python, via ctypes) binds.
//
// The scenarios are processed in chunks through TWO device slots on two streams: the kernel of chunk k+1 runs
// while chunk k's results cross PCIe, so the call is bound by the D2H copy alone and the batch size is not limited
// by HBM (every profile of 10^5 scenarios = 400 GB passes through 2 x 128 MB).  Destinations that are page-locked
// receive the DMA directly; pageable destinations are filled from a page-locked staging ring by the CopyPool
// threads (their first-touch page faults, not PCIe, are the limit).
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
    NvtxRange r_all("crt1d_solve_host");
    HostCtx& cx = g_ctx;
    rc = cx.prepare(device);
    if (rc != CRT1D_OK) return rc;

    const size_t S = (size_t)in->n_scen, nz = (size_t)in->n_z, nw = (size_t)in->n_wl;
    const size_t xrows = (scheme == CRT1D_SCHEME_N79) ? nz - 1 : nz;
    const size_t esz = out->profile_f32 ? sizeof(float) : sizeof(double);

    // per-scenario output fields (host base, bytes per scenario)
    struct Field {
        char* host;
        size_t per_scen;
        double** dev_slot;  // which crt1d_out member receives the device pointer
        bool pinned;
        size_t off;         // offset inside a chunk slot
    };
    crt1d_out dout = *out;
    std::vector<Field> fields;
    auto add = [&](double* host, size_t per_scen, double** slot) {
        if (host != nullptr) fields.push_back({reinterpret_cast<char*>(host), per_scen, slot, false, 0});
    };
    add(out->I_dr, nz * nw * esz, &dout.I_dr);
    add(out->I_df_d, nz * nw * esz, &dout.I_df_d);
    add(out->I_df_u, nz * nw * esz, &dout.I_df_u);
    add(out->F, nz * nw * esz, &dout.F);
    add(out->x0, xrows * nw * esz, &dout.x0);
    add(out->x1, xrows * nw * esz, &dout.x1);
    add(out->x2, xrows * nw * esz, &dout.x2);
    add(out->rho_c, nw * sizeof(double), &dout.rho_c);
    add(out->absorbed, (size_t)out->n_bw * sizeof(double), &dout.absorbed);
    size_t per_scen_total = 0, pageable_total = 0;
    for (Field& f : fields) {
        f.pinned = is_pinned_host(f.host);
        per_scen_total += f.per_scen;
        if (!f.pinned) pageable_total += f.per_scen * S;
    }
    size_t C = per_scen_total ? kChunkTarget / per_scen_total : S;  // scenarios per chunk
    if (C < 1) C = 1;
    if (C > S) C = S;
    const size_t n_chunks = (S + C - 1) / C;
    size_t slot_bytes = 0;
    for (Field& f : fields) {
        f.off = slot_bytes;
        slot_bytes += pad256(f.per_scen * C);
    }
    const bool staged = pageable_total >= kStageMin;

    // ---- workspace: inputs | status | slot 0 | slot 1
    struct Copy {
        void* dst;
        const void* src;
        size_t bytes;
    };
    std::vector<Copy> h2d;
    crt1d_batch din = *in;
    size_t need = 0, slots_off = 0, status_off = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            rc = cx.reserve_ws(need);
            if (rc != CRT1D_OK) return rc;
        }
        char* base = pass == 1 ? static_cast<char*>(cx.ws) : nullptr;
        size_t off = 0;
        auto in_raw = [&](const void* src, size_t bytes) -> void* {
            if (src == nullptr) return nullptr;
            void* d = base + off;
            off += pad256(bytes);
            if (pass == 1) h2d.push_back({d, src, bytes});
            return d;
        };
        auto in_d = [&](const double* src, size_t count) { return static_cast<const double*>(in_raw(src, count * sizeof(double))); };
        auto in_i = [&](const int32_t* src, size_t count) { return static_cast<const int32_t*>(in_raw(src, count * sizeof(int32_t))); };
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
        status_off = off;
        off += pad256(S * sizeof(int32_t));
        slots_off = off;
        off += (n_chunks > 1 ? 2 : 1) * slot_bytes;
        need = off;
    }
    char* const wsb = static_cast<char*>(cx.ws);
    int32_t* const d_status = reinterpret_cast<int32_t*>(wsb + status_off);

    if (staged) {
        const size_t pieces = (pageable_total + kStagePiece - 1) / kStagePiece;
        const int want = (int)std::min<size_t>(pieces, 2 * (size_t)CopyPool::get().size() + 2);
        rc = cx.reserve_stage(want);
        if (rc != CRT1D_OK) return rc;
    }
    cx.error.store(0);

    {
        NvtxRange r("h2d inputs");
        for (const Copy& c : h2d) {
            e = cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyHostToDevice, cx.s_compute);
            if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
        }
    }

    // every exit below this point must drain the streams and the copy jobs first
    int stage_next = 0;
    auto drain = [&]() {
        cudaStreamSynchronize(cx.s_compute);
        cudaStreamSynchronize(cx.s_copy);
        for (int i = 0; i < cx.n_stage; ++i) CopyPool::get().wait_clear(cx.stage_busy[i]);
    };
    auto bail = [&](cudaError_t err, const char* what) {
        drain();
        return cuda_fail(err, what);
    };

    for (size_t k = 0; k < n_chunks; ++k) {
        const size_t s0 = k * C, cs = std::min(C, S - s0);
        const int slot = (int)(k & 1);
        char* const sb = wsb + slots_off + (size_t)slot * slot_bytes;
        // the chunk's view of the batch: per-scenario arrays advanced to s0, libraries shared
        crt1d_batch cb = din;
        cb.n_scen = (int64_t)cs;
        cb.psi = din.psi + s0;
        cb.K_b = din.K_b + s0;
        cb.G = din.G ? din.G + s0 : nullptr;
        cb.mu_bar = din.mu_bar ? din.mu_bar + s0 : nullptr;
        cb.G_int = din.G_int ? din.G_int + 2 * s0 : nullptr;
        cb.tau_i = din.tau_i ? din.tau_i + s0 : nullptr;
        cb.tau_psi = din.tau_psi ? din.tau_psi + s0 : nullptr;
        cb.lai_idx = din.lai_idx + s0;
        cb.leaf_idx = din.leaf_idx + s0;
        cb.soil_idx = din.soil_idx ? din.soil_idx + s0 : nullptr;
        cb.sky_idx = din.sky_idx + s0;
        crt1d_out co = dout;
        co.I_dr = co.I_df_d = co.I_df_u = co.F = co.x0 = co.x1 = co.x2 = co.rho_c = co.absorbed = nullptr;
        for (const Field& f : fields) {
            double** member = reinterpret_cast<double**>(reinterpret_cast<char*>(&co) + (reinterpret_cast<char*>(f.dev_slot) - reinterpret_cast<char*>(&dout)));
            *member = reinterpret_cast<double*>(sb + f.off);
        }
        co.status = d_status + s0;

        if (k >= 2) {
            e = cudaStreamWaitEvent(cx.s_compute, cx.ev_free[slot], 0);
            if (e != cudaSuccess) return bail(e, "cudaStreamWaitEvent");
        }
        {
            NvtxRange r("solve chunk");
            e = crt::launch_solve(scheme, cb, co, can_vec2(&cb, &co), cx.s_compute);
            if (e != cudaSuccess) return bail(e, (std::string("launch of ") + scheme_name(scheme) + " kernel").c_str());
        }
        e = cudaEventRecord(cx.ev_done[slot], cx.s_compute);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cx.s_copy, cx.ev_done[slot], 0);
        if (e != cudaSuccess) return bail(e, "event record / wait");

        NvtxRange r("d2h chunk");
        for (const Field& f : fields) {
            char* hdst = f.host + s0 * f.per_scen;
            const char* dsrc = sb + f.off;
            const size_t bytes = cs * f.per_scen;
            if (f.pinned || !staged) {
                e = cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, cx.s_copy);
                if (e != cudaSuccess) return bail(e, "D2H copy");
                continue;
            }
            for (size_t o = 0; o < bytes; o += kStagePiece) {
                const size_t n = std::min(kStagePiece, bytes - o);
                const int ss = stage_next;
                stage_next = (stage_next + 1) % cx.n_stage;
                CopyPool::get().wait_clear(cx.stage_busy[ss]);
                char* sbuf = cx.stage + (size_t)ss * kStagePiece;
                e = cudaMemcpyAsync(sbuf, dsrc + o, n, cudaMemcpyDeviceToHost, cx.s_copy);
                if (e == cudaSuccess) e = cudaEventRecord(cx.ev_stage[ss], cx.s_copy);
                if (e != cudaSuccess) return bail(e, "D2H copy (staged)");
                cx.stage_busy[ss].store(1, std::memory_order_release);
                CopyPool::get().submit({cx.ev_stage[ss], hdst + o, sbuf, n, &cx.stage_busy[ss], &cx.error, device});
            }
        }
        e = cudaEventRecord(cx.ev_free[slot], cx.s_copy);
        if (e != cudaSuccess) return bail(e, "event record");
    }

    // status words: one small copy after the last kernel
    std::vector<int32_t> h_status;
    int32_t* st = out->status;
    if (st == nullptr) {
        h_status.resize(S);
        st = h_status.data();
    }
    e = cudaMemcpyAsync(st, d_status, S * sizeof(int32_t), cudaMemcpyDeviceToHost, cx.s_compute);
    if (e != cudaSuccess) return bail(e, "D2H copy (status)");
    e = cudaStreamSynchronize(cx.s_compute);
    if (e != cudaSuccess) return bail(e, "kernel execution / synchronise");
    e = cudaStreamSynchronize(cx.s_copy);
    if (e != cudaSuccess) return bail(e, "D2H copies / synchronise");
    for (int i = 0; i < cx.n_stage; ++i) CopyPool::get().wait_clear(cx.stage_busy[i]);
    if (cx.error.load() != 0) return fail(CRT1D_ERR_CUDA, "a staged D2H copy failed");
    size_t n_bad = 0;
    for (size_t i = 0; i < S; ++i) n_bad += st[i] != 0;
    if (n_bad != 0) {
        g_last_error = std::string(scheme_name(scheme)) + ": non-finite values in " + std::to_string(n_bad) + " of " +
                       std::to_string(S) + " scenarios (see crt1d_out.status)";
        return CRT1D_NONFINITE;
    }
    return CRT1D_OK;
}

}  // extern "C"
