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
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <type_traits>

#include "crt_internal.h"
#include "crt_scheme.cuh"
#include "crt_spectra.cuh"

namespace crt {

constexpr int BLOCK = 128;  // threads per CTA of the elementwise kernels (absorption)

// ---------------------------------------------------------------------------------------------
// Tuning hooks (CRT1D_B200_* environment variables): kernel-selection overrides for experiments and for the
// tests that force a specific kernel.  Read ONCE, at the first launch (or again on crt1d_reload_tuning());
// the steady state of the ABI touches neither getenv nor sscanf.
// ---------------------------------------------------------------------------------------------
struct Tuning {
    int tile_threads = 0;        // CRT1D_B200_TILE_THREADS   (0 = per-scheme default)
    int diag_threads = 0;        // CRT1D_B200_DIAG_THREADS
    int rows_lv = 10, rows_th = 512, rows_rec = 1;  // CRT1D_B200_ROWS_CFG "LV,threads,rec"
    int split_4s = 4;            // CRT1D_B200_4S_SPLIT (measured: 2 -> 0.833, 3 -> 0.832, 4 -> 0.853 of HBM peak at 60 levels; n_z = 1000: 1 -> 0.795, 3 -> 0.776, 4 -> 0.815)
    int rows_threads = 0;        // CRT1D_B200_ROWS_THREADS
    long long scen_min = 148;    // CRT1D_B200_SCEN_MIN
    bool force_vec1 = false;     // CRT1D_B200_FORCE_VEC1
    bool tile_2s = false;        // CRT1D_B200_2S_KERNEL=tile
    bool no_rows = false;        // CRT1D_B200_NO_ROWS
    int fixup_parts = 0;         // CRT1D_B200_FIXUP_PARTS (0 = by level count)
    bool no_flat = false;        // CRT1D_B200_NO_FLAT: tridiagonal schemes keep the (scenario, band tile) mapping
    bool no_wide_ck = false;     // CRT1D_B200_NO_WIDE_CK: deep zq keeps the checkpoint spacing of 10 levels
    int smem_pad = 0;            // CRT1D_B200_SMEM_PAD: extra dynamic shared memory (bytes) for the tile / flat kernels (occupancy experiments)
};
static Tuning read_tuning() {
    Tuning t;
    if (const char* e = getenv("CRT1D_B200_TILE_THREADS")) t.tile_threads = atoi(e);
    if (const char* e = getenv("CRT1D_B200_DIAG_THREADS")) t.diag_threads = atoi(e);
    if (const char* e = getenv("CRT1D_B200_ROWS_CFG")) sscanf(e, "%d,%d,%d", &t.rows_lv, &t.rows_th, &t.rows_rec);
    if (const char* e = getenv("CRT1D_B200_4S_SPLIT")) t.split_4s = atoi(e);
    if (const char* e = getenv("CRT1D_B200_ROWS_THREADS")) t.rows_threads = atoi(e);
    if (const char* e = getenv("CRT1D_B200_SCEN_MIN")) t.scen_min = atoll(e);
    t.force_vec1 = getenv("CRT1D_B200_FORCE_VEC1") != nullptr;
    if (const char* e = getenv("CRT1D_B200_2S_KERNEL")) t.tile_2s = e[0] != 'r';
    t.no_rows = getenv("CRT1D_B200_NO_ROWS") != nullptr;
    t.no_flat = getenv("CRT1D_B200_NO_FLAT") != nullptr;
    t.no_wide_ck = getenv("CRT1D_B200_NO_WIDE_CK") != nullptr;
    if (const char* e = getenv("CRT1D_B200_SMEM_PAD")) t.smem_pad = atoi(e);
    if (const char* e = getenv("CRT1D_B200_FIXUP_PARTS")) t.fixup_parts = atoi(e);
    return t;
}
static Tuning g_tuning = read_tuning();
static const Tuning& tuning() { return g_tuning; }
void reload_tuning() { g_tuning = read_tuning(); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) and only when the request grows.
static cudaError_t ensure_dyn_smem(const void* kern, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    struct Entry { const void* k; int dev; size_t bytes; };
    static Entry cache[256];
    static int n_cache = 0;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    Entry* hit = nullptr;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kern && cache[i].dev == dev) { hit = &cache[i]; break; }
    if (hit && hit->bytes >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (hit) hit->bytes = bytes;
    else if (n_cache < 256) cache[n_cache++] = {kern, dev, bytes};
    return cudaSuccess;
}
template <class K>
static cudaError_t ensure_smem(K kern, size_t bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(kern), bytes); }
// one-CTA-per-scenario grids: blockIdx.x is 31 bits
static bool grid_ok(int64_t ctas) { return ctas > 0 && ctas <= 2147483647LL; }

// Checkpoint spacing of the zq / n79 Thomas sweeps (levels recomputed per segment in the back sweep).
#ifndef CRT_SEG_CK
#define CRT_SEG_CK 10
#endif
constexpr int SEG_CK = CRT_SEG_CK;
__host__ __device__ constexpr bool uses_segments(int scheme) {
    return scheme == CRT1D_SCHEME_ZQ || scheme == CRT1D_SCHEME_N79 || scheme == CRT1D_SCHEME_ZQ_PA;
}
// slots (levels) of the segment store: the segment itself, plus zq_pa's checkpoints (its M-grid is not the output grid)
__host__ __device__ inline int seg_slots(int scheme, int n_z, int ck = SEG_CK) {
    if (scheme != CRT1D_SCHEME_ZQ_PA) return ck + 1;  // + the prefetch slot of the next checkpoint (pf_tmp / ld_pf)
    const int M = n_z < ZQPA_MAX_M ? n_z : ZQPA_MAX_M;
    return SEG_CK + 1 + (M - 1) / SEG_CK + 1;  // segment slots 0..CK (slot 0 = the pair below the segment) + checkpoints
}
// doubles of level tables, rounded up so that the segment store behind them is 16-byte aligned
__host__ __device__ inline size_t tab_doubles(int scheme, int n_z) { return ((size_t)n_level_tables(scheme) * n_z + 1) & ~(size_t)1; }

// doubles of segment store per thread (thread-major, one pad element: see GlobalOut)
template <int VEC>
__host__ __device__ inline int seg_thread_doubles(int scheme, int n_z, int ck = SEG_CK) {
    return uses_segments(scheme) ? (seg_slots(scheme, n_z, ck) * 2 + 1) * VEC : 0;
}

// ---------------------------------------------------------------------------------------------
// global-memory column accessor
// ---------------------------------------------------------------------------------------------
// streaming store of VEC doubles to a typed cursor (float cursors round once)
template <int VEC>
__device__ __forceinline__ void st_raw(double* q, const double (&x)[VEC]) {
    if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<double2*>(q), make_double2(x[0], x[1]));
    } else {
        __stcs(q, x[0]);
    }
}
template <int VEC>
__device__ __forceinline__ void st_raw(float* q, const double (&x)[VEC]) {
    if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<float2*>(q), make_float2((float)x[0], (float)x[1]));
    } else {
        __stcs(q, (float)x[0]);
    }
}

// Fields are addressed as  base[f] + off + j * stride : `base` are the caller's field pointers (uniform:
// kernel parameters, they cost no per-thread registers), `off` the thread's element offset of
// (scenario, level 0, first band) -- `xoff` for the extra-output slots, whose row count can differ.
// FAST = every field of the scheme requested and float64 storage: no null checks, no dtype switch
// (15 -> 3 issue slots per store); the general path keeps both.
template <int VEC, bool FAST, int CK = SEG_CK>
struct GlobalOut {
    double* base[N_FIELDS];
    int64_t off, xoff, stride;
    bool f32;  // float32 storage (general path only): base[] really are float*

    __device__ __forceinline__ int64_t at(int f, int j) const { return (f >= F_X0 ? xoff : off) + (int64_t)j * stride; }

    // final results: written once, never re-read by this kernel -> streaming (evict-first) stores
    __device__ __forceinline__ void st(int f, int j, const double (&x)[VEC]) const {
        if constexpr (FAST) {
            st_raw<VEC>(base[f] + at(f, j), x);
        } else {
            if (base[f] == nullptr) return;
            if (f32) {
                st_raw<VEC>(reinterpret_cast<float*>(base[f]) + at(f, j), x);
            } else {
                st_raw<VEC>(base[f] + at(f, j), x);
            }
        }
    }
    // the four profile fields of one level: one row offset for all of them
    __device__ __forceinline__ void st4(int j, const double (&a)[VEC], const double (&b)[VEC], const double (&c)[VEC],
                                        const double (&d)[VEC]) const {
        if constexpr (FAST) {
            const int64_t o = off + (int64_t)j * stride;
            st_raw<VEC>(base[F_IDR] + o, a);
            st_raw<VEC>(base[F_DN] + o, b);
            st_raw<VEC>(base[F_UP] + o, c);
            st_raw<VEC>(base[F_F] + o, d);
        } else {
            st(F_IDR, j, a);
            st(F_DN, j, b);
            st(F_UP, j, c);
            st(F_F, j, d);
        }
    }
    // elimination scratch parked in the output arrays: re-read by the same thread during
    // back-substitution -> default (write-back, L2-resident) stores
    __device__ __forceinline__ void st_tmp(int f, int j, const double (&x)[VEC]) const {
        double* q = base[f] + at(f, j);
        if constexpr (VEC == 2) {
            *reinterpret_cast<double2*>(q) = make_double2(x[0], x[1]);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) q[v] = x[v];
        }
    }
    __device__ __forceinline__ void ld_tmp(int f, int j, double (&x)[VEC]) const {
        const double* q = base[f] + at(f, j);
        if constexpr (VEC == 2) {
            const double2 t = __ldcs(reinterpret_cast<const double2*>(q));  // last use: evict-first
            x[0] = t.x;
            x[1] = t.y;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[v] = __ldcs(q + v);
        }
    }
    // Checkpoint prefetch: the back sweep needs the parked pair of the NEXT segment ~1000 cycles after it starts the
    // current one, and an L2 round trip at that point is exposed latency in a latency-bound kernel (ncu: 21 % of the
    // stall cycles were long-scoreboard).  pf_tmp starts the copy into this thread's slot SEG_CK of the segment store
    // (cp.async: no registers held while it is in flight); ld_pf waits for it and reads the slot.
    __device__ __forceinline__ void pf_tmp(int f, int j, int k) const {
        const double* q = base[f] + at(f, j);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(seg + (CK * 2 + k) * VEC);
        if constexpr (VEC == 2) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(q) : "memory");
        } else {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(q) : "memory");
        }
    }
    __device__ __forceinline__ void ld_pf(int f, int j, int k, double (&x)[VEC]) const {
        asm volatile("cp.async.wait_all;" ::: "memory");
        seg_ld(CK, k, x);
    }
    // segment store of the checkpointed Thomas sweeps: `slots` levels x 2 values x VEC columns per thread in
    // shared memory, thread-major with one pad element: thread stride (2 slots + 1) * VEC doubles is an odd
    // multiple of the access width, so a warp's 8/16-byte accesses are bank-conflict free for any CTA size
    double* seg;  // this thread's first element
    __device__ __forceinline__ int seg_levels() const { return CK; }
    __device__ __forceinline__ void seg_st(int slot, int k, const double (&x)[VEC]) const {
        double* q = seg + (slot * 2 + k) * VEC;
        if constexpr (VEC == 2) {
            *reinterpret_cast<double2*>(q) = make_double2(x[0], x[1]);
        } else {
            q[0] = x[0];
        }
    }
    __device__ __forceinline__ void seg_ld(int slot, int k, double (&x)[VEC]) const {
        const double* q = seg + (slot * 2 + k) * VEC;
        if constexpr (VEC == 2) {
            const double2 t = *reinterpret_cast<const double2*>(q);
            x[0] = t.x;
            x[1] = t.y;
        } else {
            x[0] = q[0];
        }
    }
};

// streaming 8/16-byte load / store helpers
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
// profile store honouring crt1d_out.profile_f32: `base` points at the scenario's field in its own dtype
template <int VEC>
__device__ __forceinline__ void st_prof(void* base, bool f32, int64_t off, const double (&x)[VEC]) {
    if (base == nullptr) return;
    if (f32) {
        float* q = static_cast<float*>(base) + off;
        if constexpr (VEC == 2) {
            __stcs(reinterpret_cast<float2*>(q), make_float2((float)x[0], (float)x[1]));
        } else {
            __stcs(q, (float)x[0]);
        }
    } else {
        double* q = static_cast<double*>(base) + off;
        if constexpr (VEC == 2) {
            __stcs(reinterpret_cast<double2*>(q), make_double2(x[0], x[1]));
        } else {
            __stcs(q, x[0]);
        }
    }
}
template <int VEC, bool F32>
__device__ __forceinline__ void st_prof_t(void* base, int64_t off, const double (&x)[VEC]) {
    if (base == nullptr) return;
    if constexpr (F32) {
        float* q = static_cast<float*>(base) + off;
        if constexpr (VEC == 2) {
            __stcs(reinterpret_cast<float2*>(q), make_float2((float)x[0], (float)x[1]));
        } else {
            __stcs(q, (float)x[0]);
        }
    } else {
        double* q = static_cast<double*>(base) + off;
        if constexpr (VEC == 2) {
            __stcs(reinterpret_cast<double2*>(q), make_double2(x[0], x[1]));
        } else {
            __stcs(q, x[0]);
        }
    }
}
__device__ __forceinline__ void* prof_base(double* ptr, bool f32, int64_t elem_off) {
    if (ptr == nullptr) return nullptr;
    return f32 ? static_cast<void*>(reinterpret_cast<float*>(ptr) + elem_off) : static_cast<void*>(ptr + elem_off);
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

// ---------------------------------------------------------------------------------------------
// the solver kernel
// ---------------------------------------------------------------------------------------------
template <int SCHEME, int VEC, int BLK, int MINB, bool FAST>
__global__ void __launch_bounds__(BLK, MINB) solve_kernel(const crt1d_batch in, const crt1d_out out, int tiles_per_scen,
                                                          int tiles_per_cta) {
    extern __shared__ double tab[];
    __shared__ double red[BLK / 32][4];

    const int nthr = blockDim.x;  // <= BLK: the launcher picks the CTA size that leaves the fewest idle lanes in the last tile
    const int ctas_per_scen = (tiles_per_scen + tiles_per_cta - 1) / tiles_per_cta;
    const int64_t s = blockIdx.x / ctas_per_scen;
    const int t0 = (blockIdx.x % ctas_per_scen) * tiles_per_cta;
    const int t1 = min(t0 + tiles_per_cta, tiles_per_scen);
    const int n_z = in.n_z, n_wl = in.n_wl;

    for (int j = threadIdx.x; j < n_z; j += nthr) fill_level_tables<SCHEME>(in, s, j, tab);
    __syncthreads();
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        for (int j = threadIdx.x; j < n_z; j += nthr) fill_level_tables_2<SCHEME>(in, s, j, tab);
        __syncthreads();
        for (int j = threadIdx.x; j < n_z; j += nthr) fill_level_tables_3<SCHEME>(in, s, j, tab);
        __syncthreads();
    }

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t prof = (int64_t)n_z * n_wl;                       // doubles per scenario in a profile
    const int64_t xprof = (int64_t)extra_rows(SCHEME, n_z) * n_wl;  // ... in an extra-output slot

    for (int t = t0; t < t1; ++t) {
        const int b0 = (t * nthr + threadIdx.x) * VEC;
        if (b0 >= n_wl) continue;
        const BandIn<VEC> b = load_bands<VEC>(in, s, b0);
        GlobalOut<VEC, FAST> o;
        o.stride = n_wl;
        o.f32 = out.profile_f32 != 0;
        o.seg = tab + tab_doubles(SCHEME, n_z) + (size_t)threadIdx.x * seg_thread_doubles<VEC>(SCHEME, n_z);
        o.off = s * prof + b0;
        o.xoff = s * xprof + b0;
        o.base[F_IDR] = out.I_dr;
        o.base[F_DN] = out.I_df_d;
        o.base[F_UP] = out.I_df_u;
        o.base[F_F] = out.F;
        o.base[F_X0] = out.x0;
        o.base[F_X1] = out.x1;
        o.base[F_X2] = out.x2;
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
        if (out.status) {  // `ab` combines all three irradiances of the ground and top levels
            bool bad = false;
#pragma unroll
            for (int v = 0; v < VEC; ++v) bad = bad || !isfinite(ab[v]);
            if (bad) atomicOr(out.status + s, CRT1D_STATUS_NONFINITE);
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
            for (int w = 0; w < nthr / 32; ++w) v += red[w][threadIdx.x];
            out.absorbed[s * out.n_bw + threadIdx.x] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Flat column mapping for the tridiagonal schemes (zq, n79, zq_pa): the batch's column groups (VEC adjacent bands of
// one scenario) are numbered through, scenario after scenario, and CTA c takes groups [c nthr, (c+1) nthr) -- so
// every CTA is full.  With the (scenario, band tile) mapping of solve_kernel the last tile of a scenario is mostly
// empty (2100 bands / 512 columns per tile: 26 of 256 threads in the fifth tile), and these kernels are bound by
// the latency of their per-column recurrences at 16 warps per SM: the early-exiting warps are latency-hiding
// capacity lost for a whole CTA lifetime (18 % of the warp slots at 2100 bands).  A CTA spans at most two scenarios
// (launcher: nthr <= groups per scenario) and keeps both scenarios' level tables.
// Canopy-absorbed sums: each CTA reduces its columns per spanned scenario in a fixed order into
// partial[c][which][band group]; absorbed_flat_finish_kernel adds a scenario's partials in CTA order (deterministic).
// ---------------------------------------------------------------------------------------------
template <int SCHEME, int VEC, int BLK, int MINB, bool FAST, int CK = SEG_CK>
__global__ void __launch_bounds__(BLK, MINB) solve_flat_kernel(const crt1d_batch in, const crt1d_out out, int gps,
                                                               double* __restrict__ partial) {
    extern __shared__ double tab[];
    __shared__ double red[BLK / 32][8];

    const int nthr = blockDim.x;
    const int n_z = in.n_z, n_wl = in.n_wl;
    const int64_t g0 = (int64_t)blockIdx.x * nthr;
    const int64_t total = in.n_scen * gps;
    const int64_t g1 = min(total, g0 + nthr);  // this CTA's groups: [g0, g1)
    const int64_t s_first = g0 / gps, s_last = (g1 - 1) / gps;
    const size_t td = tab_doubles(SCHEME, n_z);

    for (int j = threadIdx.x; j < n_z; j += nthr) {
        fill_level_tables<SCHEME>(in, s_first, j, tab);
        if (s_last != s_first) fill_level_tables<SCHEME>(in, s_last, j, tab + td);
    }
    __syncthreads();
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        for (int j = threadIdx.x; j < n_z; j += nthr) {
            fill_level_tables_2<SCHEME>(in, s_first, j, tab);
            if (s_last != s_first) fill_level_tables_2<SCHEME>(in, s_last, j, tab + td);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < n_z; j += nthr) {
            fill_level_tables_3<SCHEME>(in, s_first, j, tab);
            if (s_last != s_first) fill_level_tables_3<SCHEME>(in, s_last, j, tab + td);
        }
        __syncthreads();
    }

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int which = 0;
    const int64_t gid = g0 + threadIdx.x;
    if (gid < g1) {
        const int64_t s = gid / gps;
        const int b0 = (int)(gid - s * gps) * VEC;
        which = s != s_first;
        const int64_t prof = (int64_t)n_z * n_wl;                       // doubles per scenario in a profile
        const int64_t xprof = (int64_t)extra_rows(SCHEME, n_z) * n_wl;  // ... in an extra-output slot
        const BandIn<VEC> b = load_bands<VEC>(in, s, b0);
        GlobalOut<VEC, FAST, CK> o;
        o.stride = n_wl;
        o.f32 = out.profile_f32 != 0;
        o.seg = tab + 2 * td + (size_t)threadIdx.x * seg_thread_doubles<VEC>(SCHEME, n_z, CK);
        o.off = s * prof + b0;
        o.xoff = s * xprof + b0;
        o.base[F_IDR] = out.I_dr;
        o.base[F_DN] = out.I_df_d;
        o.base[F_UP] = out.I_df_u;
        o.base[F_F] = out.F;
        o.base[F_X0] = out.x0;
        o.base[F_X1] = out.x1;
        o.base[F_X2] = out.x2;
        double rho_c[VEC], ab[VEC];
        solve_column_group<SCHEME, VEC>(in, s, tab + which * td, b, o, rho_c, ab);
        if (partial) {
            for (int k = 0; k < out.n_bw; ++k) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[k] += out.band_w[(int64_t)k * n_wl + b0 + v] * ab[v];
            }
        }
        if (out.status) {
            bool bad = false;
#pragma unroll
            for (int v = 0; v < VEC; ++v) bad = bad || !isfinite(ab[v]);
            if (bad) atomicOr(out.status + s, CRT1D_STATUS_NONFINITE);
        }
    }

    if (partial) {  // fixed-order reduction: lanes by shuffle tree, warps in warp order
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double v = (q >> 2) == which ? acc[q & 3] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) red[warp][q] = v;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            double v = 0.0;
            for (int w = 0; w < nthr / 32; ++w) v += red[w][threadIdx.x];
            partial[(int64_t)blockIdx.x * 8 + threadIdx.x] = v;
        }
    }
}

// absorbed[s][k] = sum over the CTAs of solve_flat_kernel that hold columns of scenario s, in CTA order.
__global__ void __launch_bounds__(256) absorbed_flat_finish_kernel(const double* __restrict__ partial, int64_t n_scen, int gps,
                                                                   int nthr, int n_bw, double* __restrict__ absorbed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scen * n_bw) return;
    const int64_t s = i / n_bw;
    const int k = (int)(i - s * n_bw);
    const int64_t c_lo = (s * gps) / nthr, c_hi = ((s + 1) * gps - 1) / nthr;
    double v = 0.0;
    for (int64_t c = c_lo; c <= c_hi; ++c) {
        const int which = (c * nthr) / gps != s;  // slot 0 = the CTA's first scenario
        v += partial[c * 8 + which * 4 + k];
    }
    absorbed[i] = v;
}

// Number of output fields a scheme writes (4 profiles + its extra slots).
__host__ __device__ constexpr int n_out_fields(int scheme) {
    return scheme == CRT1D_SCHEME_ZQ ? 7 : scheme == CRT1D_SCHEME_N79 ? 6 : (scheme == CRT1D_SCHEME_BF || scheme == CRT1D_SCHEME_G77) ? 7 : 4;
}
// The compile-time fast store path exists for the schemes whose only kernel this is.
__host__ __device__ constexpr bool has_fast_tile(int scheme) {
    return scheme == CRT1D_SCHEME_ZQ || scheme == CRT1D_SCHEME_N79 || scheme == CRT1D_SCHEME_ZQ_PA;
}

#ifndef CRT_ZQPA_VEC
#define CRT_ZQPA_VEC 1
#endif
#ifndef CRT_ZQPA_THREADS
#define CRT_ZQPA_THREADS 256
#endif
// zq_pa: one column per thread.  CTA size: 256 threads when its shared memory fits (closed M-grid solution: 0.66 of HBM
// peak vs 0.59 with 128-thread CTAs -- longer row fragments per store; two columns per thread 0.51 / 0.48), else 128.
constexpr int ZQPA_VEC = CRT_ZQPA_VEC, ZQPA_THREADS = CRT_ZQPA_THREADS, ZQPA_THREADS_MIN = 128;

// Threads per CTA of a tile-kernel launch: the configured size, or CRT1D_B200_TILE_THREADS (<= the compiled
// bound; tuning).  One CTA walks a scenario's bands in passes of nthr * VEC columns and the last pass is partly
// empty (2100 bands, 256 threads x 2: 18 % of the thread-passes idle) -- but picking the size with the fullest
// last pass does NOT pay: measured zq 96 thr 0.68 | 128 0.70 | 192 0.74 | 256 0.74; n79 0.60-0.62 for all;
// zq_pa 96 0.47 | 128 0.50 | 192 0.45 | 256 0.50 (early-exiting warps free their issue slots; longer row
// fragments help the stores).
static int tile_threads(int cfg, int blk) {
    const int t = tuning().tile_threads;
    if (t >= 32 && t <= blk && t % 32 == 0) return t;
    return cfg;
}

// Stream-ordered scratch for the flat kernels' per-CTA absorbed partials (pool memory is kept across calls).
static cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t stream) {
    static std::once_flag once[64];
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) {
        std::call_once(once[dev], [dev] {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        });
    }
    return cudaMallocAsync(p, bytes, stream);
}

// Flat mapping (solve_flat_kernel) when a CTA cannot span more than two scenarios and two sets of level tables leave
// the shared memory for the same number of resident CTAs.
template <int SCHEME, int VEC, int BLK, int MINB>
static bool flat_eligible(const crt1d_batch& in, int nthr, size_t& smem, int ck = SEG_CK) {
    if constexpr (!uses_segments(SCHEME)) return false;
    // zq_pa keeps the (scenario, band tile) mapping: its three-pass scenario prologue (M-grid, interpolation tables,
    // emission order) is paid per CTA, and the flat mapping has 16 CTAs per scenario where the fused-absorbed tile
    // mapping has one: measured 0.490 (flat) vs 0.515 (tile) of HBM peak; zq 0.755 -> 0.860, n79 0.743 -> 0.797.
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) return false;
    if (tuning().no_flat) return false;
    const int gps = in.n_wl / VEC;
    if (gps * VEC != in.n_wl || nthr > gps) return false;
    smem = (2 * tab_doubles(SCHEME, in.n_z) + (size_t)nthr * seg_thread_doubles<VEC>(SCHEME, in.n_z, ck)) * sizeof(double);
    if (ck != SEG_CK) return smem <= 227u * 1024u - 2048u;  // the wide-spacing variant runs one CTA per SM by design
    const size_t old_smem = (tab_doubles(SCHEME, in.n_z) + (size_t)nthr * seg_thread_doubles<VEC>(SCHEME, in.n_z)) * sizeof(double);
    const size_t per_sm = 228u * 1024u, per_cta = 1024u + sizeof(double) * (BLK / 32) * 8;
    // resident CTAs the (scenario, band tile) mapping gets: the register bound (MINB CTAs of BLK threads) or its shared memory
    // (n79 at n_z = 1000 keeps ONE resident CTA either way: 64 B of tables per level.  Moving the tables to global memory
    // (read through L1) so that two CTAs fit was measured and dropped: 0.61 -> 0.53-0.57 of HBM peak -- twice the resident
    // columns double the parked-checkpoint working set (242 MB at two CTAs per SM against 126 MB of L2).)
    const size_t want = std::min<size_t>((size_t)MINB * (BLK / nthr), per_sm / (old_smem + per_cta));
    return smem <= 227u * 1024u && per_sm / (smem + per_cta) >= want;
}

template <int SCHEME, int VEC, int BLK, int MINB, int CK = SEG_CK>
static cudaError_t launch_flat(const crt1d_batch& in, const crt1d_out& out, int nthr, size_t smem, cudaStream_t stream) {
    const int gps = in.n_wl / VEC;
    const int64_t grid = (in.n_scen * gps + nthr - 1) / nthr;
    if (grid <= 0 || grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    auto kern = solve_flat_kernel<SCHEME, VEC, BLK, MINB, false, CK>;
    double* const f[7] = {out.I_dr, out.I_df_d, out.I_df_u, out.F, out.x0, out.x1, out.x2};
    bool all = out.profile_f32 == 0;
    for (int q = 0; q < n_out_fields(SCHEME); ++q) all = all && f[q] != nullptr;
    if (all) kern = solve_flat_kernel<SCHEME, VEC, BLK, MINB, true, CK>;
    smem = std::min<size_t>(smem + (size_t)std::max(0, tuning().smem_pad), 227u * 1024u);
    if (cudaError_t e = ensure_smem(kern, smem); e != cudaSuccess) return e;
    double* partial = nullptr;
    if (out.absorbed) {
        if (cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&partial), (size_t)grid * 8 * sizeof(double), stream); e != cudaSuccess) return e;
    }
    kern<<<(unsigned)grid, nthr, smem, stream>>>(in, out, gps, partial);
    cudaError_t e = cudaGetLastError();
    if (partial) {
        if (e == cudaSuccess) {
            const int64_t n = in.n_scen * out.n_bw;
            absorbed_flat_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(partial, in.n_scen, gps, nthr, out.n_bw, out.absorbed);
            e = cudaGetLastError();
        }
        cudaFreeAsync(partial, stream);
    }
    return e;
}

// Deep canopies, zq: the checkpoints parked by all resident columns -- n_SM x 512 threads x (n_z / CK) x 32 B at two CTAs per SM --
// outgrow the 126 MB L2 near n_z = 520 at CK = 10, and every parked pair then makes a round trip to HBM.  From there on zq runs
// the flat kernel with checkpoints every 20 levels (twice the segment store: ONE resident CTA, a quarter of the parked bytes in
// flight).  Measured at n_z = 1000 (variant builds, one box): CK = 10 0.665 | 14 0.693 | 16 0.703 | 20 0.707 of HBM peak.
// Same arithmetic in the same order: the profiles are bit-identical for any spacing.
#ifndef CRT_ZQ_DEEP_CK
#define CRT_ZQ_DEEP_CK 20
#endif
#ifndef CRT_ZQ_DEEP_NZ
#define CRT_ZQ_DEEP_NZ 512
#endif
constexpr int ZQ_DEEP_CK = CRT_ZQ_DEEP_CK;
static bool zq_wide_spacing(int n_z) { return ZQ_DEEP_CK != SEG_CK && n_z >= CRT_ZQ_DEEP_NZ && !tuning().no_flat && !tuning().no_wide_ck; }

template <int SCHEME, int VEC, int BLK, int MINB>
static cudaError_t launch_one(const crt1d_batch& in, const crt1d_out& out, cudaStream_t stream) {
    int nthr = tile_threads(BLK, BLK);
    if constexpr (SCHEME == CRT1D_SCHEME_ZQ_PA) {
        const size_t want = (tab_doubles(SCHEME, in.n_z) + (size_t)ZQPA_THREADS * seg_thread_doubles<VEC>(SCHEME, in.n_z)) * sizeof(double);
        nthr = tile_threads(want + 2048u <= 227u * 1024u ? ZQPA_THREADS : ZQPA_THREADS_MIN, BLK);
    }
    if constexpr (uses_segments(SCHEME)) {
        size_t smem_flat = 0;
        if constexpr (SCHEME == CRT1D_SCHEME_ZQ) {
            if (zq_wide_spacing(in.n_z) && flat_eligible<SCHEME, VEC, BLK, MINB>(in, nthr, smem_flat, ZQ_DEEP_CK))
                return launch_flat<SCHEME, VEC, BLK, MINB, ZQ_DEEP_CK>(in, out, nthr, smem_flat, stream);
        }
        if (flat_eligible<SCHEME, VEC, BLK, MINB>(in, nthr, smem_flat)) return launch_flat<SCHEME, VEC, BLK, MINB>(in, out, nthr, smem_flat, stream);
    }
    const int cols = nthr * VEC;
    const int tiles_per_scen = (in.n_wl + cols - 1) / cols;
    const int tiles_per_cta = out.absorbed ? tiles_per_scen : 1;
    const int ctas_per_scen = (tiles_per_scen + tiles_per_cta - 1) / tiles_per_cta;
    const int64_t grid = in.n_scen * ctas_per_scen;
    if (grid <= 0 || grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    size_t smem = (size_t)n_level_tables(SCHEME) * in.n_z * sizeof(double);
    if (uses_segments(SCHEME)) smem = (tab_doubles(SCHEME, in.n_z) + (size_t)nthr * seg_thread_doubles<VEC>(SCHEME, in.n_z)) * sizeof(double);
    auto kern = solve_kernel<SCHEME, VEC, BLK, MINB, false>;
    if constexpr (has_fast_tile(SCHEME)) {
        double* const f[7] = {out.I_dr, out.I_df_d, out.I_df_u, out.F, out.x0, out.x1, out.x2};
        bool all = out.profile_f32 == 0;
        for (int q = 0; q < n_out_fields(SCHEME); ++q) all = all && f[q] != nullptr;
        if (all) kern = solve_kernel<SCHEME, VEC, BLK, MINB, true>;
    }
    smem = std::min<size_t>(smem + (size_t)std::max(0, tuning().smem_pad), 227u * 1024u);
    if (cudaError_t e = ensure_smem(kern, smem); e != cudaSuccess) return e;
    kern<<<(unsigned)grid, nthr, smem, stream>>>(in, out, tiles_per_scen, tiles_per_cta);
    return cudaGetLastError();
}

// Tile-kernel launch configuration per scheme: (threads per CTA, resident CTAs per SM the register
// allocator must allow), the best of the measured sweep in profiles/r01_tile_kernel_config_all_schemes.txt
// (128,1 | 128,4 | 256,2 tried for every scheme).  Only the chosen one is compiled.
template <int SCHEME>
struct TileCfg {
    static constexpr int BLK = 128, MINB = 1;  // 4s (wants its 212 registers)
};
template <> struct TileCfg<CRT1D_SCHEME_2S> { static constexpr int BLK = 256, MINB = 2; };   // 4 KB row fragments: 0.79 -> 0.85
template <> struct TileCfg<CRT1D_SCHEME_BL> { static constexpr int BLK = 256, MINB = 2; };   // 0.80 -> 0.87
template <> struct TileCfg<CRT1D_SCHEME_BF> { static constexpr int BLK = 256, MINB = 2; };   // +3.5 %
template <> struct TileCfg<CRT1D_SCHEME_G77> { static constexpr int BLK = 256, MINB = 2; };  // +3.4 %
#ifndef CRT_TRI_BLK  // tuning hooks for the tridiagonal schemes (build.py `defines`)
#define CRT_TRI_BLK 128
#define CRT_TRI_MINB 4
#endif
// Tridiagonal schemes (checkpointed sweeps; measured on 66 304-scenario runs, fraction of HBM peak):
//   zq    (128,4) 0.706 | (128,3) 0.729 | (256,2) 0.748 | (128,5) 0.627 | one column per thread 0.645
//   n79   (128,4) 0.632 | (128,3) 0.609 | (256,2) 0.625 | (128,5) 0.555 | one column per thread 0.498
//   zq_pa (128,4) one column per thread 0.476 | (128,3) two columns 0.461 | (128,4) two columns 0.350 | (128,2) 0.383
#ifndef CRT_ZQPA_MINB
#define CRT_ZQPA_MINB 2
#endif
template <> struct TileCfg<CRT1D_SCHEME_ZQ_PA> { static constexpr int BLK = 256, MINB = CRT_ZQPA_MINB; };
#ifndef CRT_ZQ_MINB
#define CRT_ZQ_MINB 2
#endif
#ifndef CRT_N79_MINB
#define CRT_N79_MINB 2
#endif
template <> struct TileCfg<CRT1D_SCHEME_ZQ> { static constexpr int BLK = 256, MINB = CRT_ZQ_MINB; };
template <> struct TileCfg<CRT1D_SCHEME_N79> { static constexpr int BLK = 256, MINB = CRT_N79_MINB; };

template <int SCHEME>
static cudaError_t launch_vec(const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    constexpr int B = TileCfg<SCHEME>::BLK, M = TileCfg<SCHEME>::MINB;
    return vec2 ? launch_one<SCHEME, 2, B, M>(in, out, stream) : launch_one<SCHEME, 1, B, M>(in, out, stream);
}

size_t solve_shared_bytes(int scheme, int n_z) {
    if (!uses_segments(scheme)) return (size_t)n_level_tables(scheme) * n_z * sizeof(double);
    if (scheme == CRT1D_SCHEME_ZQ_PA)
        return (tab_doubles(scheme, n_z) + (size_t)ZQPA_THREADS_MIN * seg_thread_doubles<ZQPA_VEC>(scheme, n_z)) * sizeof(double);
    const int blk = scheme == CRT1D_SCHEME_ZQ ? TileCfg<CRT1D_SCHEME_ZQ>::BLK : TileCfg<CRT1D_SCHEME_N79>::BLK;
    return (tab_doubles(scheme, n_z) + (size_t)blk * seg_thread_doubles<2>(scheme, n_z)) * sizeof(double);
}

// ---------------------------------------------------------------------------------------------
// 2s row-sweep kernel: ONE CTA = one whole scenario, per-band coefficients in SHARED MEMORY.
//
// The level sweep of a closed-form scheme needs only the band's folded coefficients (2s: 8 doubles).
// Holding them in shared memory (8 x n_wl x 8 B = 134 KB at 2100 bands) instead of registers frees the
// mapping of lanes to columns: warps pull work items (group of LV consecutive levels x 32*VEC adjacent
// bands) from a shared counter in row-major order, so the CTA writes complete 16.8 KB band rows of
// every field, level after level, with perfect load balance and ~64 registers per thread (1024
// threads, all 64 warp slots of the SM busy).  One resident CTA per SM => 148 row streams in flight,
// the pattern that reaches ~6.8 TB/s on B200 (tools/micro/wbw.cu) where the band-tile pattern
// saturates at 5.1 TB/s.  The canopy-absorbed reduction uses the ground/top levels evaluated in the
// coefficient phase (fixed thread->column assignment), so it stays deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int ROWS_MAX_CHUNKS = 256;  // chunks of 32 lanes per band row the row-sweep kernels support

// Warp-level pieces of the row-sweep kernels' work-item protocol.
// An item is (level group lg, chunk of 32*VEC bands).  The warp that gets a chunk's FIRST item (lg = 0) computes
// that chunk's per-band coefficients, publishes them in shared memory and raises ready[chunk]; items of later
// level groups wait for the flag (their producer was handed out earlier, so it is running or done: no deadlock).
// This removes the separate coefficient phase during which an SM issued no stores (12.6 % of a CTA's life).
__device__ __forceinline__ void rows_publish(int* ready, int chunk, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int*>(ready + chunk) = 1;
}
__device__ __forceinline__ void rows_wait(int* ready, int chunk, int lane) {
    if (lane == 0) {
        while (*reinterpret_cast<volatile int*>(ready + chunk) == 0) __nanosleep(40);
    }
    __syncwarp();
    __threadfence_block();
}

// DIAG = reduced-diagnostic instantiation (no profile requested): only the first item of every chunk exists, the
// coefficients are never stored (shared memory = level tables only) and chunks are dealt round-robin.  Small CTAs,
// several per SM: a CTA's serial parts (tables, barriers, final chunk sum) overlap the other CTAs' arithmetic --
// as one 512-thread CTA per SM this mode was latency-bound at ~13 us per scenario.
template <int VEC, int LV, int MAXT, bool REC, bool F32, int MINB = 1, bool DIAG = false>
__global__ void __launch_bounds__(MAXT, MINB) solve_2s_rows_kernel(const crt1d_batch in, const crt1d_out out) {
    extern __shared__ double sm[];
    __shared__ double partial[ROWS_MAX_CHUNKS][4];  // per-chunk sums of the absorbed reduction (fixed final order)
    __shared__ int ready[ROWS_MAX_CHUNKS];
    __shared__ int counter;
    __shared__ unsigned char grp_uniform[1024];  // per level group: 1 if its levels are equally spaced (n_z <= 1024*LV)

    const int64_t s = blockIdx.x;
    const int n_z = in.n_z, n_wl = in.n_wl, T = blockDim.x;
    const int ld = (n_wl + 1) & ~1;  // coefficient row stride (even => 16-byte aligned pairs)
    double* L = sm;
    double* eK = sm + n_z;
    double* cf = sm + 2 * n_z + ((2 * n_z) & 1);  // [8][ld], 16-byte aligned
    const int n_grp = n_wl / VEC;
    const int n_chunks = (n_grp + 31) / 32;

    // ---- phase A: level tables, flags
    for (int j = threadIdx.x; j < n_z; j += T) fill_level_tables<CRT1D_SCHEME_2S>(in, s, j, sm);
    for (int q = threadIdx.x; q < n_chunks; q += T) ready[q] = 0;
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    if (REC && !DIAG) {
        const double tol = 8.0 * 2.220446049250313e-16 * L[0];
        for (int g = threadIdx.x; g < (n_z + LV - 1) / LV && g < 1024; g += T) {
            const int a = g * LV, b = min(n_z, a + LV);
            bool u = (b - a) >= 2;
            for (int j = a + 1; j + 1 < b; ++j) u = u && fabs((L[j] - L[j + 1]) - (L[a] - L[a + 1])) <= tol;
            grp_uniform[g] = u ? 1 : 0;
        }
        __syncthreads();
    }

    const Scen2s sc = scen_2s(in.psi[s], in.K_b[s], in.mu_bar[s], in.mla_deg, L[0]);
    const int64_t prof = (int64_t)n_z * n_wl;
    void* pI = prof_base(out.I_dr, F32, s * prof);
    void* pD = prof_base(out.I_df_d, F32, s * prof);
    void* pU = prof_base(out.I_df_u, F32, s * prof);
    void* pF = prof_base(out.F, F32, s * prof);
    // Reduced-diagnostic mode: with no profile requested only the first item of every chunk runs (coefficients +
    // ground/top levels for the absorbed reduction); there is nothing to sweep.
    const bool any_profile = !DIAG && (pI || pD || pU || pF);
    const bool all_profiles = pI && pD && pU && pF;
    const int n_lg = any_profile ? (n_z + LV - 1) / LV : 1;
    const int n_items = n_chunks * n_lg;
    const int lane = threadIdx.x & 31;

    // ---- row-major work items
    int item_rr = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if constexpr (DIAG) {
            item = item_rr;
            item_rr += T >> 5;
        } else {
            if (lane == 0) item = atomicAdd(&counter, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
        }
        if (item >= n_items) break;
        const int lg = item / n_chunks, chunk = item - lg * n_chunks;
        const int g = chunk * 32 + lane;
        const bool valid = g < n_grp;
        const int c0 = (valid ? g : 0) * VEC;
        Coef2s k[VEC];
        if (lg == 0) {  // first touch of this chunk: produce its coefficients (ref _solve_2s.py:65-120)
            double ab[VEC];
            if (valid) {
                const BandIn<VEC> b = load_bands<VEC>(in, s, c0);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    k[v] = coef_2s<DIAG>(sc, b.leaf_r[v], b.leaf_t[v], b.soil_r[v], b.Idr0[v], b.Idf0[v]);
                    if constexpr (DIAG) continue;
                    cf[0 * ld + c0 + v] = k[v].h;
                    cf[1 * ld + c0 + v] = k[v].Au;
                    cf[2 * ld + c0 + v] = k[v].Bu;
                    cf[3 * ld + c0 + v] = k[v].Cu;
                    cf[4 * ld + c0 + v] = k[v].Ad;
                    cf[5 * ld + c0 + v] = k[v].Bd;
                    cf[6 * ld + c0 + v] = k[v].Cd;
                    cf[7 * ld + c0 + v] = k[v].Idr0;
                }
            }
            if (out.absorbed || out.status) {  // ground/top levels -> canopy-absorbed sums of this chunk, fixed lane order
                double a4[4] = {0.0, 0.0, 0.0, 0.0};
                if (valid) {
                    bool bad = false;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        double Ig, dg, ug, Fg, It, dt, ut, Ft;
                        level_2s(k[v], sc.inv_mu, L[0], eK[0], Ig, dg, ug, Fg);
                        level_2s(k[v], sc.inv_mu, L[n_z - 1], eK[n_z - 1], It, dt, ut, Ft);
                        ab[v] = absorbed_from_ends(It, Ig, dt, dg, ut, ug);
                        bad = bad || !isfinite(ab[v]);
                    }
                    if (out.status && bad) atomicOr(out.status + s, CRT1D_STATUS_NONFINITE);
                    for (int q = 0; q < (out.absorbed ? out.n_bw : 0); ++q) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) a4[q] += out.band_w[(int64_t)q * n_wl + c0 + v] * ab[v];
                    }
                }
                if (out.absorbed) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double v = a4[q];
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                        if (lane == 0) partial[chunk][q] = v;
                    }
                }
            }
            if constexpr (DIAG) continue;
            rows_publish(ready, chunk, lane);
        } else {
            rows_wait(ready, chunk, lane);
            if (valid) {
                if constexpr (VEC == 2) {
                    const double2 a0 = *reinterpret_cast<const double2*>(cf + 0 * ld + c0);
                    const double2 a1 = *reinterpret_cast<const double2*>(cf + 1 * ld + c0);
                    const double2 a2 = *reinterpret_cast<const double2*>(cf + 2 * ld + c0);
                    const double2 a3 = *reinterpret_cast<const double2*>(cf + 3 * ld + c0);
                    const double2 a4 = *reinterpret_cast<const double2*>(cf + 4 * ld + c0);
                    const double2 a5 = *reinterpret_cast<const double2*>(cf + 5 * ld + c0);
                    const double2 a6 = *reinterpret_cast<const double2*>(cf + 6 * ld + c0);
                    const double2 a7 = *reinterpret_cast<const double2*>(cf + 7 * ld + c0);
                    k[0] = {a0.x, a1.x, a2.x, a3.x, a4.x, a5.x, a6.x, a7.x};
                    k[VEC - 1] = {a0.y, a1.y, a2.y, a3.y, a4.y, a5.y, a6.y, a7.y};
                } else {
                    k[0] = {cf[0 * ld + c0], cf[1 * ld + c0], cf[2 * ld + c0], cf[3 * ld + c0],
                            cf[4 * ld + c0], cf[5 * ld + c0], cf[6 * ld + c0], cf[7 * ld + c0]};
                }
            }
        }
        if (!valid || !any_profile) continue;

        const int j0 = lg * LV, j1 = min(n_z, j0 + LV);
        // Equally spaced levels inside the group (every profile the reference's LAI generators make:
        // lai = linspace(1, 0, n) * LAI, ref ../leaf_area.py:82-88): e^{-+h L_j} advance by the constant
        // factor e^{+-h dL}, so only the first level of the group needs exponentials.  Drift <= LV ulp.
        const bool uniform = REC && lg < 1024 && grp_uniform[lg] != 0;
        // Running store cursors (one add per field per level instead of 64-bit multiply-adds), and the common
        // "all four profiles requested, equally spaced group" case gets a loop without per-store null checks
        // and without per-level path selects: the sweep is issue/energy sensitive under the 1 kW power cap.
        using ST = typename std::conditional<F32, float, double>::type;
        const int64_t o0 = (int64_t)j0 * n_wl + c0;
        ST* qI = pI ? static_cast<ST*>(pI) + o0 : nullptr;
        ST* qD = pD ? static_cast<ST*>(pD) + o0 : nullptr;
        ST* qU = pU ? static_cast<ST*>(pU) + o0 : nullptr;
        ST* qF = pF ? static_cast<ST*>(pF) + o0 : nullptr;
        if (uniform && all_profiles) {
            double em[VEC], ep[VEC], qm[VEC], qp[VEC];
            const double dL = L[j0] - L[j0 + 1];  // > 0: levels run from the ground (largest L) upwards
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                exp_pm(k[v].h * L[j0], em[v], ep[v]);
                exp_pm(k[v].h * dL, qp[v], qm[v]);  // qp = e^{-h dL} multiplies e^{+hL}; qm = e^{+h dL} multiplies e^{-hL}
            }
            // one per-thread cursor + the CTA-uniform byte distances between the fields: 42 instead of 49 instructions per
            // level (four 64-bit cursors cost four IMAD.WIDE and ~10 register moves); A/B on one box 0.9345 -> 0.9366
            const int64_t dD = reinterpret_cast<char*>(qD) - reinterpret_cast<char*>(qI);
            const int64_t dU = reinterpret_cast<char*>(qU) - reinterpret_cast<char*>(qI);
            const int64_t dF = reinterpret_cast<char*>(qF) - reinterpret_cast<char*>(qI);
            char* q = reinterpret_cast<char*>(qI);
            const int64_t row_bytes = (int64_t)n_wl * (int64_t)sizeof(ST);
            for (int j = j0; j < j1; ++j) {
                const double eKj = eK[j];
                double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    level_2s_e(k[v], sc.inv_mu, eKj, em[v], ep[v], Idr[v], dn[v], up[v], F[v]);
                    em[v] *= qm[v];
                    ep[v] *= qp[v];
                }
                st_raw<VEC>(reinterpret_cast<ST*>(q), Idr);
                st_raw<VEC>(reinterpret_cast<ST*>(q + dD), dn);
                st_raw<VEC>(reinterpret_cast<ST*>(q + dU), up);
                st_raw<VEC>(reinterpret_cast<ST*>(q + dF), F);
                q += row_bytes;
            }
        } else {  // general path: irregular spacing and/or some profiles not requested
            double em[VEC], ep[VEC], qm[VEC], qp[VEC];
            if (uniform) {
                const double dL = L[j0] - L[j0 + 1];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    exp_pm(k[v].h * L[j0], em[v], ep[v]);
                    exp_pm(k[v].h * dL, qp[v], qm[v]);
                }
            }
            for (int j = j0; j < j1; ++j) {
                const double Lj = L[j], eKj = eK[j];
                double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (uniform) {
                        level_2s_e(k[v], sc.inv_mu, eKj, em[v], ep[v], Idr[v], dn[v], up[v], F[v]);
                        em[v] *= qm[v];
                        ep[v] *= qp[v];
                    } else {
                        level_2s(k[v], sc.inv_mu, Lj, eKj, Idr[v], dn[v], up[v], F[v]);
                    }
                }
                if (qI) { st_raw<VEC>(qI, Idr); qI += n_wl; }
                if (qD) { st_raw<VEC>(qD, dn); qD += n_wl; }
                if (qU) { st_raw<VEC>(qU, up); qU += n_wl; }
                if (qF) { st_raw<VEC>(qF, F); qF += n_wl; }
            }
        }
    }

    if (out.absorbed) {  // chunk sums added in chunk order: deterministic
        __syncthreads();
        if (threadIdx.x < out.n_bw) {
            double v = 0.0;
            for (int q = 0; q < n_chunks; ++q) v += partial[q][threadIdx.x];
            out.absorbed[s * out.n_bw + threadIdx.x] = v;
        }
    }
}

static size_t rows_2s_shared_bytes(int n_z, int n_wl) {
    const int ld = (n_wl + 1) & ~1;
    return (size_t)(2 * n_z + ((2 * n_z) & 1) + 8 * ld) * sizeof(double);
}

// Reduced-diagnostic kernels: __launch_bounds__(256, MINB) and threads per CTA, from the measured sweep
// (profiles/r01_reduced_diagnostic_mode.json; ms per 10^6 scenarios x 2100 bands, (MINB, threads)):
//   2s  (2,256) 58  (2,128) 51  (2,64) 49  (3,128) 45  (4,128) 44.7      bl  52 / 41 / 35 / 31 / 25.5
//   bf  65 / 52 / 45 / 38 / 33                                           g77 70 / 57 / 50 / 44 / 41
//   4s  114 / 102 / 100 / 104 / 107  -> 4s keeps 128 registers (its eigen-system spills badly below that)
#ifndef CRT_DIAG_MINB
#define CRT_DIAG_MINB 4
#endif
#ifndef CRT_DIAG_MINB_4S
#define CRT_DIAG_MINB_4S 2
#endif
static bool no_profile_requested(const crt1d_out& out) {
    return !out.I_dr && !out.I_df_d && !out.I_df_u && !out.F && !out.x0 && !out.x1 && !out.x2;
}
static int diag_threads(int dflt) {  // CRT1D_B200_DIAG_THREADS: tuning override (multiple of 32, <= 256)
    const int t = tuning().diag_threads;
    return (t >= 32 && t <= 256 && t % 32 == 0) ? t : dflt;
}

template <int VEC, int LV, int MAXT, bool REC, bool F32 = false>
static cudaError_t launch_rows_2s_t(const crt1d_batch& in, const crt1d_out& out, int threads, cudaStream_t stream) {
    if (!grid_ok(in.n_scen)) return cudaErrorInvalidConfiguration;
    if (no_profile_requested(out)) {
        const size_t smem = (size_t)(2 * in.n_z + 2) * sizeof(double);
        auto kd = solve_2s_rows_kernel<VEC, 10, 256, false, false, CRT_DIAG_MINB, true>;  // LV, REC, F32 play no part
        if (cudaError_t ed = ensure_smem(kd, smem); ed != cudaSuccess) return ed;
        kd<<<(unsigned)in.n_scen, diag_threads(128), smem, stream>>>(in, out);
        return cudaGetLastError();
    }
    const size_t smem = rows_2s_shared_bytes(in.n_z, in.n_wl);
    auto kern = solve_2s_rows_kernel<VEC, LV, MAXT, REC, F32>;
    if (cudaError_t e = ensure_smem(kern, smem); e != cudaSuccess) return e;
    kern<<<(unsigned)in.n_scen, threads, smem, stream>>>(in, out);
    return cudaGetLastError();
}

// rows-kernel configuration: default (LV = 10, 512 threads, recurrence) = best of the measured sweeps
// (profiles/r01_rows_kernel_config_sweep.txt and the later fused-coefficient runs: LV 10 > 6 > 4 by 1-3 % each).
// "LV,threads,rec" in CRT1D_B200_ROWS_CFG selects one of the few other compiled configurations (tests / tuning):
// 6,<=512,0|1   4,<=512,0|1   3,<=1024,1.
template <int VEC>
static cudaError_t launch_rows_2s(const crt1d_batch& in, const crt1d_out& out, cudaStream_t stream) {
    int lv = tuning().rows_lv, th = tuning().rows_th, rec = tuning().rows_rec;
    if (th % 32 != 0 || th < 64 || th > 1024) th = 512;
    if (in.n_z > 1024 * 2) rec = 0;  // the per-group spacing flags cover 1024 level groups (guarded again in the kernel)
    if (out.profile_f32)
        return rec ? launch_rows_2s_t<VEC, 10, 512, true, true>(in, out, th > 512 ? 512 : th, stream)
                   : launch_rows_2s_t<VEC, 10, 512, false, true>(in, out, th > 512 ? 512 : th, stream);
    if (lv == 3 && rec) return launch_rows_2s_t<VEC, 3, 1024, true>(in, out, th, stream);
    if (th > 512) th = 512;
    if (lv == 4) return rec ? launch_rows_2s_t<VEC, 4, 512, true>(in, out, th, stream)
                            : launch_rows_2s_t<VEC, 4, 512, false>(in, out, th, stream);
    if (lv == 6) return rec ? launch_rows_2s_t<VEC, 6, 512, true>(in, out, th, stream)
                            : launch_rows_2s_t<VEC, 6, 512, false>(in, out, th, stream);
    return rec ? launch_rows_2s_t<VEC, 10, 512, true>(in, out, th, stream)
               : launch_rows_2s_t<VEC, 10, 512, false>(in, out, th, stream);
}

// ---------------------------------------------------------------------------------------------
// Generic row-sweep kernel for the other closed-form schemes (bl, bf, g77, 4s): same structure as
// solve_2s_rows_kernel -- one CTA per scenario, folded per-band coefficients in shared memory (I_dr0
// comes from the spectra library through L1), row-major work items; 4s additionally advances its four
// exponentials by constant factors inside an equally spaced level group (LV = 10).
// ---------------------------------------------------------------------------------------------
template <int SCHEME>
struct RowsTraits;

template <>
struct RowsTraits<CRT1D_SCHEME_BL> {
    static constexpr int NC = 2, NF = 4;
    using Scen = ScenBl;
    using Coef = CoefBl;
    static __device__ __forceinline__ Scen scen(const crt1d_batch& in, int64_t s, const double*) {
        Scen sc;
        sc.K_b = in.K_b[s];
        sc.inv_mu = 1.0 / cos(in.psi[s]);
        return sc;
    }
    static __device__ __forceinline__ Coef coef(const Scen& sc, const BandIn<1>& b) {
        return coef_bl(sc, b.leaf_r[0], b.leaf_t[0], b.Idr0[0], b.Idf0[0]);
    }
    static __device__ __forceinline__ void pack(const Coef& k, double (&a)[NC]) { a[0] = k.Kg; a[1] = k.Idf0; }
    static __device__ __forceinline__ Coef unpack(const double (&a)[NC], double Idr0) { return {a[0], Idr0, a[1]}; }
    static __device__ __forceinline__ double rho_c(const Coef&) { return 0.0; }
    static __device__ __forceinline__ void level(const Scen& sc, const Coef& k, const double* tab, int n_z, int j,
                                                 double (&f)[NF]) {
        level_bl(sc, k, tab[j], tab[n_z + j], tab[2 * n_z + j], f[0], f[1], f[2], f[3]);
    }
};

template <bool G77>
struct RowsTraitsBfg {
    static constexpr int NC = 10, NF = 7;
    using Scen = ScenBf;
    using Coef = CoefBf;
    static __device__ __forceinline__ Scen scen(const crt1d_batch& in, int64_t s, const double* tab) {
        return scen_bf(in.psi[s], in.K_b[s], tab[0]);
    }
    static __device__ __forceinline__ Coef coef(const Scen& sc, const BandIn<1>& b) {
        return G77 ? coef_g77(sc, b.leaf_r[0], b.leaf_t[0], b.soil_r[0], b.Idr0[0], b.Idf0[0])
                   : coef_bf(sc, b.leaf_r[0], b.leaf_t[0], b.soil_r[0], b.Idr0[0], b.Idf0[0]);
    }
    static __device__ __forceinline__ void pack(const Coef& k, double (&a)[NC]) {
        a[0] = k.k_d; a[1] = k.ed0; a[2] = k.adf; a[3] = k.a1; a[4] = k.a2;
        a[5] = k.soil; a[6] = k.c1; a[7] = k.c2; a[8] = k.c3; a[9] = k.kg;
    }
    static __device__ __forceinline__ Coef unpack(const double (&a)[NC], double Idr0) {
        Coef k;
        k.k_d = a[0]; k.ed0 = a[1]; k.adf = a[2]; k.Idr0 = Idr0; k.a1 = a[3]; k.a2 = a[4];
        k.soil = a[5]; k.c1 = a[6]; k.c2 = a[7]; k.c3 = a[8]; k.kg = a[9]; k.rho_c = 0.0;
        return k;
    }
    static __device__ __forceinline__ double rho_c(const Coef& k) { return k.rho_c; }
    static __device__ __forceinline__ void level(const Scen& sc, const Coef& k, const double* tab, int n_z, int j,
                                                 double (&f)[NF]) {
        level_bfg<G77>(sc, k, tab[j], tab[n_z + j], f);
    }
};
template <>
struct RowsTraits<CRT1D_SCHEME_BF> : RowsTraitsBfg<false> {};
template <>
struct RowsTraits<CRT1D_SCHEME_G77> : RowsTraitsBfg<true> {};

template <>
struct RowsTraits<CRT1D_SCHEME_4S> {
    static constexpr int NC = 12, NF = 4;
    using Scen = Scen4s;
    using Coef = Coef4s;
    static __device__ __forceinline__ Scen scen(const crt1d_batch& in, int64_t s, const double* tab) {
        const double mu_s = in.mu_s > 0.0 ? in.mu_s : 0.501;
        return scen_4s(in.psi[s], in.K_b[s], in.G_int[2 * s], in.G_int[2 * s + 1], mu_s, tab[0]);
    }
    static __device__ __forceinline__ Coef coef(const Scen& sc, const BandIn<1>& b) {
        return coef_4s(sc, b.leaf_r[0], b.leaf_t[0], b.soil_r[0], b.Idr0[0], b.Idf0[0]);
    }
    static __device__ __forceinline__ void pack(const Coef& k, double (&a)[NC]) {
        a[0] = k.lam[0]; a[1] = k.lam[1]; a[2] = k.dnP[0]; a[3] = k.dnP[1]; a[4] = k.dnM[0]; a[5] = k.dnM[1];
        a[6] = k.upP[0]; a[7] = k.upP[1]; a[8] = k.upM[0]; a[9] = k.upM[1]; a[10] = k.dnK; a[11] = k.upK;
    }
    static __device__ __forceinline__ Coef unpack(const double (&a)[NC], double Idr0) {
        Coef k;
        k.lam[0] = a[0]; k.lam[1] = a[1]; k.dnP[0] = a[2]; k.dnP[1] = a[3]; k.dnM[0] = a[4]; k.dnM[1] = a[5];
        k.upP[0] = a[6]; k.upP[1] = a[7]; k.upM[0] = a[8]; k.upM[1] = a[9]; k.dnK = a[10]; k.upK = a[11];
        k.Idr0 = Idr0;
        return k;
    }
    static __device__ __forceinline__ double rho_c(const Coef&) { return 0.0; }
    static __device__ __forceinline__ void level(const Scen& sc, const Coef& k, const double* tab, int n_z, int j,
                                                 double (&f)[NF]) {
        level_4s_plain(sc, k, tab[j], tab[n_z + j], f[0], f[1], f[2], f[3]);
    }
};

template <int SCHEME, int VEC, int LV, int MAXT, bool F32, bool FUSED, int MINB, bool DIAG = false>
__global__ void __launch_bounds__(MAXT, MINB) solve_rows_kernel(const crt1d_batch in, const crt1d_out out, int split,
                                                                unsigned* __restrict__ rare_mask) {
    using TR = RowsTraits<SCHEME>;
    constexpr int NC = TR::NC, NF = TR::NF;
    extern __shared__ double sm[];  // [level tables][coefficients NC x ld][partial sums chunks x 4][ready flags chunks]
    __shared__ int counter;
    __shared__ unsigned char grp_uniform[1024];  // 4s: level group is equally spaced (recurrence allowed)

    // `split` CTAs share a scenario, each owning a contiguous range of band chunks (split = 2 for 4s: two CTAs
    // per SM, so one CTA's store-free coefficient phase overlaps the other's level sweeps)
    const int64_t s = blockIdx.x / split;
    const int part = blockIdx.x - (int)s * split;
    const int n_z = in.n_z, n_wl = in.n_wl, T = blockDim.x;
    const int n_tab = n_level_tables(SCHEME) * n_z;
    const int n_grp = n_wl / VEC;
    const int chunks_all = (n_grp + 31) / 32;
    const int chunks_per_part = (chunks_all + split - 1) / split;
    const int chunk_lo = part * chunks_per_part;
    const int n_chunks = max(0, min(chunks_all, chunk_lo + chunks_per_part) - chunk_lo);  // this CTA's chunks
    const int col_lo = chunk_lo * 32 * VEC;
    const int ld = DIAG ? 0 : min((n_wl + 1) & ~1, chunks_per_part * 32 * VEC);  // DIAG: no coefficient array
    double* cf = sm + n_tab + (n_tab & 1) - col_lo;  // [NC][ld] indexed by GLOBAL column, 16-byte aligned
    double (*partial)[4] = reinterpret_cast<double (*)[4]>(sm + n_tab + (n_tab & 1) + (size_t)NC * ld);
    int* ready = reinterpret_cast<int*>(partial + chunks_per_part);

    // ---- phase A: level tables, flags
    for (int j = threadIdx.x; j < n_z; j += T) fill_level_tables<SCHEME>(in, s, j, sm);
    for (int q = threadIdx.x; q < n_chunks; q += T) ready[q] = 0;
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    // level recurrence for g77 (two exponentials per level): A/B on one box 0.863 -> 0.900 of HBM peak.  bf (one
    // exponential per level) gained nothing (0.929 -> 0.931) and keeps the direct evaluation.
    constexpr bool BFG = SCHEME == CRT1D_SCHEME_G77;
    if constexpr ((SCHEME == CRT1D_SCHEME_4S || BFG) && !DIAG) {
        const double tol = 8.0 * 2.220446049250313e-16 * sm[0];
        for (int g = threadIdx.x; g < (n_z + LV - 1) / LV; g += T) {
            const int a = g * LV, b = min(n_z, a + LV);
            bool u = (b - a) >= 3 && g < 1024;
            for (int j = a + 1; j + 1 < b; ++j) u = u && fabs((sm[j] - sm[j + 1]) - (sm[a] - sm[a + 1])) <= tol;
            if (g < 1024) grp_uniform[g] = u ? 1 : 0;
        }
        __syncthreads();
    }

    const typename TR::Scen sc = TR::scen(in, s, sm);
    const int64_t prof = (int64_t)n_z * n_wl;
    double* praw[7] = {out.I_dr, out.I_df_d, out.I_df_u, out.F, out.x0, out.x1, out.x2};
    void* pf[7];
    bool any_profile = false;  // reduced-diagnostic mode: only the first item of every chunk runs
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        pf[q] = prof_base(praw[q], F32, s * prof);
        any_profile = any_profile || (!DIAG && q < NF && pf[q] != nullptr);
    }
    const double* idr0 = in.I_dr0_lib + (int64_t)in.sky_idx[s] * n_wl;
    const int n_lg = any_profile ? (n_z + LV - 1) / LV : 1;
    const int n_items = n_chunks * n_lg;
    const int lane = threadIdx.x & 31;

    // Coefficients of one chunk (32*VEC bands, one lane per VEC adjacent bands) -> shared memory, plus the chunk's
    // contribution to the canopy-absorbed sums from its ground and top levels (warp-reduced in fixed lane order).
    auto produce = [&](int chunk, typename TR::Coef (&k)[VEC]) {
        const int g = (chunk_lo + chunk) * 32 + lane;
        const bool valid = g < n_grp;
        const int c0 = (valid ? g : 0) * VEC;
        double a4[4] = {0.0, 0.0, 0.0, 0.0};
        if (valid) {
            const BandIn<VEC> b = load_bands<VEC>(in, s, c0);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                BandIn<1> b1;
                b1.leaf_r[0] = b.leaf_r[v];
                b1.leaf_t[0] = b.leaf_t[v];
                b1.soil_r[0] = b.soil_r[v];
                b1.Idr0[0] = b.Idr0[v];
                b1.Idf0[0] = b.Idf0[v];
                k[v] = TR::coef(sc, b1);
                bool redone_later = false;
                if constexpr (SCHEME == CRT1D_SCHEME_4S) {
                    // rare column (vanishing / negative eigenvalue, kappa ~ lambda_k): fixup_4s_kernel, launched right after
                    // this kernel, recomputes it in full and adds its share of the absorbed sums; here it sweeps harmless
                    // finite numbers (zero coefficients) and contributes nothing
                    redone_later = coef_4s_is_rare(k[v]);
                    if (redone_later) {
                        k[v].lam[0] = k[v].lam[1] = 1.0;
                        k[v].dnK = k[v].upK = 0.0;
                        // bit (scenario, band) of the rare-column mask: what fixup_4s_kernel works from
                        atomicOr(rare_mask + s * ((n_wl + 31) >> 5) + ((c0 + v) >> 5), 1u << ((c0 + v) & 31));
                    }
                }
                double a[NC];
                TR::pack(k[v], a);
                if constexpr (!DIAG) {  // reduced-diagnostic instantiation: coefficients are never stored
#pragma unroll
                    for (int i = 0; i < NC; ++i) cf[i * ld + c0 + v] = a[i];
                }
                if (SCHEME == CRT1D_SCHEME_BF && out.rho_c) out.rho_c[s * n_wl + c0 + v] = TR::rho_c(k[v]);
                if (out.absorbed || out.status) {
                    double gnd[NF], top[NF];
                    TR::level(sc, k[v], sm, n_z, 0, gnd);
                    TR::level(sc, k[v], sm, n_z, n_z - 1, top);
                    const double ab = redone_later ? 0.0 : absorbed_from_ends(top[0], gnd[0], top[1], gnd[1], top[2], gnd[2]);
                    if (out.status && !isfinite(ab)) atomicOr(out.status + s, CRT1D_STATUS_NONFINITE);
                    for (int q = 0; q < (out.absorbed ? out.n_bw : 0); ++q) a4[q] += out.band_w[(int64_t)q * n_wl + c0 + v] * ab;
                }
            }
        }
        if (out.absorbed) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double v = a4[q];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) partial[chunk][q] = v;
            }
        }
    };

    if constexpr (!FUSED || DIAG) {  // separate coefficient phase: warps take chunks round-robin, then everyone sweeps
        typename TR::Coef kk[VEC];
        for (int chunk = threadIdx.x >> 5; chunk < n_chunks; chunk += T >> 5) produce(chunk, kk);
        if constexpr (!DIAG) __syncthreads();
    }

    // ---- row-major work items (LV levels x 32*VEC bands); FUSED: a chunk's first item produces its coefficients
    const int first_item = (FUSED || any_profile) ? 0 : n_items;  // !FUSED + no profiles: nothing left to do
    for (; !DIAG;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&counter, 1) + first_item;
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int lg = item / n_chunks, chunk = item - lg * n_chunks;
        const int g = (chunk_lo + chunk) * 32 + lane;
        const bool valid = g < n_grp;
        const int c0 = (valid ? g : 0) * VEC;
        typename TR::Coef k[VEC];
        if (FUSED && lg == 0) {
            produce(chunk, k);
            rows_publish(ready, chunk, lane);
        } else {
            if (FUSED) rows_wait(ready, chunk, lane);
            if (valid) {
                double a[VEC][NC];
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    if constexpr (VEC == 2) {
                        const double2 t = *reinterpret_cast<const double2*>(cf + i * ld + c0);
                        a[0][i] = t.x;
                        a[VEC - 1][i] = t.y;
                    } else {
                        a[0][i] = cf[i * ld + c0];
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) k[v] = TR::unpack(a[v], __ldg(idr0 + c0 + v));
            }
        }
        if (!valid || !any_profile) continue;

        const int j0 = lg * LV, j1 = min(n_z, j0 + LV);
        // 4s on an equally spaced group: anchor the four exponentials at the group's first level and advance
        // them by constant factors (same scheme as column_4s); everything else evaluates each level directly.
        double m0[VEC], p0[VEC], m1[VEC], p1[VEC], qi0[VEC], qd0[VEC], qi1[VEC], qd1[VEC];
        bool rec[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) rec[v] = false;
        if constexpr (SCHEME == CRT1D_SCHEME_4S) {
            if (lg < 1024 && grp_uniform[lg]) {
                const double dL = sm[j0] - sm[j0 + 1];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    rec[v] = k[v].lam[0] > 0.0;
                    if (rec[v]) {
                        exp_pm(k[v].lam[0] * sm[j0], m0[v], p0[v]);
                        exp_pm(k[v].lam[1] * sm[j0], m1[v], p1[v]);
                        exp_pm(k[v].lam[0] * dL, qd0[v], qi0[v]);
                        exp_pm(k[v].lam[1] * dL, qd1[v], qi1[v]);
                    }
                }
            }
        }
        // 4s fast loop: equally spaced group, both columns on the recurrence, all four profiles requested -- no per-level
        // path selects, no null checks, one store cursor + the CTA-uniform byte distances between the fields
        // (same shape as solve_2s_rows_kernel's fast loop): 57 instructions per level and lane, 40 of them FP64;
        // A/B on one box 0.861 -> 0.873 of HBM peak.  (The same store path for bl/bf/g77 changed nothing and was not kept.)
        if constexpr (SCHEME == CRT1D_SCHEME_4S) {
            bool fast = pf[0] && pf[1] && pf[2] && pf[3];
#pragma unroll
            for (int v = 0; v < VEC; ++v) fast = fast && rec[v];
            if (fast) {
                using ST = typename std::conditional<F32, float, double>::type;
                char* q = reinterpret_cast<char*>(static_cast<ST*>(pf[0]) + ((int64_t)j0 * n_wl + c0));
                const int64_t d1 = static_cast<char*>(pf[1]) - static_cast<char*>(pf[0]);
                const int64_t d2 = static_cast<char*>(pf[2]) - static_cast<char*>(pf[0]);
                const int64_t d3 = static_cast<char*>(pf[3]) - static_cast<char*>(pf[0]);
                const int64_t row_bytes = (int64_t)n_wl * (int64_t)sizeof(ST);
                for (int j = j0; j < j1; ++j) {
                    const double eKj = sm[n_z + j];
                    double Idr[VEC], dn[VEC], up[VEC], F[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        level_4s_e(sc, k[v], eKj, m0[v], p0[v], m1[v], p1[v], Idr[v], dn[v], up[v], F[v]);
                        m0[v] *= qi0[v];
                        p0[v] *= qd0[v];
                        m1[v] *= qi1[v];
                        p1[v] *= qd1[v];
                    }
                    st_raw<VEC>(reinterpret_cast<ST*>(q), Idr);
                    st_raw<VEC>(reinterpret_cast<ST*>(q + d1), dn);
                    st_raw<VEC>(reinterpret_cast<ST*>(q + d2), up);
                    st_raw<VEC>(reinterpret_cast<ST*>(q + d3), F);
                    q += row_bytes;
                }
                continue;
            }
        }
        // g77 on an equally spaced group: exp(-+k_d L) and exp(-kg L) anchored at the group's first level
        // and advanced by constant factors (levels run from the ground up: L decreases by dL per level)
        bool bfg_rec = false;
        if constexpr (BFG) {
            if (lg < 1024 && grp_uniform[lg]) {
                bfg_rec = true;
                const double dL = sm[j0] - sm[j0 + 1];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    exp_pm(k[v].k_d * sm[j0], m0[v], p0[v]);    // ed, ep
                    exp_pm(k[v].k_d * dL, qd0[v], qi0[v]);      // qd0 = e^{-k_d dL} multiplies ep; qi0 = e^{+k_d dL} multiplies ed
                    if constexpr (SCHEME == CRT1D_SCHEME_G77) {
                        m1[v] = exp_neg(k[v].kg * sm[j0]);
                        exp_pm(k[v].kg * dL, qd1[v], qi1[v]);   // qi1 = e^{+kg dL} multiplies eg
                    }
                }
            }
        }
        for (int j = j0; j < j1; ++j) {
            double o[NF][VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                double f[NF];
                if constexpr (BFG) {
                    if (bfg_rec) {
                        level_bfg_e<SCHEME == CRT1D_SCHEME_G77>(sc, k[v], sm[n_z + j], m0[v], p0[v], m1[v], f);
                        m0[v] *= qi0[v];
                        p0[v] *= qd0[v];
                        if constexpr (SCHEME == CRT1D_SCHEME_G77) m1[v] *= qi1[v];
                    } else {
                        TR::level(sc, k[v], sm, n_z, j, f);
                    }
                } else if constexpr (SCHEME == CRT1D_SCHEME_4S) {
                    if (rec[v]) {
                        level_4s_e(sc, k[v], sm[n_z + j], m0[v], p0[v], m1[v], p1[v], f[0], f[1], f[2], f[3]);
                        m0[v] *= qi0[v];
                        p0[v] *= qd0[v];
                        m1[v] *= qi1[v];
                        p1[v] *= qd1[v];
                    } else {
                        TR::level(sc, k[v], sm, n_z, j, f);
                    }
                } else {
                    TR::level(sc, k[v], sm, n_z, j, f);
                }
#pragma unroll
                for (int q = 0; q < NF; ++q) o[q][v] = f[q];
            }
            const int64_t off = (int64_t)j * n_wl + c0;
#pragma unroll
            for (int q = 0; q < NF; ++q) st_prof_t<VEC, F32>(pf[q], off, o[q]);
        }
    }

    if (out.absorbed) {  // chunk sums added in chunk order: deterministic
        __syncthreads();
        if (threadIdx.x < out.n_bw) {
            double v = 0.0;
            for (int q = 0; q < n_chunks; ++q) v += partial[q][threadIdx.x];
            if (split == 1) {
                out.absorbed[s * out.n_bw + threadIdx.x] = v;
            } else if (split == 2) {  // two parts into a zeroed sum: 0 + a + b = 0 + b + a exactly, so still deterministic
                atomicAdd(&out.absorbed[s * out.n_bw + threadIdx.x], v);
            } else {  // three or more parts: the launcher passed a scratch array [scenario][part][band group]; parts_sum_kernel adds them in part order
                out.absorbed[(s * split + part) * out.n_bw + threadIdx.x] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4s, row-sweep path: the RARE columns.  coef_4s hands a column over when its smaller eigenvalue^2 is (nearly) zero or
// negative (omega -> 1) or kappa sits within CRT_4S_RESONANCE_WIDTH of an eigenvalue (~0.2 % of a sweep's columns).
// Their general closed form (coef_4s_general: entire basis, resonance-safe particular solution, direct evaluation of
// every basis function) is 10x the code and registers of the ordinary one; compiled into solve_rows_kernel it cost
// EVERY column (0.84 -> 0.56 of HBM peak: spilled recurrence factors re-read in the level loop).  So the sweep kernel
// treats a rare column as all-zero coefficients, and this kernel -- launched on the same stream right after it --
// redoes them: one warp per scenario re-evaluates the (cheap, deterministic) rarity test of every band, and for the
// rare bands solves the general form and writes all their levels and their share of the absorbed sums (added in band
// order: deterministic).
// ---------------------------------------------------------------------------------------------
#ifndef CRT_FIXUP_MINB
#define CRT_FIXUP_MINB 20
#endif
// One WARP per (scenario, level part) (a 32-thread CTA: the work is latency-bound serial arithmetic, so what counts is
// how many warps are resident -- 20+ per SM, i.e. a whole launch of ~4000 scenarios in about one wave).  The rare columns
// come from the bit mask the sweep kernel set in its coefficient stage (one bit per (scenario, band); the first version
// re-ran the rarity test over all 2100 bands here: 66 dependent load + eigenvalue rounds per warp, 2/3 of this kernel's
// time).  A warp with an empty mask exits at once.  Otherwise it walks the set bits in band order into a pending list; a
// flush then (i) lets lane i solve the general coefficients of pending column i -- different columns on different
// lanes, in parallel -- into shared memory, (ii) shares the (column, level) evaluations of ITS levels out over all 32
// lanes, (iii) part 0 adds the columns' absorbed-sum terms in list (= band) order.  Deep canopies use several parts per
// scenario (n_parts ~ n_z / 64: at n_z = 1000 one warp per scenario was 21 % of the 4s step).
struct FixupSlot {
    double c[13];  // Coef4s, field order of RowsTraits<4S>::pack + Idr0
};
__global__ void __launch_bounds__(32, CRT_FIXUP_MINB) fixup_4s_kernel(const crt1d_batch in, const crt1d_out out,
                                                                      const unsigned* __restrict__ rare_mask, int n_parts) {
    extern __shared__ double tab[];  // L[j], exp(-K_b L[j]), then 32 coefficient slots
    __shared__ int pend[32];
    const int64_t s = blockIdx.x / n_parts;
    const int part = blockIdx.x - (int)s * n_parts;
    const int n_z = in.n_z, n_wl = in.n_wl;
    const int lane = threadIdx.x;
    const int n_words = (n_wl + 31) >> 5;
    const unsigned* const mask = rare_mask + s * n_words;
    bool any = false;
    for (int w = lane; w < n_words; w += 32) any = any || mask[w] != 0u;
    if (!__any_sync(0xffffffffu, any)) return;

    const int j_lo = (int)((int64_t)n_z * part / n_parts), j_hi = (int)((int64_t)n_z * (part + 1) / n_parts);  // this warp's levels
    FixupSlot* const slot = reinterpret_cast<FixupSlot*>(tab + 2 * n_z);
    for (int j = j_lo + lane; j < j_hi; j += 32) fill_level_tables<CRT1D_SCHEME_4S>(in, s, j, tab);
    if (lane == 0 && j_lo > 0) fill_level_tables<CRT1D_SCHEME_4S>(in, s, 0, tab);
    if (lane == 1 && j_hi < n_z) fill_level_tables<CRT1D_SCHEME_4S>(in, s, n_z - 1, tab);
    __syncwarp();
    const double mu_s = in.mu_s > 0.0 ? in.mu_s : 0.501;
    const Scen4s sc = scen_4s(in.psi[s], in.K_b[s], in.G_int[2 * s], in.G_int[2 * s + 1], mu_s, tab[0]);
    const bool f32 = out.profile_f32 != 0;
    const int64_t prof = (int64_t)n_z * n_wl;
    void* const pf[4] = {prof_base(out.I_dr, f32, s * prof), prof_base(out.I_df_d, f32, s * prof),
                         prof_base(out.I_df_u, f32, s * prof), prof_base(out.F, f32, s * prof)};
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int n = 0;

    auto load_slot = [&](int i) {
        Coef4s k;
        const double* c = slot[i].c;
        k.lam[0] = c[0]; k.lam[1] = c[1]; k.dnP[0] = c[2]; k.dnP[1] = c[3]; k.dnM[0] = c[4]; k.dnM[1] = c[5];
        k.upP[0] = c[6]; k.upP[1] = c[7]; k.upM[0] = c[8]; k.upM[1] = c[9]; k.dnK = c[10]; k.upK = c[11]; k.Idr0 = c[12];
        return k;
    };
    auto flush = [&]() {
        double ab = 0.0;
        if (lane < n) {  // (i) coefficients, one pending column per lane
            const int cc = pend[lane];
            const BandIn<1> b = load_bands<1>(in, s, cc);
            Coef4s k;
            coef_4s_rare(&sc, b.leaf_r[0], b.leaf_t[0], b.soil_r[0], b.Idr0[0], b.Idf0[0], &k);
            double* c = slot[lane].c;
            c[0] = k.lam[0]; c[1] = k.lam[1]; c[2] = k.dnP[0]; c[3] = k.dnP[1]; c[4] = k.dnM[0]; c[5] = k.dnM[1];
            c[6] = k.upP[0]; c[7] = k.upP[1]; c[8] = k.upM[0]; c[9] = k.upM[1]; c[10] = k.dnK; c[11] = k.upK; c[12] = k.Idr0;
            if (part == 0) {
                double g[4], t[4];  // its ground and top levels: the absorbed-sum term
                level_4s(sc, k, tab[0], tab[n_z], g[0], g[1], g[2], g[3]);
                level_4s(sc, k, tab[n_z - 1], tab[2 * n_z - 1], t[0], t[1], t[2], t[3]);
                ab = absorbed_from_ends(t[0], g[0], t[1], g[1], t[2], g[2]);
                if (out.status && !isfinite(ab)) atomicOr(out.status + s, CRT1D_STATUS_NONFINITE);
            }
        }
        __syncwarp();
        const int n_lev = j_hi - j_lo;
        const int items = n * n_lev;  // (ii) this warp's levels of every pending column
        for (int it = lane; it < items; it += 32) {
            const int col = it / n_lev, j = j_lo + (it - col * n_lev);
            const Coef4s k = load_slot(col);
            double f[4][1];
            level_4s(sc, k, tab[j], tab[n_z + j], f[0][0], f[1][0], f[2][0], f[3][0]);
            const int64_t off = (int64_t)j * n_wl + pend[col];
#pragma unroll
            for (int q = 0; q < 4; ++q) st_prof<1>(pf[q], f32, off, f[q]);
        }
        if (out.absorbed && part == 0) {  // (iii) in list order, the same sum on every lane
            for (int col = 0; col < n; ++col) {
                const double a = __shfl_sync(0xffffffffu, ab, col);
                const int cc = pend[col];
                for (int q = 0; q < out.n_bw; ++q) acc[q] += out.band_w[(int64_t)q * n_wl + cc] * a;
            }
        }
        __syncwarp();
        n = 0;
    };

    for (int w0 = 0; w0 < n_words; w0 += 32) {  // set bits in band order; control flow is warp-uniform
        const unsigned m = (w0 + lane < n_words) ? mask[w0 + lane] : 0u;
        unsigned lanes = __ballot_sync(0xffffffffu, m != 0u);
        while (lanes) {
            const int src = __ffs(lanes) - 1;
            lanes &= lanes - 1;
            unsigned word = __shfl_sync(0xffffffffu, m, src);
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                if (lane == 0) pend[n] = (w0 + src) * 32 + bit;
                ++n;
                __syncwarp();
                if (n == 32) flush();
            }
        }
    }
    if (n > 0) flush();
    if (out.absorbed && part == 0 && lane < out.n_bw) {
        double v = 0.0;  // lane q keeps acc[q]; static indexing only
#pragma unroll
        for (int q = 0; q < 4; ++q) v = lane == q ? acc[q] : v;
        out.absorbed[s * out.n_bw + lane] += v;  // after the sweep kernel's own sum: stream order
    }
}
static cudaError_t launch_fixup_4s(const crt1d_batch& in, const crt1d_out& out, const unsigned* rare_mask, cudaStream_t stream) {
    const size_t smem = (size_t)2 * in.n_z * sizeof(double) + 32 * sizeof(FixupSlot);
    if (cudaError_t e = ensure_smem(fixup_4s_kernel, smem); e != cudaSuccess) return e;
    int n_parts = std::max(1, std::min(32, in.n_z / 64));
    if (tuning().fixup_parts > 0) n_parts = std::min(tuning().fixup_parts, in.n_z);
    if (!grid_ok(in.n_scen * n_parts)) return cudaErrorInvalidConfiguration;
    fixup_4s_kernel<<<(unsigned)(in.n_scen * n_parts), 32, smem, stream>>>(in, out, rare_mask, n_parts);
    return cudaGetLastError();
}
// the sweep kernel's rare-column bit mask: stream-ordered scratch, zeroed
static cudaError_t rare_mask_alloc(const crt1d_batch& in, unsigned** mask, cudaStream_t stream) {
    const size_t bytes = (size_t)in.n_scen * ((in.n_wl + 31) / 32) * sizeof(unsigned);
    if (cudaError_t e = scratch_alloc(reinterpret_cast<void**>(mask), bytes, stream); e != cudaSuccess) return e;
    return cudaMemsetAsync(*mask, 0, bytes, stream);
}

// absorbed[s][k] = sum of the `split` parts of scenario s, in part order (deterministic)
__global__ void __launch_bounds__(256) parts_sum_kernel(const double* __restrict__ parts, int64_t n_scen, int split, int n_bw,
                                                        double* __restrict__ absorbed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scen * n_bw) return;
    const int64_t s = i / n_bw;
    const int k = (int)(i - s * n_bw);
    double v = 0.0;
    for (int p = 0; p < split; ++p) v += parts[(s * split + p) * n_bw + k];
    absorbed[i] = v;
}

template <int SCHEME>
static size_t rows_shared_bytes(int n_z, int n_wl, int split = 1, int vec = 2) {
    int ld = (n_wl + 1) & ~1;
    if (split > 1) {
        const int chunks_all = (n_wl / vec + 31) / 32;
        ld = min(ld, (chunks_all + split - 1) / split * 32 * vec);
    }
    const int n_tab = n_level_tables(SCHEME) * n_z;
    const int chunks_per_part = ((n_wl / vec + 31) / 32 + split - 1) / split;
    return (size_t)(n_tab + (n_tab & 1) + RowsTraits<SCHEME>::NC * ld) * sizeof(double) + (size_t)chunks_per_part * (4 * sizeof(double) + sizeof(int));
}

// 4s band split of a launch: the configured split (2), widened to 3 or 4 when two CTAs would not fit (deep canopies);
// 1 = no split possible.  `smem2` = dynamic shared memory of one CTA at that split.
static int split_4s_for(int n_z, int n_wl, int vec, size_t& smem2) {
    int split = tuning().split_4s;
    smem2 = rows_shared_bytes<CRT1D_SCHEME_4S>(n_z, n_wl, split, vec);
    while (split >= 2 && split < 4 && 2 * (smem2 + 3 * 1024) > 228u * 1024u) {
        ++split;
        smem2 = rows_shared_bytes<CRT1D_SCHEME_4S>(n_z, n_wl, split, vec);
    }
    return (split >= 2 && 2 * (smem2 + 3 * 1024) <= 228u * 1024u) ? split : 1;
}

template <int SCHEME, int VEC, int LV, int MAXT, bool F32>
static cudaError_t launch_rows_t(const crt1d_batch& in, const crt1d_out& out, int th, size_t smem, cudaStream_t stream) {
    // 4s: its coefficient stage (eigen-system + 4x4 solve, 168 registers) is better kept as a separate phase
    // (0.82 vs 0.77 of HBM peak when fused into the first item); bl/bf/g77 fuse it (no store-free phase).
    constexpr bool FUSED = SCHEME != CRT1D_SCHEME_4S;
    constexpr bool IS4S = SCHEME == CRT1D_SCHEME_4S;
    if (!grid_ok(in.n_scen)) return cudaErrorInvalidConfiguration;
    unsigned* rare = nullptr;  // 4s: (scenario, band) bits of the columns fixup_4s_kernel redoes
    if constexpr (IS4S) {
        if (cudaError_t e = rare_mask_alloc(in, &rare, stream); e != cudaSuccess) return e;
    }
    // every path ends here: 4s runs its fix-up kernel behind the sweep and returns the mask to the pool
    auto finish = [&](cudaError_t e) {
        if constexpr (IS4S) {
            if (e == cudaSuccess) e = launch_fixup_4s(in, out, rare, stream);
            cudaFreeAsync(rare, stream);
        }
        return e;
    };
    if (no_profile_requested(out)) {  // reduced-diagnostic mode: small CTAs, several per SM, level tables only
        const int n_tab = n_level_tables(SCHEME) * in.n_z;
        const int chunks = (in.n_wl / VEC + 31) / 32;
        const size_t smem_d = (size_t)(n_tab + (n_tab & 1)) * sizeof(double) + (size_t)chunks * (4 * sizeof(double) + sizeof(int));
        constexpr int DM = IS4S ? CRT_DIAG_MINB_4S : CRT_DIAG_MINB;
        auto kd = solve_rows_kernel<SCHEME, VEC, 10, 256, false, FUSED, DM, true>;  // LV, F32 play no part
        if (cudaError_t e = ensure_smem(kd, smem_d); e != cudaSuccess) return finish(e);
        kd<<<(unsigned)in.n_scen, diag_threads(IS4S ? 64 : 128), smem_d, stream>>>(in, out, 1, rare);
        return finish(cudaGetLastError());
    }
    if constexpr (IS4S) {
        // Two resident CTAs per SM (register bound), each on a share of the scenario's band chunks: one CTA's store-free
        // coefficient phase (~20 % of its life) overlaps the other's level sweeps.  First version: halves (2 x ~106 KB of
        // coefficients); quarters are better at every depth (60 levels 0.833 -> 0.853, 1000 levels 0.776 (thirds) -> 0.815:
        // shorter coefficient phases, finer interleaving) and leave room for the level tables of deep canopies.
        size_t smem2 = 0;
        const int split = split_4s_for(in.n_z, in.n_wl, VEC, smem2);
        if (split >= 2 && in.n_scen * split <= 2147483647LL) {
            auto kern2 = solve_rows_kernel<SCHEME, VEC, LV, MAXT / 2, F32, FUSED, 2>;
            cudaError_t e = ensure_smem(kern2, smem2);
            if (e != cudaSuccess) return finish(e);
            crt1d_out o2 = out;
            double* parts = nullptr;
            if (out.absorbed && split == 2) {
                e = cudaMemsetAsync(out.absorbed, 0, (size_t)in.n_scen * out.n_bw * sizeof(double), stream);
                if (e != cudaSuccess) return finish(e);
            } else if (out.absorbed) {
                e = scratch_alloc(reinterpret_cast<void**>(&parts), (size_t)in.n_scen * split * out.n_bw * sizeof(double), stream);
                if (e != cudaSuccess) return finish(e);
                o2.absorbed = parts;
            }
            kern2<<<(unsigned)(in.n_scen * split), min(th, MAXT / 2), smem2, stream>>>(in, o2, split, rare);
            e = cudaGetLastError();
            if (parts) {
                if (e == cudaSuccess) {
                    const int64_t n = in.n_scen * out.n_bw;
                    parts_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(parts, in.n_scen, split, out.n_bw, out.absorbed);
                    e = cudaGetLastError();
                }
                cudaFreeAsync(parts, stream);
            }
            return finish(e);
        }
    }
    auto kern = solve_rows_kernel<SCHEME, VEC, LV, MAXT, F32, FUSED, 1>;
    if (cudaError_t e = ensure_smem(kern, smem); e != cudaSuccess) return finish(e);
    kern<<<(unsigned)in.n_scen, th, smem, stream>>>(in, out, 1, rare);
    return finish(cudaGetLastError());
}

// 4s level groups: 10 levels per work item.  Measured (two CTAs per SM): 15 levels 0.80, 30 levels 0.81 vs 0.77 --
// but the gain is the smaller number of exact exponentials, and the drift of a 30-level recurrence times the
// cancellation in weak bands reached 2.2e-9 (bar: 1e-9); 30-level items re-anchored every 10 levels: 0.786 vs 0.795.
#ifndef CRT_4S_LV
#define CRT_4S_LV 10
#endif
template <int SCHEME, int MAXT, int LV = 6>
static cudaError_t launch_rows(const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    const size_t smem = rows_shared_bytes<SCHEME>(in.n_z, in.n_wl, 1, vec2 ? 2 : 1);
    int th = MAXT;
    const int rt = tuning().rows_threads;
    if (rt >= 64 && rt <= MAXT && rt % 32 == 0) th = rt;
    if (out.profile_f32)
        return vec2 ? launch_rows_t<SCHEME, 2, LV, MAXT, true>(in, out, th, smem, stream)
                    : launch_rows_t<SCHEME, 1, LV, MAXT, true>(in, out, th, smem, stream);
    return vec2 ? launch_rows_t<SCHEME, 2, LV, MAXT, false>(in, out, th, smem, stream)
                : launch_rows_t<SCHEME, 1, LV, MAXT, false>(in, out, th, smem, stream);
}

// Batches with at least this many scenarios go to the row-sweep kernels (one CTA per SM needs >= n_SM
// scenarios in flight); smaller ones (the single-scenario plugin path) use the band-tile kernel.
static int64_t scen_kernel_min_batch() { return tuning().scen_min; }

static cudaError_t launch_solve_impl(int scheme, const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream);

cudaError_t launch_solve(int scheme, const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    if (out.status) {
        cudaError_t e = cudaMemsetAsync(out.status, 0, (size_t)in.n_scen * sizeof(int32_t), stream);
        if (e != cudaSuccess) return e;
    }
    return launch_solve_impl(scheme, in, out, vec2, stream);
}

static cudaError_t launch_solve_impl(int scheme, const crt1d_batch& in, const crt1d_out& out, bool vec2, cudaStream_t stream) {
    if (tuning().force_vec1) vec2 = false;  // tuning experiments
    if (scheme == CRT1D_SCHEME_2S && in.n_scen >= scen_kernel_min_batch()) {
        const int n_chunks = (in.n_wl / (vec2 ? 2 : 1) + 31) / 32;
        const bool rows_fit = rows_2s_shared_bytes(in.n_z, in.n_wl) <= 227u * 1024u - 12288u && n_chunks <= ROWS_MAX_CHUNKS;
        if (!tuning().tile_2s && rows_fit)
            return vec2 ? launch_rows_2s<2>(in, out, stream) : launch_rows_2s<1>(in, out, stream);
    }
    if (in.n_scen >= scen_kernel_min_batch() && !tuning().no_rows) {
        const size_t cap = 227u * 1024u - 2048u;  // dynamic + ~1 KB static shared memory of one CTA
        if ((in.n_wl / (vec2 ? 2 : 1) + 31) / 32 > ROWS_MAX_CHUNKS) goto tile;
        switch (scheme) {
            case CRT1D_SCHEME_BL:
                if (rows_shared_bytes<CRT1D_SCHEME_BL>(in.n_z, in.n_wl, 1, vec2 ? 2 : 1) <= cap) return launch_rows<CRT1D_SCHEME_BL, 512>(in, out, vec2, stream);
                break;
            case CRT1D_SCHEME_BF:
                if (rows_shared_bytes<CRT1D_SCHEME_BF>(in.n_z, in.n_wl, 1, vec2 ? 2 : 1) <= cap) return launch_rows<CRT1D_SCHEME_BF, 512>(in, out, vec2, stream);
                break;
            case CRT1D_SCHEME_G77:
                if (rows_shared_bytes<CRT1D_SCHEME_G77>(in.n_z, in.n_wl, 1, vec2 ? 2 : 1) <= cap) return launch_rows<CRT1D_SCHEME_G77, 512>(in, out, vec2, stream);
                break;
            case CRT1D_SCHEME_4S:
                if (rows_shared_bytes<CRT1D_SCHEME_4S>(in.n_z, in.n_wl, 1, vec2 ? 2 : 1) <= cap)  // 0.82 vs 0.76 tiled (with the recurrence)
#ifndef CRT_4S_MAXT
#define CRT_4S_MAXT 512  // split: 2 x 256 threads at 128 registers 0.786 vs 2 x 192 at 168 registers 0.768
#endif
                    return launch_rows<CRT1D_SCHEME_4S, CRT_4S_MAXT, CRT_4S_LV>(in, out, vec2, stream);
                break;
            default: break;
        }
    }
tile:
    switch (scheme) {
        case CRT1D_SCHEME_2S: return launch_vec<CRT1D_SCHEME_2S>(in, out, vec2, stream);
        case CRT1D_SCHEME_4S: return launch_vec<CRT1D_SCHEME_4S>(in, out, vec2, stream);
        case CRT1D_SCHEME_BF: return launch_vec<CRT1D_SCHEME_BF>(in, out, vec2, stream);
        case CRT1D_SCHEME_BL: return launch_vec<CRT1D_SCHEME_BL>(in, out, vec2, stream);
        case CRT1D_SCHEME_G77: return launch_vec<CRT1D_SCHEME_G77>(in, out, vec2, stream);
        case CRT1D_SCHEME_N79: return launch_vec<CRT1D_SCHEME_N79>(in, out, vec2, stream);
        case CRT1D_SCHEME_ZQ: return launch_vec<CRT1D_SCHEME_ZQ>(in, out, vec2, stream);
        case CRT1D_SCHEME_ZQ_PA: return launch_vec<CRT1D_SCHEME_ZQ_PA>(in, out, vec2 && ZQPA_VEC == 2, stream);
        default: return cudaErrorInvalidValue;
    }
}

// Scenarios per launch, at most `max_scen`, that fill WHOLE waves of resident CTAs for the kernel launch_solve picks.
// A kernel whose CTAs all take about the same time T runs ceil(grid / resident) * T: 4.1 waves cost 5 (the deep-canopy
// zq launch of 296 scenarios: 1215 CTAs on 296 slots).  Row-sweep kernels: one CTA per scenario and SM (4s: `split`
// CTAs per scenario, two resident); flat tridiagonal kernels: grid = ceil(n gps / nthr) on n_SM x resident slots.
int64_t preferred_batch(int scheme, int n_z, int n_wl, int64_t max_scen, int n_sm) {
    if (max_scen < scen_kernel_min_batch() || n_sm <= 0) return max_scen;
    const bool vec2 = n_wl % 2 == 0 && !tuning().force_vec1;
    int64_t num = 1, den = 1;  // scenarios per wave = n_sm * num / den
    if (scheme == CRT1D_SCHEME_ZQ || scheme == CRT1D_SCHEME_N79) {
        crt1d_batch in{};
        in.n_z = n_z;
        in.n_wl = n_wl;
        constexpr int B = TileCfg<CRT1D_SCHEME_ZQ>::BLK, M = TileCfg<CRT1D_SCHEME_ZQ>::MINB;
        const int nthr = tile_threads(B, B);
        size_t smem = 0;
        bool flat;
        if (scheme == CRT1D_SCHEME_ZQ) {
            flat = false;
            if (zq_wide_spacing(n_z))  // deep canopies: one resident CTA (launch_one)
                flat = vec2 ? flat_eligible<CRT1D_SCHEME_ZQ, 2, B, M>(in, nthr, smem, ZQ_DEEP_CK) : flat_eligible<CRT1D_SCHEME_ZQ, 1, B, M>(in, nthr, smem, ZQ_DEEP_CK);
            if (!flat) flat = vec2 ? flat_eligible<CRT1D_SCHEME_ZQ, 2, B, M>(in, nthr, smem) : flat_eligible<CRT1D_SCHEME_ZQ, 1, B, M>(in, nthr, smem);
        }
        else
            flat = vec2 ? flat_eligible<CRT1D_SCHEME_N79, 2, B, M>(in, nthr, smem) : flat_eligible<CRT1D_SCHEME_N79, 1, B, M>(in, nthr, smem);
        if (!flat) return max_scen;
        const size_t per_cta = 1024u + sizeof(double) * (B / 32) * 8;
        const int64_t resident = std::min<int64_t>((int64_t)M * (B / nthr), (int64_t)((228u * 1024u) / (smem + per_cta)));
        num = resident * nthr;
        den = n_wl / (vec2 ? 2 : 1);
    } else if (scheme == CRT1D_SCHEME_4S) {
        size_t smem2 = 0;
        const int split = split_4s_for(n_z, n_wl, vec2 ? 2 : 1, smem2);
        if (split >= 2) { num = 2; den = split; }
    }
    const int64_t waves = max_scen * den / (n_sm * num);
    if (waves <= 0) return max_scen;
    return waves * n_sm * num / den;
}

// ---------------------------------------------------------------------------------------------
// layer absorption  (ref ../model.py:573-647)
// thread = VEC adjacent bands, walks up the levels carrying the level below in registers, so every
// profile value is read exactly once; per-layer scalars (1 - tau_b, f_sl) sit in shared memory.
// ---------------------------------------------------------------------------------------------
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
        if (cudaError_t e = ensure_smem(absorption_kernel<2>, smem); e != cudaSuccess) return e;
        absorption_kernel<2><<<(unsigned)grid, BLOCK, smem, stream>>>(in, I_dr, I_df_d, I_df_u, out, tiles);
    } else {
        if (cudaError_t e = ensure_smem(absorption_kernel<1>, smem); e != cudaSuccess) return e;
        absorption_kernel<1><<<(unsigned)grid, BLOCK, smem, stream>>>(in, I_dr, I_df_d, I_df_u, out, tiles);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// canopy energy balance  (ref ../diagnostics.py:476-530 with the band weights of :56-81)
// one CTA per scenario; threads stride over bands reading the ground and top rows only; fixed-order
// block reduction of 4 x n_bw sums (deterministic).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) energy_balance_kernel(int n_z, int n_wl, const double* __restrict__ I_dr,
                                                             const double* __restrict__ I_df_d,
                                                             const double* __restrict__ I_df_u,
                                                             const double* __restrict__ band_w, int n_bw, double* ebal) {
    __shared__ double red[8][16];
    const int64_t s = blockIdx.x;
    const int64_t g0 = s * (int64_t)n_z * n_wl, t0 = g0 + (int64_t)(n_z - 1) * n_wl;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    for (int b = threadIdx.x; b < n_wl; b += blockDim.x) {
        const double drg = I_dr[g0 + b], dng = I_df_d[g0 + b], upg = I_df_u[g0 + b];
        const double drt = I_dr[t0 + b], dnt = I_df_d[t0 + b], upt = I_df_u[t0 + b];
        const double incoming = drt + dnt;                                   // ref :512
        const double outgoing = upt;                                         // ref :513
        const double soil_abs = (drg + dng) - upg;                           // ref :514-516
        const double canopy = dnt - upt + drt - drg + -(dng - upg);          // ref :518-524
        for (int k = 0; k < n_bw; ++k) {
            const double w = band_w[(int64_t)k * n_wl + b];
            acc[4 * k + 0] += w * incoming;
            acc[4 * k + 1] += w * outgoing;
            acc[4 * k + 2] += w * soil_abs;
            acc[4 * k + 3] += w * canopy;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double v = acc[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4 * n_bw) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        ebal[s * 4 * n_bw + threadIdx.x] = v;
    }
}

cudaError_t launch_energy_balance(int64_t n_scen, int n_z, int n_wl, const double* I_dr, const double* I_df_d,
                                  const double* I_df_u, const double* band_w, int n_bw, double* ebal, cudaStream_t stream) {
    if (n_scen <= 0) return cudaSuccess;
    if (n_scen > 2147483647LL) return cudaErrorInvalidConfiguration;
    energy_balance_kernel<<<(unsigned)n_scen, 256, 0, stream>>>(n_z, n_wl, I_dr, I_df_d, I_df_u, band_w, n_bw, ebal);
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

// ---------------------------------------------------------------------------------------------
// spectral binning: out[r][i] = average of y_r(x) over [bins[i], bins[i+1]]   (ref spectra.py smear_tuv)
// one thread per (row, bin); rows are spectra of a library sharing the grid x
// ---------------------------------------------------------------------------------------------
__global__ void smear_tuv_kernel(int64_t n_rows, int n_x, const double* __restrict__ x, const double* __restrict__ y,
                                 int n_bins, const double* __restrict__ bins, double* out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * n_bins) return;
    const int64_t r = t / n_bins;
    const int i = (int)(t - r * n_bins);
    out[t] = smear_tuv_bin(x, y + r * n_x, n_x, bins[i], bins[i + 1]);
}

cudaError_t launch_smear_tuv(int64_t n_rows, int n_x, const double* x, const double* y, int n_bins, const double* bins,
                             double* out, cudaStream_t stream) {
    const int64_t n = n_rows * n_bins;
    if (n == 0) return cudaSuccess;
    smear_tuv_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n_rows, n_x, x, y, n_bins, bins, out);
    return cudaGetLastError();
}

}  // namespace crt
