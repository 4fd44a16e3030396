// wbw.cu -- write-bandwidth microbenchmarks that bracket the solver kernel's HBM roofline.
//   linear : grid-stride 16-byte streaming stores over one buffer (pure write ceiling)
//   pattern: the solver's store pattern with no arithmetic -- CTA = (scenario, band tile), thread = 2
//            adjacent bands, loop over n_z levels, 4 fields, 16-byte stores at stride n_wl*8 bytes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw wbw.cu ; run: ./wbw
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int POLICY>
__device__ __forceinline__ void st16(double* p, double a, double b) {
    if (POLICY == 0) __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
    else if (POLICY == 1) *reinterpret_cast<double2*>(p) = make_double2(a, b);
    else if (POLICY == 2) __stcg(reinterpret_cast<double2*>(p), make_double2(a, b));
    else __stwt(reinterpret_cast<double2*>(p), make_double2(a, b));
}

template <int POLICY>
__global__ void linear_kernel(double* buf, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) st16<POLICY>(buf + 2 * i, (double)i, 1.0);
}

template <int POLICY>
__global__ void pattern_kernel(double* f0, double* f1, double* f2, double* f3, int n_z, int n_wl, int tiles) {
    const size_t s = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const int b0 = (t * blockDim.x + threadIdx.x) * 2;
    if (b0 >= n_wl) return;
    const size_t base = s * (size_t)n_z * n_wl + b0;
    double v = (double)b0;
    for (int j = 0; j < n_z; ++j) {
        const size_t o = base + (size_t)j * n_wl;
        st16<POLICY>(f0 + o, v, v + 1);
        st16<POLICY>(f1 + o, v + 2, v + 3);
        st16<POLICY>(f2 + o, v + 4, v + 5);
        st16<POLICY>(f3 + o, v + 6, v + 7);
        v += 0.5;
    }
}

// level-major variant: CTA = (scenario, level group), threads sweep the whole band row (long contiguous runs)
template <int POLICY>
__global__ void rowmajor_kernel(double* f0, double* f1, double* f2, double* f3, int n_z, int n_wl, int lev_per_cta) {
    const int groups = (n_z + lev_per_cta - 1) / lev_per_cta;
    const size_t s = blockIdx.x / groups;
    const int g = blockIdx.x % groups;
    for (int j = g * lev_per_cta; j < min(n_z, (g + 1) * lev_per_cta); ++j) {
        const size_t row = (s * (size_t)n_z + j) * n_wl;
        for (int b0 = threadIdx.x * 2; b0 < n_wl; b0 += blockDim.x * 2) {
            double v = (double)b0;
            st16<POLICY>(f0 + row + b0, v, v + 1);
            st16<POLICY>(f1 + row + b0, v + 2, v + 3);
            st16<POLICY>(f2 + row + b0, v + 4, v + 5);
            st16<POLICY>(f3 + row + b0, v + 6, v + 7);
        }
    }
}

// the row-sweep kernel's actual store order: one CTA per scenario, warps pull (LV levels x 64 bands) items
// from a shared counter in row-major order
template <int POLICY>
__global__ void items_kernel(double* f0, double* f1, double* f2, double* f3, int n_z, int n_wl, int LV) {
    __shared__ int counter;
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    const size_t s = blockIdx.x;
    const int n_grp = n_wl / 2, n_chunks = (n_grp + 31) / 32, n_items = n_chunks * ((n_z + LV - 1) / LV);
    const int lane = threadIdx.x & 31;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int lg = item / n_chunks, g = (item - lg * n_chunks) * 32 + lane;
        if (g >= n_grp) continue;
        const int c0 = 2 * g;
        double v = (double)c0;
        for (int j = lg * LV; j < min(n_z, lg * LV + LV); ++j) {
            const size_t o = (s * n_z + j) * (size_t)n_wl + c0;
            st16<POLICY>(f0 + o, v, v + 1);
            st16<POLICY>(f1 + o, v + 2, v + 3);
            st16<POLICY>(f2 + o, v + 4, v + 5);
            st16<POLICY>(f3 + o, v + 6, v + 7);
            v += 0.5;
        }
    }
}

template <class F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    const int S = 4096, n_z = 60, n_wl = 2100;
    const size_t per = (size_t)S * n_z * n_wl;  // doubles per field
    double* f[4];
    for (int i = 0; i < 4; ++i) CK(cudaMalloc(&f[i], per * sizeof(double)));
    const double gb = 4.0 * per * 8 / 1e9;
    const char* pol[4] = {"cs", "wb", "cg", "wt"};
    printf("bytes per launch: %.2f GB\n", gb);
#define LIN(P) { float ms = time_ms([&] { for (int i = 0; i < 4; ++i) linear_kernel<P><<<148 * 16, 256>>>(f[i], per / 2); }, 5); \
                 printf("linear   %-3s                     %8.3f ms  %8.1f GB/s\n", pol[P], ms, gb / ms * 1e3); }
    LIN(0) LIN(1) LIN(2) LIN(3)
#define PAT(P, BLK) { int tiles = (n_wl + 2 * BLK - 1) / (2 * BLK); \
                 float ms = time_ms([&] { pattern_kernel<P><<<S * tiles, BLK>>>(f[0], f[1], f[2], f[3], n_z, n_wl, tiles); }, 5); \
                 printf("pattern  %-3s block=%-4d tiles=%-3d %8.3f ms  %8.1f GB/s\n", pol[P], BLK, tiles, ms, gb / ms * 1e3); }
    PAT(0, 64) PAT(0, 128) PAT(0, 256) PAT(0, 512) PAT(0, 1024)
    PAT(1, 128) PAT(1, 256) PAT(2, 128) PAT(2, 256) PAT(3, 256)
#define ROW(P, BLK, LPC) { int groups = (n_z + LPC - 1) / LPC; \
                 float ms = time_ms([&] { rowmajor_kernel<P><<<S * groups, BLK>>>(f[0], f[1], f[2], f[3], n_z, n_wl, LPC); }, 5); \
                 printf("rowmajor %-3s block=%-4d lev/cta=%-3d %8.3f ms  %8.1f GB/s\n", pol[P], BLK, LPC, ms, gb / ms * 1e3); }
    ROW(0, 256, 1) ROW(0, 256, 4) ROW(0, 512, 4) ROW(0, 1024, 6) ROW(1, 256, 4)
    ROW(0, 352, 60) ROW(0, 544, 60) ROW(0, 288, 60) ROW(0, 224, 60) ROW(0, 192, 60) ROW(0, 1024, 60)
    ROW(0, 352, 30) ROW(0, 352, 15) ROW(0, 544, 30) ROW(1, 352, 60) ROW(2, 352, 60)
    ROW(0, 352, 2) ROW(0, 352, 3) ROW(0, 352, 4) ROW(0, 352, 5) ROW(0, 352, 6) ROW(0, 352, 10) ROW(0, 352, 12)
    ROW(0, 544, 4) ROW(0, 544, 6) ROW(0, 544, 10) ROW(0, 1024, 4) ROW(0, 1024, 10) ROW(0, 1024, 12) ROW(0, 1024, 15)
    ROW(0, 128, 4) ROW(0, 128, 6) ROW(0, 192, 6)
    // occupancy-limited variants: dynamic shared memory forces 1 or 2 resident CTAs per SM
#define ROWS(P, BLK, LPC, SMEM) { int groups = (n_z + LPC - 1) / LPC; \
                 CK(cudaFuncSetAttribute(rowmajor_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
                 float ms = time_ms([&] { rowmajor_kernel<P><<<S * groups, BLK, SMEM>>>(f[0], f[1], f[2], f[3], n_z, n_wl, LPC); }, 5); \
                 printf("rowmajor %-3s block=%-4d lev/cta=%-3d smem=%-6d %8.3f ms  %8.1f GB/s\n", pol[P], BLK, LPC, SMEM, ms, gb / ms * 1e3); }
    ROWS(0, 1024, 60, 120 * 1024) ROWS(0, 1024, 60, 60 * 1024) ROWS(0, 512, 60, 120 * 1024) ROWS(0, 512, 60, 60 * 1024)
    ROWS(0, 544, 60, 120 * 1024) ROWS(0, 544, 60, 60*1024) ROWS(0, 352, 60, 120 * 1024) ROWS(0, 1024, 30, 120 * 1024) ROWS(0, 1024, 20, 120 * 1024)
    ROWS(0, 1024, 10, 120 * 1024) ROWS(0, 1024, 6, 120 * 1024)
#define ITEMS(P, BLK, LVV, SMEM) { CK(cudaFuncSetAttribute(items_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
                 float ms = time_ms([&] { items_kernel<P><<<S, BLK, SMEM>>>(f[0], f[1], f[2], f[3], n_z, n_wl, LVV); }, 5); \
                 printf("items    %-3s block=%-4d LV=%-3d smem=%-6d %8.3f ms  %8.1f GB/s\n", pol[P], BLK, LVV, SMEM, ms, gb / ms * 1e3); }
    ITEMS(0, 512, 6, 134 * 1024) ITEMS(0, 512, 1, 134 * 1024) ITEMS(0, 512, 2, 134 * 1024) ITEMS(0, 512, 10, 134 * 1024)
    ITEMS(0, 384, 6, 134 * 1024) ITEMS(0, 1024, 6, 134 * 1024) ITEMS(0, 512, 60, 134 * 1024) ITEMS(0, 256, 6, 134 * 1024)
    return 0;
}
