// wbw2.cu -- round-2 store-pattern microbenchmarks (no arithmetic), two questions:
//  (1) column-per-thread kernels (the tridiagonal schemes: a thread owns VEC adjacent bands and walks all n_z levels,
//      NF fields per level): which CTA shape / residency / cluster shape gets the best HBM write bandwidth?
//  (2) the row-sweep kernels' item pattern: per-thread 16-byte st.global.cs vs staging in shared memory and
//      cp.async.bulk (TMA bulk store, 512 B per warp and field), same order of rows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw2 wbw2.cu ; run: ./wbw2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Fields { double* f[7]; };

__device__ __forceinline__ void st16(double* p, double a, double b) { __stcs(reinterpret_cast<double2*>(p), make_double2(a, b)); }

// column-per-thread: CTA = (scenario, tile); thread = 2 adjacent bands, loops over levels, NF fields per level
template <int NF>
__global__ void column_kernel(Fields F, int n_z, int n_wl, int tiles) {
    extern __shared__ double dummy[];
    const size_t s = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const int b0 = (t * blockDim.x + threadIdx.x) * 2;
    if (b0 >= n_wl) return;
    const size_t base = s * (size_t)n_z * n_wl + b0;
    double v = (double)b0;
    for (int j = 0; j < n_z; ++j) {
        const size_t o = base + (size_t)j * n_wl;
#pragma unroll
        for (int f = 0; f < NF; ++f) st16(F.f[f] + o, v + f, v + 1);
        v += 0.5;
    }
}

// same, but the CTAs of one scenario advance level by level together (cluster barrier every LB levels)
template <int NF>
__global__ void column_cluster_kernel(Fields F, int n_z, int n_wl, int tiles, int LB) {
    extern __shared__ double dummy[];
    const size_t s = blockIdx.x / tiles;
    const int t = blockIdx.x % tiles;
    const int b0 = (t * blockDim.x + threadIdx.x) * 2;
    const bool live = b0 < n_wl;
    const size_t base = s * (size_t)n_z * n_wl + b0;
    double v = (double)b0;
    for (int j = 0; j < n_z; ++j) {
        if (live) {
            const size_t o = base + (size_t)j * n_wl;
#pragma unroll
            for (int f = 0; f < NF; ++f) st16(F.f[f] + o, v + f, v + 1);
        }
        v += 0.5;
        if (LB > 0 && (j + 1) % LB == 0) {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        }
    }
}

// the row-sweep item pattern (one CTA per scenario, warps pull (LV levels x 64 bands) items), per-thread stores
template <int NF>
__global__ void items_kernel(Fields F, int n_z, int n_wl, int LV) {
    extern __shared__ double dummy[];
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
#pragma unroll
            for (int f = 0; f < NF; ++f) st16(F.f[f] + o, v + f, v + 1);
            v += 0.5;
        }
    }
}

// same order of rows, but each warp stages one level of its item (NF fields x 64 bands = NF x 512 B) in shared
// memory (two buffers) and lanes 0..NF-1 issue one cp.async.bulk shared -> global each
template <int NF>
__global__ void items_bulk_kernel(Fields F, int n_z, int n_wl, int LV) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int counter;
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    const size_t s = blockIdx.x;
    const int n_grp = n_wl / 2, n_chunks = (n_grp + 31) / 32, n_items = n_chunks * ((n_z + LV - 1) / LV);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* stage = reinterpret_cast<double*>(smem) + (size_t)warp * 2 * NF * 64;  // [2][NF][64]
    int buf = 0;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int lg = item / n_chunks, ch = item - lg * n_chunks, g = ch * 32 + lane;
        const int c0 = 2 * g;
        const int cols = min(64, n_wl - ch * 64);  // bands of this chunk
        double v = (double)c0;
        for (int j = lg * LV; j < min(n_z, lg * LV + LV); ++j) {
            // the buffer we are about to fill was handed to the bulk engine two levels ago
            if (lane < NF) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            double* sb = stage + (size_t)buf * NF * 64;
            if (g < n_grp) {
#pragma unroll
                for (int f = 0; f < NF; ++f) *reinterpret_cast<double2*>(sb + f * 64 + 2 * lane) = make_double2(v + f, v + 1);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < NF) {
                double* gdst = F.f[lane] + (s * n_z + j) * (size_t)n_wl + ch * 64;
                const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sb + lane * 64);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(cols * 8) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
            v += 0.5;
        }
    }
    if (lane < NF) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <class F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

template <class K>
void launch_cluster(K kernel, int grid, int block, size_t smem, int cluster, Fields F, int n_z, int n_wl, int tiles, int LB) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kernel, F, n_z, n_wl, tiles, LB));
}

int main(int argc, char** argv) {
    const int S = 2072, n_z = 60, n_wl = 2100;
    const int reps = argc > 1 ? atoi(argv[1]) : 5;
    const size_t per = (size_t)S * n_z * n_wl;
    Fields F;
    for (int i = 0; i < 7; ++i) CK(cudaMalloc(&F.f[i], per * sizeof(double)));
    printf("S = %d scenarios x %d x %d; %d reps\n", S, n_z, n_wl, reps);
#define COL(NF, BLK, SMEM) { int tiles = (n_wl + 2 * BLK - 1) / (2 * BLK); double gb = NF * per * 8.0 / 1e9; \
        CK(cudaFuncSetAttribute(column_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        float ms = time_ms([&] { column_kernel<NF><<<S * tiles, BLK, SMEM>>>(F, n_z, n_wl, tiles); }, reps); \
        printf("column  NF=%d block=%-4d tiles=%-2d smem=%-6d            %8.3f ms  %8.1f GB/s\n", NF, BLK, tiles, SMEM, ms, gb / ms * 1e3); }
    COL(4, 256, 0) COL(4, 256, 86 * 1024) COL(4, 256, 110 * 1024)
    COL(7, 256, 0) COL(7, 256, 86 * 1024) COL(7, 256, 110 * 1024) COL(7, 256, 70 * 1024)
    COL(7, 544, 110 * 1024) COL(7, 544, 200 * 1024) COL(7, 352, 110 * 1024) COL(7, 352, 200 * 1024) COL(7, 352, 70 * 1024)
    COL(7, 1024, 110 * 1024) COL(7, 1024, 200 * 1024) COL(7, 128, 50 * 1024) COL(7, 128, 100 * 1024) COL(7, 192, 70 * 1024)
    COL(7, 384, 110 * 1024) COL(7, 384, 200 * 1024) COL(7, 512, 200 * 1024) COL(7, 512, 110 * 1024)
#define COLC(NF, BLK, SMEM, CL, LB) { int tiles = (n_wl + 2 * BLK - 1) / (2 * BLK); double gb = NF * per * 8.0 / 1e9; \
        if (tiles % CL == 0) { \
        CK(cudaFuncSetAttribute(column_cluster_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        float ms = time_ms([&] { launch_cluster(column_cluster_kernel<NF>, S * tiles, BLK, SMEM, CL, F, n_z, n_wl, tiles, LB); }, reps); \
        printf("column  NF=%d block=%-4d tiles=%-2d smem=%-6d cluster=%d LB=%-2d %8.3f ms  %8.1f GB/s\n", NF, BLK, tiles, SMEM, CL, LB, ms, gb / ms * 1e3); } }
    COLC(7, 256, 86 * 1024, 5, 0) COLC(7, 256, 86 * 1024, 5, 1) COLC(7, 256, 86 * 1024, 5, 5) COLC(7, 256, 86 * 1024, 5, 10)
    COLC(7, 544, 200 * 1024, 2, 0) COLC(7, 544, 200 * 1024, 2, 1) COLC(7, 544, 200 * 1024, 2, 5)
    COLC(7, 352, 200 * 1024, 3, 0) COLC(7, 352, 200 * 1024, 3, 1) COLC(7, 352, 200 * 1024, 3, 5) COLC(7, 352, 110 * 1024, 3, 1) COLC(7, 352, 110 * 1024, 3, 5)
    COLC(4, 256, 86 * 1024, 5, 1) COLC(4, 256, 86 * 1024, 5, 5) COLC(4, 352, 200 * 1024, 3, 1) COLC(4, 352, 110 * 1024, 3, 1)
#define ITEMS(K, NAME, NF, BLK, LVV, SMEM) { double gb = NF * per * 8.0 / 1e9; \
        CK(cudaFuncSetAttribute(K<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        float ms = time_ms([&] { K<NF><<<S, BLK, SMEM>>>(F, n_z, n_wl, LVV); }, reps); \
        printf("%-11s NF=%d block=%-4d LV=%-3d smem=%-6d            %8.3f ms  %8.1f GB/s\n", NAME, NF, BLK, LVV, SMEM, ms, gb / ms * 1e3); }
    ITEMS(items_kernel, "items", 4, 512, 10, 134 * 1024) ITEMS(items_bulk_kernel, "items_bulk", 4, 512, 10, 200 * 1024)
    ITEMS(items_kernel, "items", 4, 512, 6, 134 * 1024) ITEMS(items_bulk_kernel, "items_bulk", 4, 512, 6, 200 * 1024)
    ITEMS(items_kernel, "items", 7, 512, 10, 134 * 1024) ITEMS(items_bulk_kernel, "items_bulk", 7, 512, 10, 200 * 1024)
    ITEMS(items_bulk_kernel, "items_bulk", 4, 1024, 10, 200 * 1024) ITEMS(items_bulk_kernel, "items_bulk", 4, 256, 10, 200 * 1024)
    // sustained (power-capped) comparison: ~2 s of back-to-back launches each
    if (argc > 2) {
        const int long_reps = atoi(argv[2]);
        for (int rep = 0; rep < 2; ++rep) {
            { double gb = 4 * per * 8.0 / 1e9; float ms = time_ms([&] { items_kernel<4><<<S, 512, 134 * 1024>>>(F, n_z, n_wl, 10); }, long_reps);
              printf("SUSTAINED items      NF=4 %8.3f ms  %8.1f GB/s\n", ms, gb / ms * 1e3); }
            { double gb = 4 * per * 8.0 / 1e9; float ms = time_ms([&] { items_bulk_kernel<4><<<S, 512, 200 * 1024>>>(F, n_z, n_wl, 10); }, long_reps);
              printf("SUSTAINED items_bulk NF=4 %8.3f ms  %8.1f GB/s\n", ms, gb / ms * 1e3); }
        }
    }
    return 0;
}
