// fp64peak.cu -- FP64 FMA throughput of the GPU (the secondary roofline of SURVEY.md section 8d: the
// reduced-diagnostic mode issues no profile stores and is bounded by the FP64 pipe, not by HBM).
//   every thread runs ILP independent DFMA chains for ITERS iterations; the grid fills every SM with
//   `warps` warps.  Reports DFMA warp-instructions/clk/SM and TFLOP/s (2 flops per FMA) at the clock the
//   run sustained (elapsed SM cycles from clock64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64peak fp64peak.cu ; run: ./fp64peak
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, long long* cycles, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (double)(threadIdx.x + i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
static void run(int n_sm, int threads, int ctas_per_sm, int iters) {
    const int grid = n_sm * ctas_per_sm;
    double* out;
    long long* cyc;
    CK(cudaMalloc(&out, (size_t)grid * threads * sizeof(double)));
    CK(cudaMalloc(&cyc, (size_t)grid * sizeof(long long)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    dfma_kernel<ILP><<<grid, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        dfma_kernel<ILP><<<grid, threads>>>(out, cyc, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long* h = (long long*)malloc((size_t)grid * sizeof(long long));
    CK(cudaMemcpy(h, cyc, (size_t)grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean_cyc = 0.0;
    for (int i = 0; i < grid; ++i) mean_cyc += (double)h[i];
    mean_cyc /= grid;
    const double fma_total = (double)grid * threads * (double)iters * ILP;
    const double tflops = 2.0 * fma_total / (best * 1e-3) / 1e12;
    const double per_clk_sm = (double)threads * ctas_per_sm * (double)iters * ILP / mean_cyc;  // thread-FMAs / clk / SM
    printf("ILP %d  %4d thr x %d CTA/SM  %8.3f ms  %6.2f TFLOP/s fp64  %6.1f FMA/clk/SM  (%.0f MHz effective)\n", ILP, threads,
           ctas_per_sm, best, tflops, per_clk_sm, mean_cyc / (best * 1e-3) / 1e6);
    free(h);
    CK(cudaFree(out));
    CK(cudaFree(cyc));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    const int n = p.multiProcessorCount;
    run<1>(n, 512, 1, 1 << 16);   // 16 warps/SM, dependent chain: latency-bound reference point
    run<4>(n, 512, 1, 1 << 15);
    run<8>(n, 512, 1, 1 << 14);
    run<8>(n, 1024, 1, 1 << 14);
    run<8>(n, 1024, 2, 1 << 14);
    run<16>(n, 256, 2, 1 << 13);
    return 0;
}
