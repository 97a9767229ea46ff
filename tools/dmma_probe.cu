// Probe (not part of the product): does FP64 mma.sync (DMMA m8n8k4) on B200 run on a unit separate
// from the FP64 FMA pipe?  Times DFMA-only, DMMA-only and an interleaved mix on all SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>   // 0 dfma, 1 dmma, 2 both
__global__ void __launch_bounds__(256) probe(double* sink, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        if (MODE != 1) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
        if (MODE != 0) {
            dmma(c0, c1, a, b); dmma(c2, c3, a, b); dmma(c4, c5, a, b); dmma(c6, c7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    if (s == 123.456) sink[0] = s;
}

template <int MODE>
double run(int iters) {
    double* sink; cudaMalloc(&sink, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, 256>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    cudaFree(sink);
    const double threads = (double)grid * 256;
    const double dfma_flops = MODE != 1 ? 2.0 * 8 * iters * threads : 0;
    const double dmma_flops = MODE != 0 ? 4.0 * (2.0 * 8 * 8 * 4) * iters * (threads / 32) : 0;   // 4 mma per warp-iteration
    printf("mode %d: %.3f ms  dfma %.2f TF  dmma %.2f TF  total %.2f TF\n", MODE, best, dfma_flops / best / 1e9,
           dmma_flops / best / 1e9, (dfma_flops + dmma_flops) / best / 1e9);
    return best;
}

int main() {
    run<0>(1 << 16); run<1>(1 << 16); run<2>(1 << 16);
    return 0;
}
