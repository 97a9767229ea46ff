// Probe (not part of the product): how close can a pure DFMA stream get to the nominal FP64 rate?
// Variants differ in operand register sharing (register-bank pressure) and chain count.
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256) probe(double* sink, int iters, double a, double b) {
    double x[8], aa[8], bb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-9 + k; aa[k] = a + k * 1e-12; bb[k] = b + k * 1e-13; }
#pragma unroll 16
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (V == 0) x[k] = fma(x[k], a, b);                 // shared a, b
            if (V == 1) x[k] = fma(x[k], aa[k], bb[k]);         // private a_k, b_k
            if (V == 2) x[k] = fma(x[k], aa[k], x[k]);          // x*a + x: two distinct register operands
            if (V == 3) x[k] = fma(x[k], 0.999999, 1e-9);       // immediates / constant bank
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 123.456) sink[0] = s;
}

template <int V>
void run(int iters) {
    double* sink; cudaMalloc(&sink, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<V><<<grid, 256>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    cudaFree(sink);
    printf("variant %d: %.3f ms  %.2f TFLOP/s\n", V, best, 2.0 * 8 * iters * (double)grid * 256 / best / 1e9);
}

int main() { run<0>(1 << 16); run<1>(1 << 16); run<2>(1 << 16); run<3>(1 << 16); return 0; }
