// What the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4) sustain on this GPU, with nothing else in the loop:
// NACC independent accumulator tiles per warp, W warps per CTA, CTAs per SM chosen by the launch.  Prints TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/dmma_peak scratch/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_loop(int iters, double* out, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int t = 0; t < NACC; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
    double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < NACC; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NACC; ++t) s += c[t][0] + c[t][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
static void run(int warps_per_cta, int ctas_per_sm, int sms, double* out) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_loop<NACC><<<sms * ctas_per_sm, 32 * warps_per_cta>>>(100, out, 1.0, 2.0);
    cudaEventRecord(e0);
    dmma_loop<NACC><<<sms * ctas_per_sm, 32 * warps_per_cta>>>(iters, out, 1.0, 2.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 256.0 * NACC * (double)iters * warps_per_cta * ctas_per_sm * sms;
    printf("NACC %2d  warps/SM %2d : %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", NACC, warps_per_cta * ctas_per_sm, ms,
           flop / ms * 1e-9, flop / 2.0 / (ms * 1e-3) / sms / 1.9e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* out;
    cudaMalloc(&out, 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<15>(4, 1, sms, out);
    run<15>(8, 1, sms, out);
    run<15>(8, 2, sms, out);
    run<15>(8, 4, sms, out);
    run<4>(8, 2, sms, out);
    run<1>(8, 4, sms, out);
    run<1>(8, 8, sms, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
