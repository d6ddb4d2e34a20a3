// Micro-benchmark: sustained FP64 FMA throughput of one B200 -- the denominator of the peeling
// kernels' "fraction of the FP64 pipe" (BASELINE.md section 2 leaves it "to be measured").
// 8 independent DFMA chains per thread, 256 threads, 148 x 8 CTAs; CUDA-event timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_peak dfma_peak.cu && ./dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 8192
__global__ void __launch_bounds__(256) k(double *out, double a, double b)
{
    double v[8];
    for (int i = 0; i < 8; i++)
        v[i] = a * (threadIdx.x + 1) + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            v[i] = fma(v[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; i++)
        s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int ctas = sms * 8;
    double *out;
    cudaMalloc(&out, (size_t)ctas * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++)
        k<<<ctas, 256>>>(out, 1.0000001, 1e-9);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<<<ctas, 256>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best)
            best = ms;
    }
    const double flop = 2.0 * 8 * ITERS * 256.0 * ctas;
    const double tf = flop / (best * 1e-3) / 1e12;
    std::printf("{\"sms\": %d, \"sm_clock_mhz\": %.0f, \"dfma_tflops\": %.2f, "
                "\"dfma_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n",
                sms, clk / 1e3, tf, flop / 2 / (best * 1e-3) / (clk * 1e3) / sms, best);
    return cudaGetLastError() != cudaSuccess;
}
