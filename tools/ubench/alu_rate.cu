// Micro-benchmark (development): issue rate of candidate compare-exchange instructions on sm_100a.
// Each kernel runs 8 independent chains per thread, 256 threads x 148*4 CTAs; prints
// warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cyc)
{
    uint32_t v[8], w[8];
    for (int i = 0; i < 8; i++) {
        v[i] = seed * (threadIdx.x + 1) + i * 7919u;
        w[i] = seed ^ (i * 104729u + threadIdx.x);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) {          // VIMNMX.U32 pair (compare-exchange)
                uint32_t lo = min(v[i], w[i]), hi = max(v[i], w[i]);
                v[i] = lo + it; w[i] = hi;
            } else if (OP == 1) {   // FMNMX pair
                float a = __uint_as_float(v[i]), b = __uint_as_float(w[i]);
                float lo = fminf(a, b), hi = fmaxf(a, b);
                v[i] = __float_as_uint(lo) + it; w[i] = __float_as_uint(hi);
            } else if (OP == 2) {   // ISETP + 2 SEL
                bool p = v[i] < w[i];
                uint32_t lo = p ? v[i] : w[i], hi = p ? w[i] : v[i];
                v[i] = lo + it; w[i] = hi;
            } else if (OP == 3) {   // LOP3 pair (reference rate of the ALU pipe)
                uint32_t lo = v[i] & w[i], hi = v[i] | w[i];
                v[i] = lo + it; w[i] = hi;
            } else if (OP == 4) {   // signed VIMNMX
                int lo = min((int)v[i], (int)w[i]), hi = max((int)v[i], (int)w[i]);
                v[i] = (uint32_t)lo + it; w[i] = (uint32_t)hi;
            } else if (OP == 5) {   // u16x2 min/max
                uint32_t lo = __vminu2(v[i], w[i]), hi = __vmaxu2(v[i], w[i]);
                v[i] = lo + it; w[i] = hi;
            } else if (OP == 6) {   // min + (a + b - min) : one VIMNMX, two adds
                uint32_t lo = min(v[i], w[i]);
                uint32_t hi = v[i] + w[i] - lo;
                v[i] = lo + it; w[i] = hi;
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += v[i] ^ w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int opsPerIter)
{
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMalloc(&cyc, 8);
    k<OP><<<148 * 4, 256>>>(out, 12345u, cyc);     // 4 CTAs x 8 warps = 32 warps / SM = 8 / SMSP
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++)
        k<OP><<<148 * 4, 256>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // warp-instructions of interest issued per SMSP: 8 warps * ITERS * 8 chains * opsPerIter
    double inst = 8.0 * ITERS * 8 * opsPerIter;
    printf("%-28s clock64 %8lld  event %.3f ms = %.0f cycles at 1.965 GHz -> %.3f cycles per counted warp-instruction per SMSP\n",
           name, h, ms, ms * 1.965e6, ms * 1.965e6 / inst);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<3>("LOP3 x2 + IADD", 3);
    run<0>("VIMNMX.U32 x2 + IADD", 3);
    run<4>("VIMNMX.S32 x2 + IADD", 3);
    run<1>("FMNMX x2 + IADD", 3);
    run<2>("ISETP + SEL x2 + IADD", 4);
    run<5>("VMIN/VMAX u16x2 + IADD", 3);
    run<6>("VIMNMX + 2 adds + IADD", 4);
    return 0;
}
