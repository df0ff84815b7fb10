#include <cstdio>
#include <cuda_runtime.h>
// throughput probes: FFMA reg, FFMA imm, FFMA2 reg
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 aa = make_float2(a, a * 1.01f), bb = make_float2(b, b * 0.99f);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = fmaf(x[i].y, aa.y, bb.y); }
                else if (MODE == 1) { x[i].x = fmaf(x[i].x, 0.999f, x[(i + 1) & 7].y); x[i].y = fmaf(x[i].y, 1.001f, x[(i + 1) & 7].x); }
                else if (MODE == 2) { x[i] = __ffma2_rn(x[i], aa, bb); }
                else if (MODE == 3) { x[i] = __ffma2_rn(x[(i + 3) & 7], make_float2(0.999f, 1.001f), x[i]); }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    double fma = 2.0 * 8 * 8 * (double)iters * 148 * 8 * 256;   // scalar FMAs
    printf("%s: %.3f ms  %.2f TFLOP/s\n", name, best, 2 * fma / best / 1e9);
}
int main() { run<0>("ffma reg"); run<1>("ffma imm"); run<2>("ffma2 reg"); run<3>("ffma2 const"); return 0; }
