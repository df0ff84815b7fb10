// Probe: how many shared-memory wavefronts does an LDS.128 / LDS.64 cost when pairs of lanes read the SAME address?
// (design question for the tiled sweep: the two lanes that share a 4-site z segment read overlapping plane windows)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters)
{
    extern __shared__ __align__(16) unsigned char sm[];
    float4 *s4 = reinterpret_cast<float4 *>(sm);
    float2 *s2 = reinterpret_cast<float2 *>(sm);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s4[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, h = (lane >> 2) & 1, j = lane >> 3;
    int idx;
    if (MODE == 0 || MODE == 3) idx = lane;                       // all distinct, consecutive
    else if (MODE == 1 || MODE == 4) idx = lane >> 1;             // adjacent lane pairs share an address
    else if (MODE == 2 || MODE == 5) idx = 2 * kk + 56 * j;       // the tile pattern: lanes h = 0 / 1 share
    else idx = 2 * kk + h + 56 * j;                               // (6, 7) today's pattern: all distinct
    idx += warp * 300;
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE <= 2 || MODE == 6) { const float4 v = s4[idx + u * 14]; acc += v.x + v.y + v.z + v.w; }
            else { const float2 v = s2[idx + u * 14]; acc += v.x + v.y; }
        }
        idx ^= (it & 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char *name)
{
    float *out; cudaMalloc(&out, 148 * 256 * 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<MODE><<<148, 256, 8192 * 16>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    const double warp_instr = 8.0 * 16 * iters;                    // per SM
    printf("%-44s %.3f ms  %.2f cycles per warp-LDS (1.965 GHz)\n", name, best, best * 1e-3 * 1.965e9 / warp_instr);
}
int main()
{
    run<0>("LDS.128 distinct consecutive");
    run<1>("LDS.128 lane pairs share (consecutive)");
    run<2>("LDS.128 tile pattern, h lanes share");
    run<6>("LDS.128 tile pattern, all distinct (v11)");
    run<3>("LDS.64 distinct consecutive");
    run<4>("LDS.64 lane pairs share (consecutive)");
    run<5>("LDS.64 tile pattern, h lanes share");
    run<7>("LDS.64 tile pattern, all distinct (v11)");
    return 0;
}
