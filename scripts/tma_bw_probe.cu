// Probe: per-SM TMA throughput for the tile box, strided-z (elementStrides=4) vs contiguous rows.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int mode, int ntx, int iters, int box_bytes, unsigned* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), ss = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    uint32_t parity = 0; unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        const int t = (blockIdx.x + it * gridDim.x) % (ntx * ntx * ntx);
        const int x0 = (t % ntx) * 32, y0 = ((t / ntx) % ntx) * 32, z0 = (t / (ntx * ntx)) * 32;
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sb), "r"(4 * box_bytes));
            for (int r = 0; r < 4; r++) {
                const uint32_t dst = ss + r * 54272;
                if (mode == 0)      // canonical layout [x][y][z][4], strided samples along z
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
                                 :: "r"(dst), "l"(&tm), "r"(0), "r"(z0 + r), "r"(y0), "r"(x0), "r"(sb) : "memory");
                else if (mode == 2) // de-interleaved layout, dim0 = whole 112-byte row of 7 float4
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
                                 :: "r"(dst), "l"(&tm), "r"(z0), "r"(r), "r"(y0), "r"(x0), "r"(sb) : "memory");
                else                // de-interleaved layout [x][y][r][q][4]: 5-D, contiguous 7-sample rows
                    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];"
                                 :: "r"(dst), "l"(&tm), "r"(0), "r"(z0 / 4), "r"(r), "r"(y0), "r"(x0), "r"(sb) : "memory");
            }
        }
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(sb), "r"(parity));
        parity ^= 1;
        acc += ((unsigned*)smem)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 0x12345) sink[0] = acc;
}

int main()
{
    const int N = 512, P = N + 8;            // padded cube
    const size_t cells = (size_t)P * P * P;
    float4* d; cudaMalloc(&d, cells * 16); cudaMemset(d, 0, cells * 16);
    unsigned* sink; cudaMalloc(&sink, 4);
    EncodeFn enc; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    CUtensorMap tm0, tm1, tm2;
    {
        const int Q = P / 4;
        cuuint64_t gdim[4] = {(cuuint64_t)Q * 4, 4, (cuuint64_t)P, (cuuint64_t)P};
        cuuint64_t gstr[3] = {(cuuint64_t)Q * 16, (cuuint64_t)P * 16, (cuuint64_t)P * P * 16};
        cuuint32_t box[4] = {28, 1, 22, 22}, estr[4] = {1, 1, 1, 1};
        printf("enc2 %d\n", (int)enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    }
    {
        cuuint64_t gdim[4] = {4, (cuuint64_t)P, (cuuint64_t)P, (cuuint64_t)P};
        cuuint64_t gstr[3] = {16, (cuuint64_t)P * 16, (cuuint64_t)P * P * 16};
        cuuint32_t box[4] = {4, 28, 22, 22}, estr[4] = {1, 4, 1, 1};
        printf("enc0 %d\n", (int)enc(&tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    }
    {
        const int Q = P / 4;
        cuuint64_t gdim[5] = {4, (cuuint64_t)Q, 4, (cuuint64_t)P, (cuuint64_t)P};
        cuuint64_t gstr[4] = {16, (cuuint64_t)Q * 16, (cuuint64_t)P * 16, (cuuint64_t)P * P * 16};
        cuuint32_t box[5] = {4, 7, 1, 22, 22}, estr[5] = {1, 1, 1, 1, 1};
        printf("enc1 %d\n", (int)enc(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 54272);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 64;
    for (int grid : {1, 148}) for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            probe<<<grid, 128, 4 * 54272>>>(mode == 2 ? tm2 : mode ? tm1 : tm0, mode, 16, iters, 54208, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("grid %3d mode %d (%s): %.3f ms, %.2f us per tile, %.1f GB/s per SM, %.1f GB/s total  [%s]\n", grid, mode, mode == 2 ? "112 B rows as dim0" : mode ? "contiguous rows" : "strided z",
                            ms, 1e3 * ms / iters, 4 * 54208.0 * iters / (ms * 1e-3) / 1e9, grid * 4 * 54208.0 * iters / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
