// Probe: does cp.async.bulk.tensor with elementStrides={1,4,1,1} compact the strided samples in smem?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float4* out, int nq, int ny, int nx, int cz, int cy, int cx)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), ss = (uint32_t)__cvta_generic_to_shared(smem);
    const int n = nq * ny * nx;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sb), "r"(n * 16));
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
                     :: "r"(ss), "l"(&tm), "r"(0), "r"(cz), "r"(cy), "r"(cx), "r"(sb) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(sb), "r"(0));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ((float4*)smem)[i];
}

int main()
{
    const int PX = 10, PY = 9, PZ = 40;
    std::vector<float4> h(PX * PY * PZ);
    for (int x = 0; x < PX; x++) for (int y = 0; y < PY; y++) for (int z = 0; z < PZ; z++) h[(x * PY + y) * PZ + z] = make_float4(x, y, z, 1000 * x + 100 * y + z);
    float4 *d, *o; cudaMalloc(&d, h.size() * 16); cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
    const int nq = 7, ny = 3, nx = 2;
    cudaMalloc(&o, nq * ny * nx * 16); cudaMemset(o, 0xff, nq * ny * nx * 16);
    EncodeFn enc; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    CUtensorMap tm;
    cuuint64_t gdim[4] = {4, (cuuint64_t)PZ, (cuuint64_t)PY, (cuuint64_t)PX};
    cuuint64_t gstr[3] = {16, (cuuint64_t)PZ * 16, (cuuint64_t)PZ * PY * 16};
    cuuint32_t box[4] = {4, 28, (cuuint32_t)ny, (cuuint32_t)nx};
    cuuint32_t estr[4] = {1, 4, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    for (int trial = 0; trial < 2; trial++) {
        int cz = trial == 0 ? 5 : -2, cy = 2, cx = 3;
        probe<<<1, 64, nq * ny * nx * 16>>>(tm, o, nq, ny, nx, cz, cy, cx);
        cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
        std::vector<float4> r_(nq * ny * nx); cudaMemcpy(r_.data(), o, r_.size() * 16, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int x = 0; x < nx; x++) for (int y = 0; y < ny; y++) for (int q = 0; q < nq; q++) {
            float4 v = r_[(x * ny + y) * nq + q]; int z = cz + 4 * q;
            float ex = (z < 0 || z >= PZ) ? 0 : (float)(1000 * (cx + x) + 100 * (cy + y) + z);
            if (v.w != ex) { if (bad < 8) printf("mismatch x%d y%d q%d got (%g %g %g %g) want w=%g\n", x, y, q, v.x, v.y, v.z, v.w, ex); bad++; }
        }
        printf("trial %d (cz=%d): %s (%d mismatches); first row w:", trial, cz, bad ? "BAD" : "OK compacted stride-4 samples", bad);
        for (int q = 0; q < nq; q++) printf(" %g", r_[q].w); printf("\n");
    }
    return 0;
}
