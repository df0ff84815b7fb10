// sn_sweep_tiled.cuh -- the fast Metropolis sweep: TMA-staged shared-memory tiles.
//
// Replaces MC_moves -> MC_move -> site_energy (montecarlo-core.c:76-191) for
// DipoleCutOff = 3 lattices whose X, Y and slab height are multiples of 32.
//
// Decomposition
//   * The lattice is cut into 16^3 tiles.  A sweep is 8 launches ("phases"), one
//     per tile parity (px,py,pz): active tiles are 32 apart, so the 22^3 read set
//     of one active tile never meets the 16^3 write set of another.
//   * One persistent CTA per SM walks the phase's tiles.  Per tile, one thread
//     issues four cp.async.bulk.tensor (TMA) loads from the padded float4 lattice:
//     box 22 x 22 x 28(z) with elementStrides = 4 along z, start shifted by the
//     residue r = 0..3.  Shared memory therefore holds the tile + halo
//     de-interleaved in z: box r keeps planes z0-4+r, z0+r, ..., 7 per (x,y).  Lanes
//     that own sites 4 apart in z (same colour) then read consecutive float4 --
//     LDS.128 without bank conflicts -- and every neighbour address is
//     base + compile-time immediate.
//   * Inside a tile the 64 site colours are visited as 16 super-passes (cx,cy).
//     A thread owns a segment of 4 consecutive z sites of one (x,y) column, i.e.
//     the four colours (cx,cy,0..3).  None of the 28 neighbour columns around it
//     changes during the super-pass (they belong to other (cx,cy) classes), so
//     their contribution to the local fields of all 4 sites is gathered in one
//     go with a sliding z window: 8 loads serve 4 x 5 neighbours (2.2x fewer
//     shared-memory reads than 4 independent gathers), all 798 tensor FFMAs per
//     attempt are still executed.  Only the centre column changes: the 4 sites are
//     then decided in sequence, the field of the later ones corrected in registers
//     for the earlier accepted moves (own thread, and the segment above via
//     shuffle).  Every attempt is a full fresh dE over the cut-off sphere.
//   * Warp specialisation: warps 0-1 gather half of the neighbour columns (and
//     the centre column) and run the sequential chain; warps 2-3 gather the other
//     half, hand their partial fields over through shared memory, and meanwhile
//     draw the Philox proposals for the next super-pass.
//   * Accepted moves are written to the shared tile and straight to global
//     memory together with their ghost images (periodic faces, and the
//     neighbouring GPU's ghost planes over NVLink for a Z-slab handle).
#pragma once

#include <cuda.h>

#include "sn_field.cuh"
#include "sn_sweep_colour.cuh"

namespace snt {
constexpr int T = 16;                    // tile edge
constexpr int H = 3;                     // halo = cut-off
constexpr int BX = T + 2 * H;            // 22 columns per axis in the box
constexpr int NQ = 7;                    // z samples per residue box (window of 28 planes)
constexpr int BOX_F4 = BX * BX * NQ;     // float4 per residue box
constexpr int BOX_BYTES = BOX_F4 * 16;   // 54208 bytes moved by each TMA
constexpr int BOX_STRIDE_F4 = 3392;      // 54272 B: box pitch rounded up to 128 B (TMA destination alignment)
constexpr int SEGS = 64;                 // segments (threads of one warp group) per super-pass
constexpr int OFF_XF = 4 * BOX_STRIDE_F4 * 16;             // partial fields: float4[6][64]
constexpr int OFF_XP = OFF_XF + 6 * SEGS * 16;             // proposals: float4[2][4][64]
constexpr int OFF_BAR = OFF_XP + 2 * 4 * SEGS * 16;        // mbarrier
constexpr int SMEM_BYTES = OFF_BAR + 16;
constexpr int THREADS = 128;

__host__ __device__ constexpr int half_height(int r2xy) { return 9 - r2xy >= 9 ? 3 : 9 - r2xy >= 4 ? 2 : 9 - r2xy >= 1 ? 1 : 0; }
__host__ __device__ constexpr int residue(int e) { return ((e % 4) + 4) % 4; }
__host__ __device__ constexpr int qshift(int e) { return (e - residue(e)) / 4; }      // floor(e / 4)
}  // namespace snt

__device__ __forceinline__ uint32_t sn_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sn_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// Gather the contribution of this warp group's neighbour columns to the local
// fields of the 4 sites of a segment.  `tile` = shared float4 array, B = float4
// index of (column lx,ly; q = k+1) in residue box 0.
template <int GROUP, bool SPECIES>
__device__ __forceinline__ void sn_tile_gather(const float4 *__restrict__ tile, const int B, float3 (&F)[4], float3 (&G)[4],
                                               float4 (&old)[4])
{
    sn_static_for<-3, 4>([&](auto dxc) {
        sn_static_for<-3, 4>([&](auto dyc) {
            constexpr int DX = decltype(dxc)::value, DY = decltype(dyc)::value;
            constexpr int r2xy = DX * DX + DY * DY;
            if constexpr (r2xy <= 9) {
                constexpr bool centre = (DX == 0 && DY == 0);
                constexpr bool upper = DX > 0 || (DX == 0 && DY > 0);
                constexpr bool mine = centre ? GROUP == 0 : (upper ? GROUP == 0 : GROUP == 1);
                if constexpr (mine) {
                    constexpr int M = snt::half_height(r2xy);
                    sn_static_for<-M, 4 + M>([&](auto ec) {
                        constexpr int E = decltype(ec)::value;
                        const float4 w = tile[snt::residue(E) * snt::BOX_STRIDE_F4 + B + (DX * snt::BX + DY) * snt::NQ + snt::qshift(E)];
                        if constexpr (centre && E >= 0 && E <= 3) old[E] = w;
                        sn_static_for<0, 4>([&](auto sc) {
                            constexpr int S = decltype(sc)::value, DZ = E - S;
                            if constexpr (DZ * DZ <= 9 - r2xy && !(centre && DZ == 0))
                                sn_accumulate<DX, DY, DZ, SPECIES>(F[S], G[S], w);
                        });
                    });
                }
            }
        });
    });
}

struct SnTilePhase {
    int px, py, pz;             // tile parity of this launch
    int hx, hy, hz;             // number of active tiles per axis (= tiles / 2)
    int nrep;
};

template <bool SPECIES>
__global__ void __launch_bounds__(snt::THREADS, 1)
sn_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const SnSweepArgs a, const SnTilePhase ph)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *tile = reinterpret_cast<float4 *>(smem);
    float4 *xF = reinterpret_cast<float4 *>(smem + snt::OFF_XF);
    float4 *xP = reinterpret_cast<float4 *>(smem + snt::OFF_XP);
    const uint32_t bar = sn_smem_u32(smem + snt::OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = warp >> 1;                         // 0: gather + chain, 1: gather + proposals
    const int k = lane & 3, j = (lane >> 2) & 3, i = ((warp & 1) << 1) | (lane >> 4);
    const int seg = ((warp & 1) << 5) | lane;            // 0..63 inside the group

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (long long)ph.hx * ph.hy * ph.hz * ph.nrep;
    uint32_t parity = 0;
    const SnGeom &G = a.G;

    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int iz = (int)(t % ph.hz), iy = (int)((t / ph.hz) % ph.hy), ix = (int)((t / ((long long)ph.hz * ph.hy)) % ph.hx);
        const int rep = (int)(t / ((long long)ph.hz * ph.hy * ph.hx));
        const int x0 = (2 * ix + ph.px) * snt::T, y0 = (2 * iy + ph.py) * snt::T, z0 = (2 * iz + ph.pz) * snt::T;

        if (tid == 0) {
            // shared memory was last touched through the generic proxy; order it before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4 * snt::BOX_BYTES) : "memory");
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t dst = sn_smem_u32(smem) + r * snt::BOX_STRIDE_F4 * 16;
                // padded coordinates: x0-3 -> x0, y0-3 -> y0, window start z0-4+r -> z0-1+r (ghost width 3)
                asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(dst), "l"(&tmap), "r"(0), "r"(z0 - 1 + r), "r"(y0), "r"(x0), "r"(rep), "r"(bar) : "memory");
            }
        }

        SnTerms tm;
        tm.cage = a.cage; tm.K = a.K; tm.beta = a.beta[rep];
        { const float4 E = a.efield[rep]; tm.E = make_float3(E.x, E.y, E.z); }
        tm.constrain = a.constrain; tm.dim = a.dim;
        float4 *glat = a.lat + (long long)rep * G.rep_stride;
        float4 *plo = a.peer_lo ? a.peer_lo + (long long)rep * G.rep_stride : nullptr;
        float4 *phi = a.peer_hi ? a.peer_hi + (long long)rep * G.rep_stride : nullptr;

        // trial orientations + accept uniforms for the 4 sites of this thread's segment in super-pass sp
        auto draw = [&](int sp) {
            const int cx = sp >> 2, cy = sp & 3;
            const int x = x0 + cx + 4 * i, y = y0 + cy + 4 * j;
            float4 *dst = xP + (sp & 1) * 4 * snt::SEGS + seg;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int z = z0 + 4 * k + s;
                const unsigned long long gsite = ((unsigned long long)x * G.Y + y) * G.Z + (G.z0 + z);
                const Philox4 r = sn_philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32) ^ ((uint32_t)rep << 8),
                                                   a.sweep_lo, a.sweep_hi, a.key0, a.key1);
                const float3 np = sn_propose(tm, sn_u01(r.x), sn_u01(r.y));
                dst[s * snt::SEGS] = make_float4(np.x, np.y, np.z, sn_u01(r.z));
            }
        };
        if (group == 1) draw(0);                        // overlaps the TMA flight

        sn_mbar_wait(bar, parity);
        parity ^= 1;
        __syncthreads();

        int n_acc = 0, n_rej = 0, n_vac = 0;
#pragma unroll 1
        for (int sp = 0; sp < 16; sp++) {
            const int cx = sp >> 2, cy = sp & 3;
            const int lx = snt::H + cx + 4 * i, ly = snt::H + cy + 4 * j;
            const int B = (lx * snt::BX + ly) * snt::NQ + k + 1;
            float3 F[4], Gc[4];
            float4 old[4];
#pragma unroll
            for (int s = 0; s < 4; s++) { F[s] = make_float3(0.f, 0.f, 0.f); Gc[s] = make_float3(0.f, 0.f, 0.f); old[s] = make_float4(0.f, 0.f, 0.f, 0.f); }

            if (group == 0) {
                sn_tile_gather<0, SPECIES>(tile, B, F, Gc, old);
            } else {
                sn_tile_gather<1, SPECIES>(tile, B, F, Gc, old);
                xF[0 * snt::SEGS + seg] = make_float4(F[0].x, F[0].y, F[0].z, F[1].x);
                xF[1 * snt::SEGS + seg] = make_float4(F[1].y, F[1].z, F[2].x, F[2].y);
                xF[2 * snt::SEGS + seg] = make_float4(F[2].z, F[3].x, F[3].y, F[3].z);
                xF[3 * snt::SEGS + seg] = make_float4(Gc[0].x, Gc[0].y, Gc[0].z, Gc[1].x);
                xF[4 * snt::SEGS + seg] = make_float4(Gc[1].y, Gc[1].z, Gc[2].x, Gc[2].y);
                xF[5 * snt::SEGS + seg] = make_float4(Gc[2].z, Gc[3].x, Gc[3].y, Gc[3].z);
            }
            __syncthreads();

            if (group == 0) {
                {   // add the other group's partial fields
                    const float4 v0 = xF[0 * snt::SEGS + seg], v1 = xF[1 * snt::SEGS + seg], v2 = xF[2 * snt::SEGS + seg];
                    const float4 v3 = xF[3 * snt::SEGS + seg], v4 = xF[4 * snt::SEGS + seg], v5 = xF[5 * snt::SEGS + seg];
                    F[0].x += v0.x; F[0].y += v0.y; F[0].z += v0.z; F[1].x += v0.w;
                    F[1].y += v1.x; F[1].z += v1.y; F[2].x += v1.z; F[2].y += v1.w;
                    F[2].z += v2.x; F[3].x += v2.y; F[3].y += v2.z; F[3].z += v2.w;
                    Gc[0].x += v3.x; Gc[0].y += v3.y; Gc[0].z += v3.z; Gc[1].x += v3.w;
                    Gc[1].y += v4.x; Gc[1].z += v4.y; Gc[2].x += v4.z; Gc[2].y += v4.w;
                    Gc[2].z += v5.x; Gc[3].x += v5.y; Gc[3].y += v5.z; Gc[3].z += v5.w;
                }
                const float4 *prop = xP + (sp & 1) * 4 * snt::SEGS + seg;
                float3 dp[4], dm[4], dmu[4], dpu0 = make_float3(0.f, 0.f, 0.f);
                const int gx = x0 + cx + 4 * i, gy = y0 + cy + 4 * j;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    float3 Fs = F[s], Gs = Gc[s];
                    // earlier sites of this segment, dz = s2 - s in {-1,-2,-3}: T(0,0,dz) = diag(1,1,-2)/|dz|^3
#pragma unroll
                    for (int s2 = 0; s2 < s; s2++) {
                        const int d = s - s2;
                        const float w3 = d == 1 ? 1.0f : d == 2 ? 0.125f : (1.0f / 27.0f);
                        Fs.x = fmaf(w3, dm[s2].x, Fs.x); Fs.y = fmaf(w3, dm[s2].y, Fs.y); Fs.z = fmaf(-2.0f * w3, dm[s2].z, Fs.z);
                        if (d == 1) { Gs.x += dp[s2].x; Gs.y += dp[s2].y; Gs.z += dp[s2].z; }
                    }
                    // earlier sites of the segment above (lane+1), dz = 4 + s2 - s in {1,2,3}
#pragma unroll
                    for (int s2 = 0; s2 < s; s2++) {
                        const int d = 4 + s2 - s;
                        const float w3 = d == 1 ? 1.0f : d == 2 ? 0.125f : (1.0f / 27.0f);
                        Fs.x = fmaf(w3, dmu[s2].x, Fs.x); Fs.y = fmaf(w3, dmu[s2].y, Fs.y); Fs.z = fmaf(-2.0f * w3, dmu[s2].z, Fs.z);
                        if (d == 1) { Gs.x += dpu0.x; Gs.y += dpu0.y; Gs.z += dpu0.z; }   // only s = 3, s2 = 0
                    }
                    const float4 o = old[s];
                    const float4 pr = prop[s * snt::SEGS];
                    const float3 np = make_float3(pr.x, pr.y, pr.z);
                    const bool vacant = o.w == 0.0f;                                   // montecarlo-core.c:163
                    const float dE = sn_delta_e(o, np, Fs, Gs, tm);
                    const bool acc = !vacant && sn_accept(dE, tm.beta, pr.w);          // montecarlo-core.c:179
                    dp[s] = acc ? make_float3(np.x - o.x, np.y - o.y, np.z - o.z) : make_float3(0.f, 0.f, 0.f);
                    dm[s] = SPECIES ? make_float3(o.w * dp[s].x, o.w * dp[s].y, o.w * dp[s].z) : dp[s];
                    if (s < 3) {
                        dmu[s].x = __shfl_down_sync(0xffffffffu, dm[s].x, 1);
                        dmu[s].y = __shfl_down_sync(0xffffffffu, dm[s].y, 1);
                        dmu[s].z = __shfl_down_sync(0xffffffffu, dm[s].z, 1);
                        if (k == 3) dmu[s] = make_float3(0.f, 0.f, 0.f);              // the segment above lies in the (static) halo
                        if (SPECIES && s == 0) {
                            dpu0.x = __shfl_down_sync(0xffffffffu, dp[0].x, 1);
                            dpu0.y = __shfl_down_sync(0xffffffffu, dp[0].y, 1);
                            dpu0.z = __shfl_down_sync(0xffffffffu, dp[0].z, 1);
                            if (k == 3) dpu0 = make_float3(0.f, 0.f, 0.f);
                        } else if (!SPECIES && s == 0) dpu0 = dmu[0];
                    }
                    if (acc) {
                        const float4 nv = make_float4(np.x, np.y, np.z, o.w);
                        tile[snt::residue(s) * snt::BOX_STRIDE_F4 + B + snt::qshift(s)] = nv;
                        sn_store_site(glat, plo, phi, G, gx, gy, z0 + 4 * k + s, nv);
                    }
                    n_acc += acc; n_rej += (!acc && !vacant); n_vac += vacant;
                }
            } else if (sp < 15) {
                draw(sp + 1);
            }
            __syncthreads();
        }
        if (group == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
                n_rej += __shfl_xor_sync(0xffffffffu, n_rej, o);
                n_vac += __shfl_xor_sync(0xffffffffu, n_vac, o);
            }
            if (lane == 0) {
                unsigned long long *c = a.counters + 3 * rep;
                if (n_acc) atomicAdd(c + 0, (unsigned long long)n_acc);
                if (n_rej) atomicAdd(c + 1, (unsigned long long)n_rej);
                if (n_vac) atomicAdd(c + 2, (unsigned long long)n_vac);
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------
typedef CUresult (*SnEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool sn_tiled_supported(const sn_handle *h, std::string *why)
{
    const SnGeom &G = h->G;
    const char *msg = nullptr;
    if (h->p.cutoff != 3) msg = "DipoleCutOff must be 3";
    else if (G.Z == 1) msg = "lattice is flat (Z == 1)";
    else if (G.X % 32 || G.Y % 32 || G.nz % 32 || G.z0 % 32) msg = "X, Y, slab height and slab origin must be multiples of 32";
    if (msg) { if (why) *why = msg; return false; }
    return true;
}

int sn_tiled_prepare(sn_handle *h)
{
    const SnGeom &G = h->G;
    SnEncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    SN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr));
    if (!enc || qr != cudaDriverEntryPointSuccess) return sn_fail(SN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap *tm = new CUtensorMap;
    const cuuint64_t gdim[5] = {4, (cuuint64_t)G.PZ, (cuuint64_t)G.PY, (cuuint64_t)(G.X + 2 * G.g), (cuuint64_t)h->p.nreplicas};
    const cuuint64_t gstr[4] = {16, (cuuint64_t)G.PZ * 16, (cuuint64_t)G.sx * 16, (cuuint64_t)G.rep_stride * 16};
    const cuuint32_t box[5] = {4, 4 * snt::NQ, snt::BX, snt::BX, 1};
    const cuuint32_t estr[5] = {1, 4, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, h->lat, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete tm; return sn_fail(SN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r); }
    h->tmap = tm;
    SN_CUDA_CHECK(cudaFuncSetAttribute(sn_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, snt::SMEM_BYTES));
    SN_CUDA_CHECK(cudaFuncSetAttribute(sn_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, snt::SMEM_BYTES));
    return SN_OK;
}

void sn_tiled_release(sn_handle *h)
{
    delete reinterpret_cast<CUtensorMap *>(h->tmap);
    h->tmap = nullptr;
}

static SnSweepArgs sn_sweep_args(sn_handle *h);
static int sn_slab_phase_sync(sn_handle *h, long long *launches);

int sn_sweep_tiled_launch(sn_handle *h, long long nsweeps, long long *launches)
{
    const SnGeom &G = h->G;
    const CUtensorMap &tm = *reinterpret_cast<const CUtensorMap *>(h->tmap);
    for (long long s = 0; s < nsweeps; s++) {
        const SnSweepArgs a = sn_sweep_args(h);
        for (int p = 0; p < 8; p++) {
            SnTilePhase ph;
            ph.px = (p >> 2) & 1; ph.py = (p >> 1) & 1; ph.pz = p & 1;
            ph.hx = G.X / 32; ph.hy = G.Y / 32; ph.hz = G.nz / 32; ph.nrep = h->p.nreplicas;
            const long long ntiles = (long long)ph.hx * ph.hy * ph.hz * ph.nrep;
            const int grid = (int)std::min<long long>(ntiles, h->num_sms);
            if (h->species) sn_tiled_kernel<true><<<grid, snt::THREADS, snt::SMEM_BYTES, h->stream>>>(tm, a, ph);
            else sn_tiled_kernel<false><<<grid, snt::THREADS, snt::SMEM_BYTES, h->stream>>>(tm, a, ph);
            if (launches) (*launches)++;
            if (!G.periodic_z) { int rc = sn_slab_phase_sync(h, launches); if (rc) return rc; }
        }
        h->sweep++;
    }
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}
