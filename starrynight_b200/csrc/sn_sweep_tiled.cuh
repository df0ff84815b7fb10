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
//     Two lanes own a segment of 4 consecutive z sites of one (x,y) column, i.e.
//     the four colours (cx,cy,0..3), 2 sites each.  None of the 28 neighbour
//     columns around it changes during the super-pass (they belong to other
//     (cx,cy) classes), so their contribution to the local fields is gathered
//     once with a sliding z window (one load serves both sites).  Only the centre
//     column changes: the 4 sites are then decided in sequence, the fields of the
//     later ones corrected in registers for the earlier accepted moves (partner
//     lane and the segment above via shuffles).  Every attempt is a full fresh dE
//     over the cut-off sphere.
//   * All 128 threads run one ~22 KB instruction stream (it has to stay inside the
//     instruction cache: a first version with two 30 KB warp roles spent half its
//     cycles waiting for instructions).  A thread gathers the fields of 2 sites;
//     neighbours r and -r share the tensor T(r) = T(-r), so their moments are added
//     first and the tensor applied once (582 instead of 798 FP ops per attempt).
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
constexpr int OFF_BAR = 4 * BOX_STRIDE_F4 * 16;            // mbarrier
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

// One neighbour pair r = (DX,DY,DZ) and -r: T(r) = T(-r), so the two moments are
// added first and the tensor applied once (12 instead of 18 FP ops for a full tensor).
template <int DX, int DY, int DZ, bool SPECIES>
__device__ __forceinline__ void sn_accumulate_pair(float3 &F, float3 &G, const float4 a, const float4 b)
{
    constexpr float txx = sn_T(DX, DY, DZ, 0, 0), tyy = sn_T(DX, DY, DZ, 1, 1), tzz = sn_T(DX, DY, DZ, 2, 2);
    constexpr float txy = sn_T(DX, DY, DZ, 0, 1), txz = sn_T(DX, DY, DZ, 0, 2), tyz = sn_T(DX, DY, DZ, 1, 2);
    float ax, ay, az;
    if constexpr (SPECIES) { ax = fmaf(a.x, a.w, b.x * b.w); ay = fmaf(a.y, a.w, b.y * b.w); az = fmaf(a.z, a.w, b.z * b.w); }
    else { ax = a.x + b.x; ay = a.y + b.y; az = a.z + b.z; }
    F.x = fmaf(txx, ax, F.x);
    F.y = fmaf(tyy, ay, F.y);
    F.z = fmaf(tzz, az, F.z);
    if constexpr (txy != 0.0f) { F.x = fmaf(txy, ay, F.x); F.y = fmaf(txy, ax, F.y); }
    if constexpr (txz != 0.0f) { F.x = fmaf(txz, az, F.x); F.z = fmaf(txz, ax, F.z); }
    if constexpr (tyz != 0.0f) { F.y = fmaf(tyz, az, F.y); F.z = fmaf(tyz, ay, F.z); }
    if constexpr (DX * DX + DY * DY + DZ * DZ == 1) {
        if constexpr (SPECIES) { G.x += a.x + b.x; G.y += a.y + b.y; G.z += a.z + b.z; }
        else { G.x += ax; G.y += ay; G.z += az; }
    }
}

// Local fields of the thread's 2 consecutive z sites from all 29 columns.  pe[e+3]
// points at the thread's own column, plane (first site + e); a neighbour column is
// a compile-time immediate away.  Each column pair (+c, -c) is loaded once for
// both sites (sliding z window) and combined with the pair symmetry.
template <bool SPECIES>
__device__ __forceinline__ void sn_tile_gather2(const float4 *const (&pe)[8], float3 (&F)[2], float3 (&G)[2], float4 (&old)[2])
{
    sn_static_for<-3, 4>([&](auto dxc) {
        sn_static_for<-3, 4>([&](auto dyc) {
            constexpr int DX = decltype(dxc)::value, DY = decltype(dyc)::value;
            constexpr int r2xy = DX * DX + DY * DY;
            constexpr bool upper = DX > 0 || (DX == 0 && DY > 0);
            if constexpr (r2xy <= 9 && upper) {
                constexpr int M = snt::half_height(r2xy);
                constexpr int C = (DX * snt::BX + DY) * snt::NQ;
                float4 wp[2 * M + 2], wm[2 * M + 2];
                sn_static_for<-M, 2 + M>([&](auto ec) {
                    constexpr int E = decltype(ec)::value;
                    wp[E + M] = pe[E + 3][C];
                    wm[E + M] = pe[E + 3][-C];
                });
                sn_static_for<0, 2>([&](auto sc) {
                    sn_static_for<-M, M + 1>([&](auto dzc) {
                        constexpr int S = decltype(sc)::value, DZ = decltype(dzc)::value;
                        sn_accumulate_pair<DX, DY, DZ, SPECIES>(F[S], G[S], wp[S + DZ + M], wm[S - DZ + M]);
                    });
                });
            }
        });
    });
    {   // own column: pairs (0,0,+-dz); also yields the current values of the 2 sites
        float4 w[8];
        sn_static_for<0, 8>([&](auto ec) { constexpr int E = decltype(ec)::value; w[E] = pe[E][0]; });
        old[0] = w[3]; old[1] = w[4];
        sn_static_for<0, 2>([&](auto sc) {
            sn_static_for<1, 4>([&](auto dzc) {
                constexpr int S = decltype(sc)::value, DZ = decltype(dzc)::value;
                sn_accumulate_pair<0, 0, DZ, SPECIES>(F[S], G[S], w[3 + S + DZ], w[3 + S - DZ]);
            });
        });
    }
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
    const uint32_t bar = sn_smem_u32(smem + snt::OFF_BAR);

    // thread -> (column i,j ; segment k of 4 z sites ; half h of the segment).  k and the low bit of j
    // vary inside a quarter-warp, so its 8 lanes read 8 distinct 16-byte bank groups (28 j + k mod 8).
    const int tid = threadIdx.x, lane = tid & 31, i = tid >> 5;
    const int k = lane & 3, h = (lane >> 3) & 1, j = ((lane >> 2) & 1) | ((lane >> 4) << 1);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (long long)ph.hx * ph.hy * ph.hz * ph.nrep;
    uint32_t parity = 0;
    const SnGeom &G = a.G;

    // z part of the shared-memory index for plane (4k + 2h + e), e = -3..4
    int zoff[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int w = 4 + 4 * k + 2 * h + (e - 3);          // plane index inside the 28-plane window
        zoff[e] = (w & 3) * snt::BOX_STRIDE_F4 + (w >> 2);
    }

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

        sn_mbar_wait(bar, parity);
        parity ^= 1;

        int n_acc = 0, n_rej = 0, n_vac = 0;
#pragma unroll 1
        for (int sp = 0; sp < 16; sp++) {
            const int cx = sp >> 2, cy = sp & 3;
            const int gx = x0 + cx + 4 * i, gy = y0 + cy + 4 * j, gz = z0 + 4 * k + 2 * h;
            const int colbase = ((snt::H + cx + 4 * i) * snt::BX + (snt::H + cy + 4 * j)) * snt::NQ;
            const float4 *pe[8];
#pragma unroll
            for (int e = 0; e < 8; e++) pe[e] = tile + colbase + zoff[e];

            // trial orientations for the 2 sites (Philox keyed by global site, replica, sweep)
            float3 np[2]; float ua[2];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const unsigned long long gsite = ((unsigned long long)gx * G.Y + gy) * G.Z + (G.z0 + gz + s);
                const Philox4 r = sn_philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32) ^ ((uint32_t)rep << 8),
                                                   a.sweep_lo, a.sweep_hi, a.key0, a.key1);
                np[s] = sn_propose(tm, sn_u01(r.x), sn_u01(r.y));
                ua[s] = sn_u01(r.z);
            }

            float3 F[2], Gc[2];
            float4 old[2];
#pragma unroll
            for (int s = 0; s < 2; s++) { F[s] = make_float3(0.f, 0.f, 0.f); Gc[s] = make_float3(0.f, 0.f, 0.f); }
            sn_tile_gather2<SPECIES>(pe, F, Gc, old);

            // The 4 sites of a segment interact: decide them in z order.  Step t belongs to the lane with
            // h == t/2 (local site t&1); its accepted change is broadcast to the partner lane (lane^8) and to
            // the segment below (lane-1 reads lane's value as "the segment above"), and folded into the
            // fields of the later sites: T(0,0,dz) = diag(1,1,-2)/|dz|^3, cage term for |dz| = 1.
            float3 dmS[4], dpS[4], dmU[4], dpU[4];
            bool accepted[2] = {false, false};
#pragma unroll
            for (int t4 = 0; t4 < 4; t4++) {
                const int s = t4 & 1;
                float3 Fs = F[s], Gs = Gc[s];
#pragma unroll
                for (int t2 = 0; t2 < t4; t2++) {                 // earlier sites of this segment, dz = t2 - t4
                    const int d = t4 - t2;
                    const float w3 = d == 1 ? 1.0f : d == 2 ? 0.125f : (1.0f / 27.0f);
                    Fs.x = fmaf(w3, dmS[t2].x, Fs.x); Fs.y = fmaf(w3, dmS[t2].y, Fs.y); Fs.z = fmaf(-2.0f * w3, dmS[t2].z, Fs.z);
                    if (d == 1) { Gs.x += dpS[t2].x; Gs.y += dpS[t2].y; Gs.z += dpS[t2].z; }
                }
#pragma unroll
                for (int t2 = 0; t2 < t4; t2++) {                 // earlier sites of the segment above, dz = 4 + t2 - t4
                    const int d = 4 + t2 - t4;
                    const float w3 = d == 1 ? 1.0f : d == 2 ? 0.125f : (1.0f / 27.0f);
                    Fs.x = fmaf(w3, dmU[t2].x, Fs.x); Fs.y = fmaf(w3, dmU[t2].y, Fs.y); Fs.z = fmaf(-2.0f * w3, dmU[t2].z, Fs.z);
                    if (d == 1) { Gs.x += dpU[t2].x; Gs.y += dpU[t2].y; Gs.z += dpU[t2].z; }
                }
                const float4 o = old[s];
                const bool mine = h == (t4 >> 1);
                const bool vacant = o.w == 0.0f;                                            // montecarlo-core.c:163
                const float dE = sn_delta_e(o, np[s], Fs, Gs, tm);
                const bool acc = mine && !vacant && sn_accept(dE, tm.beta, ua[s]);          // montecarlo-core.c:179
                float3 dp = acc ? make_float3(np[s].x - o.x, np[s].y - o.y, np[s].z - o.z) : make_float3(0.f, 0.f, 0.f);
                if (acc) *const_cast<float4 *>(pe[3 + s]) = make_float4(np[s].x, np[s].y, np[s].z, o.w);
                if (mine) accepted[s] = acc;
                n_acc += acc; n_rej += (mine && !acc && !vacant); n_vac += (mine && vacant);
                if (t4 < 3) {
                    const int src = (lane & 23) | ((t4 >> 1) << 3);                        // the lane that owns step t4
                    dpS[t4].x = __shfl_sync(0xffffffffu, dp.x, src);
                    dpS[t4].y = __shfl_sync(0xffffffffu, dp.y, src);
                    dpS[t4].z = __shfl_sync(0xffffffffu, dp.z, src);
                    if constexpr (SPECIES) {
                        const float ow = __shfl_sync(0xffffffffu, o.w, src);
                        dmS[t4] = make_float3(ow * dpS[t4].x, ow * dpS[t4].y, ow * dpS[t4].z);
                    } else dmS[t4] = dpS[t4];
                    dpU[t4].x = __shfl_down_sync(0xffffffffu, dpS[t4].x, 1);
                    dpU[t4].y = __shfl_down_sync(0xffffffffu, dpS[t4].y, 1);
                    dpU[t4].z = __shfl_down_sync(0xffffffffu, dpS[t4].z, 1);
                    if constexpr (SPECIES) {
                        dmU[t4].x = __shfl_down_sync(0xffffffffu, dmS[t4].x, 1);
                        dmU[t4].y = __shfl_down_sync(0xffffffffu, dmS[t4].y, 1);
                        dmU[t4].z = __shfl_down_sync(0xffffffffu, dmS[t4].z, 1);
                    } else dmU[t4] = dpU[t4];
                    if (k == 3) { dpU[t4] = make_float3(0.f, 0.f, 0.f); dmU[t4] = make_float3(0.f, 0.f, 0.f); }   // above lies the static halo
                }
            }
            // accepted moves go straight to global memory (and to every ghost image / the neighbour GPU)
#pragma unroll 1
            for (int s = 0; s < 2; s++)
                if (s == 0 ? accepted[0] : accepted[1])
                    sn_store_site(glat, plo, phi, G, gx, gy, gz + s, make_float4(s == 0 ? np[0].x : np[1].x, s == 0 ? np[0].y : np[1].y,
                                                                                 s == 0 ? np[0].z : np[1].z, s == 0 ? old[0].w : old[1].w));
            __syncthreads();
        }
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
