// sn_sweep_resident.cuh -- Metropolis sweeps of a lattice that lives in shared memory.
//
// Replaces MC_moves -> MC_move -> site_energy (montecarlo-core.c:76-191) for lattices of up to
// ~14 000 sites (X*Y*Z*16 B <= 227 KB): the 2-D 100x100 case of the reference's figures, its
// `make test` 20x20x28 lattice, and the small 3-D lattices of temperature / field sweeps, usually
// batched as replicas.  One CTA owns one replica: it loads the lattice into shared memory once,
// runs EVERY sweep of the sn_mc_sweeps call there -- all colour sublattices back to back with a
// block barrier between them instead of a kernel launch -- and writes the lattice back once.  HBM
// sees 32 B per site per call instead of ~2 KB per site per sweep, and there are no launch gaps
// (the colour-pass kernel needs 16-64 launches per sweep, each a few microseconds of mostly
// latency at this size).
//
// Same colour order, same Philox streams and the same arithmetic (sn_local_field_*, sn_delta_e)
// as sn_colour_pass_kernel, so the chain is bit-identical to it; the periodic wrap that the
// padded global array resolves with ghost cells is resolved here by per-axis offset tables.
#pragma once

#include "sn_sweep_colour.cuh"

namespace snr {
constexpr int MAX_SMEM = 227 * 1024;
__host__ __device__ constexpr int threads(int mode) { return mode == 1 ? 640 : 256; }   // 2-D: 625 sites per colour at 100x100

// Shared-memory layout: colour-major.  The threads of a colour pass own sites P = cutoff + 1 apart along every
// interacting axis; stored in lattice order their float4 would sit 64 B (and whole rows) apart and every LDS.128
// of a warp would take several times the wavefronts.  Every axis is therefore de-interleaved by its colour
// period: site (x,y,z) lives at
//     ((((x%Px) Py + y%Py) Pz + z%Pz) Qx + x/Px) Qy Qz + (y/Py) Qz + z/Pz,     Q = ceil(extent / P),
// so all sites of one colour are contiguous in the order the threads walk them and a neighbour class is another
// contiguous block: the lanes of a pass read consecutive float4 (bank-conflict free).  The index is a sum of
// per-axis terms, kept in tables (built once per CTA) for every coordinate -cutoff .. n + cutoff - 1 with the
// periodic wrap already applied; a neighbour address is three table entries added up.
struct Layout {
    int Px, Qx, Py, Qy, Pz, Qz;
    int A, B, C, D, E;      // element strides of x%Px, x/Px, y%Py, y/Py, z%Pz (z/Pz has stride 1)
    int cells;              // float4 slots of the tile (>= X*Y*Z when an extent is not a multiple of its period)
    int g;                  // table margin = cutoff
    int tab_off;            // byte offset of the tables behind the tile
};
__host__ __device__ inline Layout layout(const SnGeom &G, int cutoff, int Px, int Py, int Pz)
{
    Layout L;
    L.g = cutoff;
    L.Px = Px; L.Py = Py; L.Pz = Pz;
    L.Qx = (G.X + Px - 1) / Px; L.Qy = (G.Y + Py - 1) / Py; L.Qz = (G.nz + Pz - 1) / Pz;
    L.D = L.Qz; L.B = L.Qy * L.Qz; L.E = L.Qx * L.B; L.C = Pz * L.E; L.A = Py * L.C;
    L.cells = Px * L.A;
    L.tab_off = L.cells * 16;
    return L;
}
__host__ __device__ inline int smem_bytes(const SnGeom &G, const Layout &L) { return L.tab_off + 4 * (G.X + G.Y + G.nz + 6 * L.g); }
}

template <int MODE, bool SPECIES>
__global__ void __launch_bounds__(snr::threads(MODE), 1)
sn_resident_kernel(const SnSweepArgs a, const snr::Layout L, const int nrep, const unsigned long long sweep0, const int nsweeps)
{
    extern __shared__ __align__(16) unsigned char sn_resident_smem[];
    float4 *tile = reinterpret_cast<float4 *>(sn_resident_smem);
    const SnGeom &G = a.G;
    const int X = G.X, Y = G.Y, Z = G.nz, N = X * Y * Z, tid = threadIdx.x, nthr = blockDim.x, g = L.g;
    int *xtab = reinterpret_cast<int *>(sn_resident_smem + L.tab_off), *ytab = xtab + X + 2 * g, *ztab = ytab + Y + 2 * g;
    auto wrapn = [](int v, int n) { v %= n; return v < 0 ? v + n : v; };
    for (int i = tid; i < X + 2 * g; i += nthr) { const int x = wrapn(i - g, X); xtab[i] = (x % L.Px) * L.A + (x / L.Px) * L.B; }
    for (int i = tid; i < Y + 2 * g; i += nthr) { const int y = wrapn(i - g, Y); ytab[i] = (y % L.Py) * L.C + (y / L.Py) * L.D; }
    for (int i = tid; i < Z + 2 * g; i += nthr) { const int z = wrapn(i - g, Z); ztab[i] = (z % L.Pz) * L.E + z / L.Pz; }
    __syncthreads();

    for (int rep = blockIdx.x; rep < nrep; rep += gridDim.x) {
        float4 *lat = a.lat + (long long)rep * G.rep_stride;
        for (int i = tid; i < N; i += nthr) {
            const int z = i % Z, y = (i / Z) % Y, x = i / (Z * Y);
            tile[xtab[x + g] + ytab[y + g] + ztab[z + g]] = lat[sn_pidx(G, x, y, z)];
        }
        __syncthreads();

        SnTerms t;
        t.K = a.K; t.beta = a.beta[rep];
        { const float4 E = a.efield[rep]; t.E = make_float3(E.x, E.y, E.z); t.cage = E.w; }
        t.constrain = a.constrain; t.dim = a.dim;
        const uint4 key = a.rep_key[rep];
        int n_acc = 0, n_rej = 0, n_vac = 0;
        // The regular colours (c < P) all have the same site counts, so a thread's first site of a pass has the same
        // (i,j,k) in every one of them: decode once, not with four integer divisions per pass.
        const int rnx = sn_axis_count(a.ax, 0), rny = sn_axis_count(a.ay, 0), rnz = sn_axis_count(a.az, 0);
        const int rk = tid % rnz, rj = (tid / rnz) % rny, ri = tid / (rnz * rny);

        for (int s = 0; s < nsweeps; s++) {
            const unsigned long long sw = sweep0 + (unsigned long long)s;
            const uint32_t sweep_lo = (uint32_t)sw, sweep_hi = (uint32_t)(sw >> 32);
            for (int cx = 0; cx < a.ax.ncol; cx++) for (int cy = 0; cy < a.ay.ncol; cy++) for (int cz = 0; cz < a.az.ncol; cz++) {
                const int nx = sn_axis_count(a.ax, cx), ny = sn_axis_count(a.ay, cy), nz = sn_axis_count(a.az, cz);
                const int total = nx * ny * nz;
                const bool regular = nx == rnx && ny == rny && nz == rnz;
                for (int idx = tid; idx < total; idx += nthr) {
                    int k, j, i;
                    if (regular && idx == tid) { k = rk; j = rj; i = ri; }
                    else { k = idx % nz; j = (idx / nz) % ny; i = idx / (nz * ny); }
                    const int x = sn_axis_coord(a.ax, cx, i), y = sn_axis_coord(a.ay, cy, j), z = sn_axis_coord(a.az, cz, k);
                    const int c = xtab[x + g] + ytab[y + g] + ztab[z + g];
                    const float4 old = tile[c];
                    if (old.w == 0.0f) { n_vac++; continue; }                     // montecarlo-core.c:163
                    float3 F = make_float3(0.f, 0.f, 0.f), Gc = make_float3(0.f, 0.f, 0.f);
                    if constexpr (MODE == 2) {
                        const int *xt = xtab + x + g, *yt = ytab + y + g, *zt = ztab + z + g;
                        auto load = [&](int dx, int dy, int dz) { return tile[xt[dx] + yt[dy] + zt[dz]]; };
                        sn_local_field_table(a.nb, a.nnb, load, F, Gc);
                    } else {
                        // element offsets of the 7 (wrapped) coordinates per axis: registers after unrolling
                        int xa[7], ya[7], za[7];
#pragma unroll
                        for (int d = 0; d < 7; d++) {
                            xa[d] = xtab[x + d]; ya[d] = ytab[y + d];
                            za[d] = MODE == 1 ? 0 : ztab[z + d];
                        }
                        auto load = [&](int dx, int dy, int dz) { return tile[xa[dx + 3] + ya[dy + 3] + za[dz + 3]]; };
                        sn_local_field_cut3<MODE == 1, SPECIES>(load, F, Gc);
                    }
                    const unsigned long long gsite = ((unsigned long long)x * G.Y + y) * G.Z + (G.z0 + z);
                    const Philox4 r = sn_philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32) ^ key.z, sweep_lo, sweep_hi, key.x, key.y);
                    const float3 np = sn_propose(t, sn_u01(r.x), sn_u01(r.y));
                    const float dE = sn_delta_e(old, np, F, Gc, t);
                    const bool accepted = sn_accept(dE, t.beta, sn_u01_32(r.z));
                    if (accepted) tile[c] = make_float4(np.x, np.y, np.z, old.w);
                    n_acc += accepted; n_rej += !accepted;
                }
                __syncthreads();
            }
        }

        for (int i = tid; i < N; i += nthr) {
            const int z = i % Z, y = (i / Z) % Y, x = i / (Z * Y);
            lat[sn_pidx(G, x, y, z)] = tile[xtab[x + g] + ytab[y + g] + ztab[z + g]];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
            n_rej += __shfl_xor_sync(0xffffffffu, n_rej, o);
            n_vac += __shfl_xor_sync(0xffffffffu, n_vac, o);
        }
        if ((tid & 31) == 0) {
            unsigned long long *cnt = a.counters + 3 * rep;
            if (n_acc) atomicAdd(cnt + 0, (unsigned long long)n_acc);
            if (n_rej) atomicAdd(cnt + 1, (unsigned long long)n_rej);
            if (n_vac) atomicAdd(cnt + 2, (unsigned long long)n_vac);
        }
        __syncthreads();                                  // the tile is reused by the CTA's next replica
    }
}

// ---- host side -------------------------------------------------------------------
bool sn_resident_supported(const sn_handle *h, std::string *why)
{
    const SnGeom &G = h->G;
    const int g = h->p.cutoff;
    const char *msg = nullptr;
    if (!G.periodic_z) msg = "Z-slab handle";
    else if ((long long)G.X * G.Y * G.Z * 16 > snr::MAX_SMEM ||
             snr::smem_bytes(G, snr::layout(G, g, sn_axis_colour(G.X, g, false).P, sn_axis_colour(G.Y, g, false).P,
                                            sn_axis_colour(G.Z, g, G.Z == 1).P)) > snr::MAX_SMEM)
        msg = "lattice does not fit in 227 KB of shared memory";
    else if (G.X < g || G.Y < g || (G.Z > 1 && G.Z < g)) msg = "an extent is smaller than DipoleCutOff";
    if (msg) { if (why) *why = msg; return false; }
    return true;
}

// Sweeps per launch: the kernel counts accepted / rejected / vacant attempts in 32-bit registers per thread and adds them
// up over a warp before they go to the 64-bit counters, so one launch may make at most 2^31 attempts per warp.
static long long sn_resident_chunk(const sn_handle *h, int threads)
{
    const long long sites = (long long)h->G.X * h->G.Y * h->G.Z;
    const long long per_warp = 32 * ((sites + threads - 1) / threads);       // attempts of one warp per sweep, at most
    return std::max<long long>(1, std::min<long long>(1LL << 30, ((1LL << 31) - 1) / per_warp));
}

template <int MODE, bool SPECIES>
static int sn_resident_launch_t(sn_handle *h, const SnSweepArgs &a, long long nsweeps, long long *launches)
{
    const snr::Layout L = snr::layout(h->G, h->p.cutoff, a.ax.P, a.ay.P, a.az.P);
    const int smem = snr::smem_bytes(h->G, L);
    auto kern = sn_resident_kernel<MODE, SPECIES>;
    SN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = std::min(h->p.nreplicas, h->num_sms);
    long long done = 0;
    const long long limit = sn_resident_chunk(h, snr::threads(MODE));
    while (done < nsweeps) {
        const int chunk = (int)std::min<long long>(nsweeps - done, limit);
        kern<<<grid, snr::threads(MODE), smem, h->stream>>>(a, L, h->p.nreplicas, h->sweep + (unsigned long long)done, chunk);
        done += chunk;
        if (launches) (*launches)++;
    }
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

int sn_sweep_resident_launch(sn_handle *h, long long nsweeps, long long *launches)
{
    if (nsweeps <= 0) return SN_OK;
    { int rc = sn_sync_canonical(h); if (rc) return rc; }
    h->lat2_valid = false;
    const SnSweepArgs a = sn_sweep_args(h);
    const int mode = h->p.cutoff == 3 ? (h->p.Z == 1 ? 1 : 0) : 2;
    int rc;
    if (mode == 0) rc = h->species ? sn_resident_launch_t<0, true>(h, a, nsweeps, launches) : sn_resident_launch_t<0, false>(h, a, nsweeps, launches);
    else if (mode == 1) rc = h->species ? sn_resident_launch_t<1, true>(h, a, nsweeps, launches) : sn_resident_launch_t<1, false>(h, a, nsweeps, launches);
    else rc = sn_resident_launch_t<2, true>(h, a, nsweeps, launches);
    if (rc) return rc;
    h->sweep += (unsigned long long)nsweeps;
    if ((rc = sn_refresh_ghosts(h))) return rc;           // the kernel wrote interior cells only
    if (launches) (*launches)++;
    return SN_OK;
}
