// sn_sweep_colour.cuh -- colour-sublattice Metropolis passes over global memory.
//
// Replaces MC_moves/MC_move (montecarlo-core.c:143-191) for any lattice shape and
// cut-off.  One launch updates every site of one colour (cx,cy,cz); same-colour
// sites are more than DipoleCutOff apart, so their dE (site_energy,
// montecarlo-core.c:76-141) never reads a site written in the same launch.
// It is the general path (odd sizes, Z==1, cut-off != 3, small lattices batched
// as replicas); the TMA/shared-memory kernel in sn_sweep_tiled.cuh is the fast
// path for large cut-off-3 lattices.  HBM/L2 bound: ~123 float4 gathers per
// attempt.
#pragma once

#include "sn_field.cuh"

struct SnSweepArgs {
    float4 *lat;                    // replica 0 base (padded)
    SnGeom G;
    SnAxisColour ax, ay, az;
    const float *beta;              // per replica
    const float4 *efield;           // per replica: (E_x, E_y, E_z, CageStrain)
    float K;
    int constrain, dim;
    unsigned long long *counters;   // per replica {accept, reject, vacant}
    const uint4 *rep_key;           // per replica: Philox key (x, y) and the tag xor-ed into counter word 1 (z)
    uint32_t sweep_lo, sweep_hi;
    const SnNbEntry *nb;
    int nnb;
    float4 *peer_lo, *peer_hi;      // Z-slab neighbours' padded lattices (replica 0 base) or null
    float *audit;                   // sn_mc_sweep_audit: one SN_AUDIT_WORDS record per attempt, [rep][x][y][z], or null
    int audit_group;                // ordinal of this launch's group of mutually independent sites
};

// One audit record (include/starrynight_b200.h, sn_mc_sweep_audit): what the kernel proposed, drew, computed and decided.
__device__ __forceinline__ void sn_audit_write(float *__restrict__ audit, const SnGeom &G, int rep, int x, int y, int z,
                                               const float3 np, float u, float dE, bool accepted, bool vacant, int group)
{
    float4 *r = reinterpret_cast<float4 *>(audit + ((((long long)rep * G.X + x) * G.Y + y) * G.nz + z) * SN_AUDIT_WORDS);
    r[0] = make_float4(np.x, np.y, np.z, u);
    r[1] = make_float4(dE, vacant ? 2.0f : (accepted ? 1.0f : 0.0f), (float)group, 0.0f);
}

// Ghost images of a boundary site: periodic images in x, y (and z when the handle
// owns the whole Z), and the neighbouring slabs' ghost planes.  Rare (surface
// sites only), so kept out of line to keep the sweep kernels' hot code small.
__device__ __noinline__ void sn_store_images(float4 *__restrict__ lat, float4 *__restrict__ peer_lo,
                                             float4 *__restrict__ peer_hi, const SnGeom &G,
                                             int x, int y, int z, const float4 v)
{
    const int g = G.g, gz = G.gz;
    const bool bx = x < g || x >= G.X - g, by = y < g || y >= G.Y - g;
    const bool bz = gz > 0 && (z < gz || z >= G.nz - gz);
    const int kx = bx ? (g + G.X - 1) / G.X : 0, ky = by ? (g + G.Y - 1) / G.Y : 0;
    const int kz = (bz && G.periodic_z) ? (gz + G.nz - 1) / G.nz : 0;
    for (int ix = -kx; ix <= kx; ix++) {
        const int xi = x + ix * G.X;
        if (xi < -g || xi >= G.X + g) continue;
        for (int iy = -ky; iy <= ky; iy++) {
            const int yi = y + iy * G.Y;
            if (yi < -g || yi >= G.Y + g) continue;
            for (int iz = -kz; iz <= kz; iz++) {
                const int zi = z + iz * G.nz;
                if (zi < -gz || zi >= G.nz + gz) continue;
                if (ix | iy | iz) lat[sn_pidx(G, xi, yi, zi)] = v;
            }
            if (bz && !G.periodic_z) {              // push to the slab neighbours over NVLink
                if (z < gz && peer_lo) peer_lo[sn_pidx(G, xi, yi, z + G.nz)] = v;
                if (z >= G.nz - gz && peer_hi) peer_hi[sn_pidx(G, xi, yi, z - G.nz)] = v;
            }
        }
    }
}

// Write a site and, if it lies within the ghost width of a face, every image of it.  When every
// interacting extent is at least twice the ghost width a site has at most one image shift per axis
// (<= 7 images, a few predicated stores); tiny lattices take the general out-of-line path.
__device__ __forceinline__ void sn_store_site(float4 *__restrict__ lat, float4 *__restrict__ peer_lo,
                                              float4 *__restrict__ peer_hi, const SnGeom &G,
                                              int x, int y, int z, const float4 v)
{
    lat[sn_pidx(G, x, y, z)] = v;
    const int g = G.g, gz = G.gz;
    if (G.X < 2 * g || G.Y < 2 * g || (gz > 0 && G.nz < 2 * gz)) {
        const bool bx = x < g || x >= G.X - g, by = y < g || y >= G.Y - g;
        const bool bz = gz > 0 && (z < gz || z >= G.nz - gz);
        if (bx || by || bz) sn_store_images(lat, peer_lo, peer_hi, G, x, y, z, v);
        return;
    }
    const int ix = x < g ? G.X : (x >= G.X - g ? -G.X : 0);
    const int iy = y < g ? G.Y : (y >= G.Y - g ? -G.Y : 0);
    const int iz = gz == 0 ? 0 : (z < gz ? G.nz : (z >= G.nz - gz ? -G.nz : 0));
    if ((ix | iy | iz) == 0) return;
    if (ix) lat[sn_pidx(G, x + ix, y, z)] = v;
    if (iy) lat[sn_pidx(G, x, y + iy, z)] = v;
    if (ix && iy) lat[sn_pidx(G, x + ix, y + iy, z)] = v;
    if (iz) {
        float4 *__restrict__ dst = G.periodic_z ? lat : (iz > 0 ? peer_lo : peer_hi);
        if (dst) {
            dst[sn_pidx(G, x, y, z + iz)] = v;
            if (ix) dst[sn_pidx(G, x + ix, y, z + iz)] = v;
            if (iy) dst[sn_pidx(G, x, y + iy, z + iz)] = v;
            if (ix && iy) dst[sn_pidx(G, x + ix, y + iy, z + iz)] = v;
        }
    }
}

__device__ __forceinline__ void sn_count(unsigned long long *__restrict__ c, bool attempted, bool accepted, bool vacant)
{
    const unsigned m = 0xffffffffu;          // every thread of the block reaches this point
    const int na = __popc(__ballot_sync(m, accepted));
    const int nr = __popc(__ballot_sync(m, attempted && !accepted));
    const int nv = __popc(__ballot_sync(m, vacant));
    if ((threadIdx.x & 31) == 0) {
        if (na) atomicAdd(c + 0, (unsigned long long)na);
        if (nr) atomicAdd(c + 1, (unsigned long long)nr);
        if (nv) atomicAdd(c + 2, (unsigned long long)nv);
    }
}

// MODE 0: cut-off 3, 3-D (122 neighbours); 1: cut-off 3, Z==1 (28); 2: table-driven
template <int MODE, bool SPECIES>
__global__ void __launch_bounds__(128) sn_colour_pass_kernel(const SnSweepArgs a, const int cx, const int cy, const int cz)
{
    const int nx = sn_axis_count(a.ax, cx), ny = sn_axis_count(a.ay, cy), nz = sn_axis_count(a.az, cz);
    const long long total = (long long)nx * ny * nz;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int rep = blockIdx.y;
    bool attempted = false, accepted = false, vacant = false;
    if (idx < total) {
        const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((long long)nz * ny));
        const int x = sn_axis_coord(a.ax, cx, i), y = sn_axis_coord(a.ay, cy, j), z = sn_axis_coord(a.az, cz, k);
        float4 *lat = a.lat + (long long)rep * a.G.rep_stride;
        const float4 *site = lat + sn_pidx(a.G, x, y, z);
        const float4 old = *site;
        if (old.w == 0.0f) {                                    // montecarlo-core.c:163
            vacant = true;
            if (a.audit) sn_audit_write(a.audit, a.G, rep, x, y, z, make_float3(0.f, 0.f, 0.f), 0.f, 0.f, false, true, a.audit_group);
        } else {
            attempted = true;
            float3 F = make_float3(0.f, 0.f, 0.f), Gc = make_float3(0.f, 0.f, 0.f);
            const long long sx = a.G.sx, sy = a.G.sy;
            auto load = [&](int dx, int dy, int dz) { return site[dx * sx + dy * sy + dz]; };
            if constexpr (MODE == 0) sn_local_field_cut3<false, SPECIES>(load, F, Gc);
            else if constexpr (MODE == 1) sn_local_field_cut3<true, SPECIES>(load, F, Gc);
            else sn_local_field_table(a.nb, a.nnb, load, F, Gc);
            SnTerms t;
            const float4 E = a.efield[rep];
            t.cage = E.w; t.K = a.K; t.beta = a.beta[rep];
            t.E = make_float3(E.x, E.y, E.z);
            t.constrain = a.constrain; t.dim = a.dim;
            const unsigned long long gsite = ((unsigned long long)x * a.G.Y + y) * a.G.Z + (a.G.z0 + z);
            const uint4 key = a.rep_key[rep];
            const Philox4 r = sn_philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32) ^ key.z, a.sweep_lo, a.sweep_hi, key.x, key.y);
            const float3 np = sn_propose(t, sn_u01(r.x), sn_u01(r.y));
            const float dE = sn_delta_e(old, np, F, Gc, t);
            const float ua = sn_u01_32(r.z);
            accepted = sn_accept(dE, t.beta, ua);
            if (a.audit) sn_audit_write(a.audit, a.G, rep, x, y, z, np, ua, dE, accepted, false, a.audit_group);
            if (accepted) {
                float4 *plo = a.peer_lo ? a.peer_lo + (long long)rep * a.G.rep_stride : nullptr;
                float4 *phi = a.peer_hi ? a.peer_hi + (long long)rep * a.G.rep_stride : nullptr;
                sn_store_site(lat, plo, phi, a.G, x, y, z, make_float4(np.x, np.y, np.z, old.w));
            }
        }
    }
    sn_count(a.counters + 3 * rep, attempted, accepted, vacant);
}

// Rebuild every ghost cell from the cells it mirrors (after sn_set_lattice /
// sn_set_ghost).  One thread per padded cell.
__global__ void sn_refresh_ghosts_kernel(float4 *lat, const SnGeom G)
{
    const long long n = G.rep_stride;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 *base = lat + (long long)blockIdx.y * G.rep_stride;
    const int zp = (int)(i % G.PZ) - G.gz, yp = (int)((i / G.PZ) % G.PY) - G.g, xp = (int)(i / ((long long)G.PZ * G.PY)) - G.g;
    auto wrap = [](int v, int n_) { v %= n_; return v < 0 ? v + n_ : v; };
    const int xs = wrap(xp, G.X), ys = wrap(yp, G.Y);
    const int zs = G.periodic_z ? wrap(zp, G.nz) : zp;
    if (xs == xp && ys == yp && zs == zp) return;
    base[i] = base[sn_pidx(G, xs, ys, zs)];
}
