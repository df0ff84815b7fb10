// sn_observables.cuh -- lattice-wide observables as warp-shuffle + block reductions.
//
// Replaces the serial loops of starrynight-analysis.c:
//   polarisation()            :48-62     sum_i p_i                       -> sn_sum_dipoles_kernel
//   landau_order()            :506-526   |sum_i p_i|^2                   -> same sums, finished on the host
//   radial_order_parameter()  :528-598   FE / AFE correlations by r^2    -> sn_rdf_kernel
//   dipole_potential()        :65-94     V_i = sum_j l_j p_j.r / d^3     -> sn_potential_kernel
//   dipole_electricfield()    :393-465   |E_i|, E_i = sum_j (3 n n.p_j - p_j)/d^3 - p_i/3  -> sn_efield_kernel
//   dipole_electricfieldoffset() :310-376  the same half a lattice step off the sites    -> sn_efield_kernel
//   recombination_calculator():96-170    Boltzmann / Fermi-Dirac partition sums of V     -> sn_recombination_kernel
// and the lattice energy the reference never finished (main.c:63)        -> sn_energy_f32_kernel.
// Accumulation is FP64 (int64 for counts): the reference's float sums and int
// counts stop being sound beyond ~128^3.
// Partial sums are written per block and added in a fixed order, so results are
// bit-reproducible run to run.
#pragma once

#include "sn_field.cuh"

__device__ __forceinline__ double sn_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV doubles per thread; result valid in thread 0
template <int NV, int BLOCK>
__device__ __forceinline__ void sn_block_sum(double (&v)[NV], double *smem /* NV * BLOCK/32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        v[k] = sn_warp_sum(v[k]);
        if (lane == 0) smem[k * (BLOCK / 32) + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = lane < BLOCK / 32 ? smem[k * (BLOCK / 32) + lane] : 0.0;
            v[k] = sn_warp_sum(t);
        }
    }
    __syncthreads();
}

// periodic wrap of v = coordinate + offset.  NEAR: |offset| <= n is guaranteed by the caller (every extent at least
// the stencil radius), so one conditional add / subtract replaces the integer division.
template <bool NEAR>
__device__ __forceinline__ int sn_wrap(int v, int n)
{
    if constexpr (NEAR) { v += v < 0 ? n : 0; v -= v >= n ? n : 0; return v; }
    else { v %= n; return v < 0 ? v + n : v; }
}

__device__ __forceinline__ void sn_site_of(const SnGeom &G, long long i, int &x, int &y, int &z)
{
    z = (int)(i % G.nz); y = (int)((i / G.nz) % G.Y); x = (int)(i / ((long long)G.nz * G.Y));
}

// out[block][3] = partial sum of p over the block's sites (grid-stride)
__global__ void __launch_bounds__(256) sn_sum_dipoles_kernel(const float4 *__restrict__ lat, const SnGeom G, double *__restrict__ out)
{
    __shared__ double sm[3 * 8];
    const long long n = (long long)G.X * G.Y * G.nz;
    double v[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int x, y, z; sn_site_of(G, i, x, y, z);
        const float4 p = lat[sn_pidx(G, x, y, z)];
        v[0] += p.x; v[1] += p.y; v[2] += p.z;
    }
    sn_block_sum<3, 256>(v, sm);
    if (threadIdx.x == 0) { out[3 * blockIdx.x] = v[0]; out[3 * blockIdx.x + 1] = v[1]; out[3 * blockIdx.x + 2] = v[2]; }
}

// 64-bit content hash of the handle's own sites: sum over sites (mod 2^64) of a splitmix64 of the GLOBAL site index
// and the four float bit patterns.  A sum, so slabs add up: the hash of a Z-slab decomposed lattice is the wrapped sum
// of the slabs' hashes and equals the single-GPU hash iff every site holds the same bits.
__device__ __forceinline__ unsigned long long sn_mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) sn_state_hash_kernel(const float4 *__restrict__ lat, const SnGeom G, unsigned long long *__restrict__ out)
{
    const long long n = (long long)G.X * G.Y * G.nz;
    unsigned long long acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int x, y, z; sn_site_of(G, i, x, y, z);
        const float4 p = lat[sn_pidx(G, x, y, z)];
        const unsigned long long gsite = ((unsigned long long)x * G.Y + y) * G.Z + (G.z0 + z);
        unsigned long long hsh = sn_mix64(gsite);
        hsh = sn_mix64(hsh ^ (((unsigned long long)__float_as_uint(p.x) << 32) | __float_as_uint(p.y)));
        hsh = sn_mix64(hsh ^ (((unsigned long long)__float_as_uint(p.z) << 32) | __float_as_uint(p.w)));
        acc += hsh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// generic deterministic reduction of n doubles into gridDim.x partials
__global__ void __launch_bounds__(256) sn_sum_doubles_kernel(const double *__restrict__ in, long long n, double *__restrict__ out)
{
    __shared__ double sm[8];
    double v[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[0] += in[i];
    sn_block_sum<1, 256>(v, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = v[0];
}

// Lattice energy in the sweep kernel's own FP32 arithmetic (SN_PREC_F32):
// per site e_dd = l_i p_i.F_i, e_cage = -Cs p_i.G_i, e_field = p_i.E, e_K.
// out[block][4] partial sums (un-halved).
template <int MODE, bool SPECIES>
__global__ void __launch_bounds__(256) sn_energy_f32_kernel(const float4 *__restrict__ lat, const SnGeom G, const SnNbEntry *__restrict__ nb,
                                                            int nnb, float cage, float K, float3 E, double *__restrict__ out)
{
    __shared__ double sm[4 * 8];
    const long long n = (long long)G.X * G.Y * G.nz;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int x, y, z; sn_site_of(G, i, x, y, z);
        const float4 *site = lat + sn_pidx(G, x, y, z);
        const float4 p = *site;
        float3 F = make_float3(0.f, 0.f, 0.f), Gc = make_float3(0.f, 0.f, 0.f);
        const long long sx = G.sx, sy = G.sy;
        auto load = [&](int dx, int dy, int dz) { return site[dx * sx + dy * sy + dz]; };
        if constexpr (MODE == 0) sn_local_field_cut3<false, SPECIES>(load, F, Gc);
        else if constexpr (MODE == 1) sn_local_field_cut3<true, SPECIES>(load, F, Gc);
        else sn_local_field_table(nb, nnb, load, F, Gc);
        v[0] += p.w * (p.x * F.x + p.y * F.y + p.z * F.z);
        v[1] += -cage * (p.x * Gc.x + p.y * Gc.y + p.z * Gc.z);
        v[2] += p.x * E.x + p.y * E.y + p.z * E.z;
        if (K > 0.0f) v[3] += -K * (fabsf(p.x) + fabsf(p.y));
    }
    sn_block_sum<4, 256>(v, sm);
    if (threadIdx.x == 0) for (int k = 0; k < 4; k++) out[4 * blockIdx.x + k] = v[k];
}

// SN_PREC_F32 audit of single trial moves: the sweep kernel's own dE
template <int MODE, bool SPECIES>
__global__ void __launch_bounds__(128) sn_site_energy_f32_kernel(const float4 *__restrict__ lat, const SnGeom G, const SnNbEntry *__restrict__ nb,
                                                                 int nnb, SnTerms t, int n, const int *__restrict__ sites,
                                                                 const float *__restrict__ newdip, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *site = lat + sn_pidx(G, sites[3 * i], sites[3 * i + 1], sites[3 * i + 2]);
    const float4 old = *site;
    float3 F = make_float3(0.f, 0.f, 0.f), Gc = make_float3(0.f, 0.f, 0.f);
    const long long sx = G.sx, sy = G.sy;
    auto load = [&](int dx, int dy, int dz) { return site[dx * sx + dy * sy + dz]; };
    if constexpr (MODE == 0) sn_local_field_cut3<false, SPECIES>(load, F, Gc);
    else if constexpr (MODE == 1) sn_local_field_cut3<true, SPECIES>(load, F, Gc);
    else sn_local_field_table(nb, nnb, load, F, Gc);
    out[i] = (double)sn_delta_e(old, make_float3(newdip[3 * i], newdip[3 * i + 1], newdip[3 * i + 2]), F, Gc, t);
}

// ---- radial order parameter (analysis.c:528-598) ------------------------------
// Offsets inside the radius-9 sphere are sorted by r^2 on the host; `first[b]`
// is the first offset of bin b.  Only one of every pair {d, -d} is walked: on a periodic lattice the
// pairs (i, i+d) over all sites i are the pairs (j-d, j) over all j, and both correlations are symmetric
// in their two dipoles, so the host doubles the sums of r^2 > 0 (half the 3071 pair terms per site).  Each thread walks every offset for its site,
// keeps the running FE / AFE sums of the current bin in registers, and the block
// reduces them once per bin -> out[block][bin][2].
struct SnRdfOffset { short dx, dy, dz, r2; };

template <bool NEAR>
__global__ void __launch_bounds__(256) sn_rdf_kernel(const float4 *__restrict__ lat, const SnGeom G, const SnRdfOffset *__restrict__ off,
                                                     const int *__restrict__ first, int nbins, double *__restrict__ out)
{
    __shared__ double sm[2 * 8];
    const long long n = (long long)G.X * G.Y * G.nz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    int x = 0, y = 0, z = 0;
    if (live) sn_site_of(G, i, x, y, z);
    const float4 a = live ? lat[sn_pidx(G, x, y, z)] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < nbins; b++) {
        double v[2] = {0.0, 0.0};
        const int e = first[b + 1];
        for (int o = first[b]; o < e && live; o++) {
            const SnRdfOffset f = off[o];
            const int xx = sn_wrap<NEAR>(x + f.dx, G.X), yy = sn_wrap<NEAR>(y + f.dy, G.Y), zz = sn_wrap<NEAR>(z + f.dz, G.nz);
            const float4 c = lat[sn_pidx(G, xx, yy, zz)];
            const double fe = (double)a.x * c.x + (double)a.y * c.y + (double)a.z * c.z;
            double afe = fe;
            if (f.r2 > 0) {
                const double na = (double)f.dx * a.x + (double)f.dy * a.y + (double)f.dz * a.z;
                const double nc = (double)f.dx * c.x + (double)f.dy * c.y + (double)f.dz * c.z;
                afe = fe - 3.0 * na * nc / (double)f.r2;
            } else afe = fe - 3.0 * 0.0;       // d forced to 1, n = 0 (analysis.c:571-573)
            v[0] += fe; v[1] += afe;
        }
        if (e > first[b]) {                    // uniform across the block
            sn_block_sum<2, 256>(v, sm);
            if (threadIdx.x == 0) { out[((long long)blockIdx.x * nbins + b) * 2] = v[0]; out[((long long)blockIdx.x * nbins + b) * 2 + 1] = v[1]; }
        } else if (threadIdx.x == 0) { out[((long long)blockIdx.x * nbins + b) * 2] = 0.0; out[((long long)blockIdx.x * nbins + b) * 2 + 1] = 0.0; }
    }
}

// ---- electrostatic potential map (analysis.c:65-94) ---------------------------
struct SnPotOffset { short dx, dy, dz, pad; double w; };   // w = 1/d^3

template <bool NEAR>
__global__ void __launch_bounds__(128) sn_potential_kernel(const float4 *__restrict__ lat, const SnGeom G, const SnPotOffset *__restrict__ off,
                                                           int noff, double *__restrict__ V)
{
    const long long n = (long long)G.X * G.Y * G.nz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, z; sn_site_of(G, i, x, y, z);
    double pot = 0.0;
    for (int o = 0; o < noff; o++) {
        const SnPotOffset f = off[o];
        const int xx = sn_wrap<NEAR>(x + f.dx, G.X), yy = sn_wrap<NEAR>(y + f.dy, G.Y), zz = sn_wrap<NEAR>(z + f.dz, G.nz);
        const float4 c = lat[sn_pidx(G, xx, yy, zz)];
        pot += (double)c.w * ((double)c.x * f.dx + (double)c.y * f.dy + (double)c.z * f.dz) * f.w;
    }
    V[i] = pot;
}

// ---- dipole electric-field maps (analysis.c:310-376, 393-465) -------------------
struct SnEfOffset { short dx, dy, dz, pad; double nx, ny, nz, w; };   // n = r/d, w = 1/d^3

template <bool NEAR>
__global__ void __launch_bounds__(128) sn_efield_kernel(const float4 *__restrict__ lat, const SnGeom G, const SnEfOffset *__restrict__ off,
                                                        int noff, int self_term, double *__restrict__ Emag)
{
    const long long n = (long long)G.X * G.Y * G.nz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, z; sn_site_of(G, i, x, y, z);
    double ex = 0.0, ey = 0.0, ez = 0.0;
    for (int o = 0; o < noff; o++) {
        const SnEfOffset f = off[o];
        const int xx = sn_wrap<NEAR>(x + f.dx, G.X), yy = sn_wrap<NEAR>(y + f.dy, G.Y), zz = sn_wrap<NEAR>(z + f.dz, G.nz);
        const float4 c = lat[sn_pidx(G, xx, yy, zz)];
        const double radial = f.nx * c.x + f.ny * c.y + f.nz * c.z;           // species length not applied (analysis.c:429-434)
        ex += (3.0 * f.nx * radial - c.x) * f.w;
        ey += (3.0 * f.ny * radial - c.y) * f.w;
        ez += (3.0 * f.nz * radial - c.z) * f.w;
    }
    if (self_term) {                                                          // analysis.c:457-459
        const float4 c = lat[sn_pidx(G, x, y, z)];
        ex -= c.x / 3.0; ey -= c.y / 3.0; ez -= c.z / 3.0;
    }
    Emag[i] = sqrt(ex * ex + ey * ey + ez * ez);
}

// ---- recombination model (analysis.c:96-170) ------------------------------------
// One pass over the potential map: partial sums of exp(-bV), exp(bV), f_e = 1/(exp(bV)+1), f_h = 1/(exp(-bV)+1)
// and f_e f_h, and the maxima of f_e, f_h, f_e f_h over the z = 0 plane -> out[block][8].  The
// normalisations by Z_FDe, Z_FDh are applied on the host.
__global__ void __launch_bounds__(256) sn_recombination_kernel(const double *__restrict__ V, long long n, int nz, double scale,
                                                               double *__restrict__ out)
{
    __shared__ double sm[5 * 8];
    __shared__ double smax[3 * 8];
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, m[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = scale * V[i];                                        // pot * BETA
        const double ep = exp(a), em = exp(-a);
        const double fe = 1.0 / (ep + 1.0), fh = 1.0 / (em + 1.0);
        v[0] += em; v[1] += ep; v[2] += fe; v[3] += fh; v[4] += fe * fh;
        if (i % nz == 0) { m[0] = fmax(m[0], fe); m[1] = fmax(m[1], fh); m[2] = fmax(m[2], fe * fh); }
    }
    sn_block_sum<5, 256>(v, sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        if (lane == 0) smax[k * 8 + warp] = m[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 5; k++) out[8 * blockIdx.x + k] = v[k];
        for (int k = 0; k < 3; k++) { double t = 0.0; for (int w = 0; w < 8; w++) t = fmax(t, smax[k * 8 + w]); out[8 * blockIdx.x + 5 + k] = t; }
    }
}
