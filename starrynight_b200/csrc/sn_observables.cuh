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

// ---- shared-memory-tiled stencils ------------------------------------------------------------------
// RDF (radius 9), potential (radius 6) and E-field (radius <= 6) walk hundreds to thousands of neighbours per site.
// A CTA owns an 8^3 block of sites, stages the block plus its halo in shared memory once (every site of the box is
// then read ~500-1500 times from there instead of through L1/L2), and each thread walks the offset list for 2 sites.
// The loader resolves the periodic wrap (any extent, also smaller than the radius: images repeat, as the
// reference's % arithmetic does) and, for a Z-slab handle, reads planes beyond its own from the neighbouring
// GPUs' lattices over NVLink -- the observables of a decomposed lattice need no gather.

// Where a replica's sites live: the handle's own array, the slab neighbours' (same geometry), in the canonical
// padded layout or the tiled kernel's split one.
struct SnLatView {
    const float4 *own, *lo, *hi;        // replica bases; lo / hi = own for a handle that owns the whole Z axis
    int tiled;
};

__device__ __forceinline__ float4 sn_view_site(const SnLatView &v, const SnGeom &G, int x, int y, int z)
{
    const float4 *b = v.own;            // x, y already wrapped into the lattice; z in [-nz, 2 nz)
    if (z < 0) { b = v.lo; z += G.nz; } else if (z >= G.nz) { b = v.hi; z -= G.nz; }
    return v.tiled ? sn_ld2(b, G, sn_pidx2(G, x, y, z)) : b[sn_pidx(G, x, y, z)];
}

namespace sno {
constexpr int T = 8;                     // tile edge
constexpr int THREADS = 256;             // each thread: sites t and t + 256 of the tile (x and x + 4)
__host__ __device__ inline int tiles(int n) { return (n + T - 1) / T; }
__device__ __forceinline__ int wrap(int v, int n) { v %= n; return v < 0 ? v + n : v; }
}

// Stage the box [bx0, bx0+nx) x [by0, by0+ny) x [bz0, bz0+nzb) (lattice coordinates before wrapping) through
// put(cell, float4), cell = (lx * ny + ly) * nzb + lz.  One (x, y) row per warp at a time, the lanes along z (the
// contiguous direction of the lattice): two integer divisions per row instead of per cell.
template <class Put>
__device__ __forceinline__ void sn_load_box(const SnLatView &view, const SnGeom &G, int bx0, int by0, int bz0, int nx, int ny, int nzb, int pitch, Put &&put)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int row = warp; row < nx * ny; row += nwarps) {
        const int lx = row / ny, ly = row - lx * ny;
        const int x = sno::wrap(bx0 + lx, G.X), y = sno::wrap(by0 + ly, G.Y);
        for (int lz = lane; lz < nzb; lz += 32) {
            int z = bz0 + lz;
            if (G.periodic_z) z = sno::wrap(z, G.nz);
            put(row * pitch + lz, sn_view_site(view, G, x, y, z));
        }
    }
}

// Site of the 8^3 tile owned by thread slot t (0..511) in the kernels that keep the box as FP64 component arrays.
// Eight lanes run along z (64 B of one row); the two rows of a half-warp must fall into complementary halves of the 32
// banks, i.e. lie `step` rows apart with (step * pitch * 8) % 128 == 64 (sn_obs_layout picks pitch and step): the
// 8-byte loads of a warp then cost the minimum of two wavefronts instead of four.
__device__ __forceinline__ void sn_obs_site(int t, int step, int &lx, int &ly, int &lz)
{
    lz = t & 7; lx = t >> 6;
    const int a = (t >> 3) & 1, r = (t >> 4) & 3;
    ly = step == 1 ? a + 2 * r : step == 2 ? 2 * a + (r & 1) + 4 * (r >> 1) : 4 * a + r;
}

// z pitch (in cells) and row step for a box of n cells per axis
inline void sn_obs_layout(int n, int *pitch, int *step)
{
    for (int p = n; p <= n + 2; p++)
        for (int s = 1; s <= 4; s *= 2)
            if ((s * p * 8) % 128 == 64) { *pitch = p; *step = s; return; }
    *pitch = n; *step = 1;                        // not reached for even n
}

__device__ __forceinline__ void sn_tile_origin(const SnGeom &G, int &x0, int &y0, int &z0)
{
    const int tz = sno::tiles(G.nz), ty = sno::tiles(G.Y);
    const int b = blockIdx.x;
    z0 = (b % tz) * sno::T; y0 = ((b / tz) % ty) * sno::T; x0 = (b / (tz * ty)) * sno::T;
}

// sum of nrows rows of nv doubles, one block per column, fixed order: bit-reproducible
__global__ void __launch_bounds__(256) sn_reduce_rows_kernel(const double *__restrict__ in, long long nrows, int nv, double *__restrict__ out)
{
    __shared__ double sm[8];
    double v[1] = {0.0};
    for (long long r = threadIdx.x; r < nrows; r += 256) v[0] += in[r * nv + blockIdx.x];
    sn_block_sum<1, 256>(v, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = v[0];
}

// ---- radial order parameter (analysis.c:528-598) ------------------------------
// Offsets inside the radius-9 sphere are sorted by r^2 on the host; `first[b]` is the first offset of bin b.  Only
// one of every pair {d, -d} is walked: on a periodic lattice the pairs (i, i+d) over all sites i are the pairs
// (j-d, j) over all j, and both correlations are symmetric in their two dipoles, so the host doubles the sums of
// r^2 > 0 (half the 3071 pair terms per site).  The half space walked is dx > 0, or dx = 0 and dy > 0, or
// dx = dy = 0 and dz >= 0, so the box is 17 x 26 x 26 sites (184 KB as float4).  FP64 throughout, block-reduced
// once per bin -> out[block][bin][2]; sn_reduce_rows_kernel adds the blocks.
struct SnRdfOffset { double dx, dy, dz; int delta; int r2; };       // delta: box index step (32 B per entry)
namespace sno { constexpr int RDF_R = 9, RDF_NX = T + RDF_R, RDF_NY = T + 2 * RDF_R, RDF_NZ = T + 2 * RDF_R;
                constexpr int RDF_SMEM = RDF_NX * RDF_NY * RDF_NZ * 16, RDF_NOFF = 1488, RDF_NBINS = 81; }   // 1485 lattice vectors with r^2 <= 80 in the half space
// The offset table is the same for every lattice (radius 9, fixed box): constant memory, read with a warp-uniform
// index -- no global-memory latency in the pair loop (the first tiled version spent 37 % of its stall samples there).
__constant__ SnRdfOffset sn_c_rdf_off[sno::RDF_NOFF];
__constant__ int sn_c_rdf_first[sno::RDF_NBINS + 1];

// Per pair term and site: FE += a.c (3 DFMA straight into the accumulator), S += (d.a)(d.c) (7 DFMA); the AFE sum of a
// bin is FE - (3 / r^2) S with the factor applied once per bin (n = d / |d|, analysis.c:571-576; the origin has n = 0).
__global__ void __launch_bounds__(sno::THREADS, 1) sn_rdf_tiled_kernel(const SnLatView view, const SnGeom G, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char sn_obs_smem[];
    float4 *box = reinterpret_cast<float4 *>(sn_obs_smem);
    __shared__ double sm[2 * 8];
    int x0, y0, z0; sn_tile_origin(G, x0, y0, z0);
    sn_load_box(view, G, x0, y0 - sno::RDF_R, z0 - sno::RDF_R, sno::RDF_NX, sno::RDF_NY, sno::RDF_NZ, sno::RDF_NZ, [&](int i, float4 v) { box[i] = v; });
    __syncthreads();
    int base[2]; bool live[2]; double ax[2], ay[2], az[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int t = threadIdx.x + s * sno::THREADS, lz = t & 7, ly = (t >> 3) & 7, lx = t >> 6;
        live[s] = x0 + lx < G.X && y0 + ly < G.Y && z0 + lz < G.nz;
        base[s] = (lx * sno::RDF_NY + ly + sno::RDF_R) * sno::RDF_NZ + lz + sno::RDF_R;
        const float4 a = box[base[s]];
        ax[s] = live[s] ? (double)a.x : 0.0; ay[s] = live[s] ? (double)a.y : 0.0; az[s] = live[s] ? (double)a.z : 0.0;   // a dead site adds exact zeros
    }
    for (int b = 0; b < sno::RDF_NBINS; b++) {
        const int o0 = sn_c_rdf_first[b], o1 = sn_c_rdf_first[b + 1];
        double fe[2] = {0.0, 0.0}, sq[2] = {0.0, 0.0};
#pragma unroll 4
        for (int o = o0; o < o1; o++) {
            const double dx = sn_c_rdf_off[o].dx, dy = sn_c_rdf_off[o].dy, dz = sn_c_rdf_off[o].dz;
            const int delta = sn_c_rdf_off[o].delta;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const float4 c = box[base[s] + delta];
                const double cx = c.x, cy = c.y, cz = c.z;
                fe[s] = fma(ax[s], cx, fma(ay[s], cy, fma(az[s], cz, fe[s])));
                const double na = fma(dx, ax[s], fma(dy, ay[s], dz * az[s]));
                const double nc = fma(dx, cx, fma(dy, cy, dz * cz));
                sq[s] = fma(na, nc, sq[s]);
            }
        }
        double v[2];
        v[0] = fe[0] + fe[1];
        v[1] = v[0] - (b > 0 ? 3.0 / (double)b : 0.0) * (sq[0] + sq[1]);
        if (o1 > o0) sn_block_sum<2, sno::THREADS>(v, sm);      // uniform across the block
        if (threadIdx.x == 0) { out[((long long)blockIdx.x * sno::RDF_NBINS + b) * 2] = v[0]; out[((long long)blockIdx.x * sno::RDF_NBINS + b) * 2 + 1] = v[1]; }
    }
}

// ---- electrostatic potential map (analysis.c:65-94) ---------------------------
// V_i = sum_{0 < d <= 6} l_j (p_j . r) / d^3.  The box holds the moments m_j = l_j p_j as doubles (the product of two
// floats is exact in double), component-wise arrays so that consecutive lanes read consecutive words; the offset
// table holds r / d^3.  3 FMA and 3 shared-memory loads per pair term.
struct SnPotOffset { double kx, ky, kz; int delta; int pad; };      // 32 B: two broadcast LDS.128 per term
namespace sno { constexpr int POT_R = 6, POT_N = T + 2 * POT_R, POT_MAXOFF = 1024; }

__global__ void __launch_bounds__(sno::THREADS, 1) sn_potential_tiled_kernel(const SnLatView view, const SnGeom G, const SnPotOffset *__restrict__ off,
                                                                             int noff, int pitch, int step, double *__restrict__ V)
{
    extern __shared__ __align__(16) unsigned char sn_obs_smem[];
    const int cells = sno::POT_N * sno::POT_N * pitch;
    double *mx = reinterpret_cast<double *>(sn_obs_smem), *my = mx + cells, *mz = my + cells;
    SnPotOffset *tab = reinterpret_cast<SnPotOffset *>(mz + cells);
    int x0, y0, z0; sn_tile_origin(G, x0, y0, z0);
    for (int i = threadIdx.x; i < noff; i += sno::THREADS) tab[i] = off[i];
    sn_load_box(view, G, x0 - sno::POT_R, y0 - sno::POT_R, z0 - sno::POT_R, sno::POT_N, sno::POT_N, sno::POT_N, pitch,
                [&](int i, float4 v) { mx[i] = (double)v.w * v.x; my[i] = (double)v.w * v.y; mz[i] = (double)v.w * v.z; });
    __syncthreads();
    int base[2]; double pot[2] = {0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int lx, ly, lz; sn_obs_site(threadIdx.x + s * sno::THREADS, step, lx, ly, lz);
        base[s] = ((lx + sno::POT_R) * sno::POT_N + ly + sno::POT_R) * pitch + lz + sno::POT_R;
    }
#pragma unroll 4
    for (int o = 0; o < noff; o++) {
        const SnPotOffset f = tab[o];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int c = base[s] + f.delta;
            pot[s] = fma(mx[c], f.kx, fma(my[c], f.ky, fma(mz[c], f.kz, pot[s])));
        }
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int lx, ly, lz; sn_obs_site(threadIdx.x + s * sno::THREADS, step, lx, ly, lz);
        if (x0 + lx < G.X && y0 + ly < G.Y && z0 + lz < G.nz) V[((long long)(x0 + lx) * G.Y + y0 + ly) * G.nz + z0 + lz] = pot[s];
    }
}

// ---- dipole electric-field maps, tiled (analysis.c:310-376, 393-465); reach <= 6 --------------------
struct __align__(16) SnEfOffset2 { double nx, ny, nz, w; int delta; int pad[3]; };    // n = r / d, w = 1 / d^3; 48 B

__global__ void __launch_bounds__(sno::THREADS, 1) sn_efield_tiled_kernel(const SnLatView view, const SnGeom G, const SnEfOffset2 *__restrict__ off,
                                                                          int noff, int reach, int pitch, int step, int self_term, double *__restrict__ Emag)
{
    extern __shared__ __align__(16) unsigned char sn_obs_smem[];
    const int N = sno::T + 2 * reach, cells = N * N * pitch;
    double *px = reinterpret_cast<double *>(sn_obs_smem), *py = px + cells, *pz = py + cells;
    SnEfOffset2 *tab = reinterpret_cast<SnEfOffset2 *>(pz + cells);
    int x0, y0, z0; sn_tile_origin(G, x0, y0, z0);
    for (int i = threadIdx.x; i < noff; i += sno::THREADS) tab[i] = off[i];
    sn_load_box(view, G, x0 - reach, y0 - reach, z0 - reach, N, N, N, pitch, [&](int i, float4 v) { px[i] = v.x; py[i] = v.y; pz[i] = v.z; });
    __syncthreads();
    int base[2]; double ex[2] = {0.0, 0.0}, ey[2] = {0.0, 0.0}, ez[2] = {0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int lx, ly, lz; sn_obs_site(threadIdx.x + s * sno::THREADS, step, lx, ly, lz);
        base[s] = ((lx + reach) * N + ly + reach) * pitch + lz + reach;
    }
#pragma unroll 2
    for (int o = 0; o < noff; o++) {
        const SnEfOffset2 f = tab[o];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int c = base[s] + f.delta;
            const double cx = px[c], cy = py[c], cz = pz[c];
            const double radial = f.nx * cx + f.ny * cy + f.nz * cz;          // species length not applied (analysis.c:429-434)
            ex[s] += (3.0 * f.nx * radial - cx) * f.w;
            ey[s] += (3.0 * f.ny * radial - cy) * f.w;
            ez[s] += (3.0 * f.nz * radial - cz) * f.w;
        }
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int lx, ly, lz; sn_obs_site(threadIdx.x + s * sno::THREADS, step, lx, ly, lz);
        if (!(x0 + lx < G.X && y0 + ly < G.Y && z0 + lz < G.nz)) continue;
        if (self_term) { ex[s] -= px[base[s]] / 3.0; ey[s] -= py[base[s]] / 3.0; ez[s] -= pz[base[s]] / 3.0; }   // analysis.c:457-459
        Emag[((long long)(x0 + lx) * G.Y + y0 + ly) * G.nz + z0 + lz] = sqrt(ex[s] * ex[s] + ey[s] * ey[s] + ez[s] * ez[s]);
    }
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
__global__ void __launch_bounds__(256) sn_recombination_kernel(const double *__restrict__ V, long long n, int nz, int has_z0, double scale,
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
        if (has_z0 && i % nz == 0) { m[0] = fmax(m[0], fe); m[1] = fmax(m[1], fh); m[2] = fmax(m[2], fe * fh); }
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
