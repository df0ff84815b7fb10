// sn_common.cuh -- shared definitions for libstarrynight_b200.so (sm_100a only)
//
// Device lattice layout ("padded AoS"): one float4 (x, y, z, length) per site --
// the reference's `struct dipole` (config.c:32-36) -- in an array of
// (X+2g) x (Y+2g) x (nz+2g) float4 per replica, z fastest, where g = cutoff is a
// ghost shell that mirrors the periodic images (and, for a Z-slab handle, the
// neighbouring GPUs' boundary planes).  Every kernel that walks the cut-off
// sphere therefore addresses neighbours as base + constant offset, with none of
// the reference's `(X+x+dx)%X` arithmetic (montecarlo-core.c:97).  Whoever
// changes a site within g of a face also writes its ghost images.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/starrynight_b200.h"

#define SN_FLAGS_NEXT 32        // h->flags[32..33]: the tiled kernel's work counter (u64)
#define SN_FLAGS_SPECIES 40     // h->flags[40]: result of the species scan in sn_set_lattice
#define SN_FLAGS_ERR 44         // h->flags[44]: set by a device-side wait that ran out of time (a slab neighbour never arrived)
#define SN_FLAGS_DESC 48        // h->flags[48..55]: slab descriptor (magic, X, Y, nz, replicas, cutoff, tiled, Z) checked by sn_ipc_attach
#define SN_DESC_WORDS 8
#define SN_FLAGS_VER 64         // h->flags[64..]: tile versions, [rep][X/16][Y/16][nz/16 + 2]
#define SN_MAX_NB 1024          // neighbour-table capacity in constant memory (cutoff <= 6)

struct SnGeom {
    int X, Y, Z;                // global lattice
    int z0, nz;                 // this handle's slab
    int g;                      // ghost width in x and y (= cutoff)
    int gz;                     // ghost width in z (0 when Z == 1: the axis does not interact)
    int PY, PZ;                 // padded extents of y and z
    long long sx, sy;           // strides (in sites) of x and y in the padded array
    long long rep_stride;       // sites per replica in the padded array
    int periodic_z;             // nz == Z: z ghosts are periodic images of this handle's own planes
};

// Device-side waits (slab handshake, tile dependencies) are bounded: a peer that never launches must not hang the
// node.  When the bound passes the waiter raises SN_FLAGS_ERR and carries on; every sweep kernel of the handle then
// drains without taking new work and the next host call that synchronises returns SN_ERR_CUDA.
__device__ __forceinline__ unsigned long long sn_globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__host__ __device__ inline long long sn_pidx(const SnGeom &G, int x, int y, int z)
{
    return ((long long)(x + G.g) * G.PY + (y + G.g)) * G.PZ + (z + G.gz);
}

// Layout of the tiled kernel's copy ("split layout"): the same padded (x, y) grid, z padded to PZ2 = nz + 8 planes
// (plane z lives at row position z + 4: a tile's 28-plane window z0-4 .. z0+23 starts at an even, 16-byte aligned
// position), and the four floats of a site split over three arrays per replica:
//     xy  float2[N2]   at float offset 0          (x, y)
//     z   float [N2]   at float offset 2 N2
//     len float [N2]   at float offset 3 N2       (never written by a sweep)
// N2 = PX * PY * PZ2 cells, i.e. 16 N2 bytes per replica -- the size of the float4 array it replaces, so the buffer
// stays a float4 * whose replica stride is N2.  The tiled kernel loads two consecutive planes of a column with one
// LDS.128 (xy) + one LDS.64 (z): 12 bytes per site through the shared-memory pipe instead of 16, and nothing at
// all for the lengths when every site has length 1.  One TMA row is 28 planes = 224 B (xy) / 112 B (z, len).
__host__ __device__ inline int sn_pz2(const SnGeom &G) { return G.nz + 8; }
__host__ __device__ inline long long sn_rep_stride2(const SnGeom &G) { return (long long)(G.X + 2 * G.g) * G.PY * sn_pz2(G); }   // cells = float4 units
__host__ __device__ inline long long sn_pidx2(const SnGeom &G, int x, int y, int z)
{
    return ((long long)(x + G.g) * G.PY + (y + G.g)) * sn_pz2(G) + (z + 4);
}
__host__ __device__ inline float4 sn_ld2(const float4 *rep_base, const SnGeom &G, long long cell)
{
    const long long n2 = sn_rep_stride2(G);
    const float *f = reinterpret_cast<const float *>(rep_base);
    const float2 xy = reinterpret_cast<const float2 *>(f)[cell];
    return make_float4(xy.x, xy.y, f[2 * n2 + cell], f[3 * n2 + cell]);
}
__host__ __device__ inline void sn_st2(float4 *rep_base, const SnGeom &G, long long cell, const float4 v)
{
    const long long n2 = sn_rep_stride2(G);
    float *f = reinterpret_cast<float *>(rep_base);
    reinterpret_cast<float2 *>(f)[cell] = make_float2(v.x, v.y);
    f[2 * n2 + cell] = v.z;
    f[3 * n2 + cell] = v.w;
}

// One colour sublattice per axis.  Period P = cutoff+1; if the extent is not a
// multiple of P the r = extent % P trailing coordinates get colours of their
// own, so same-colour sites are always > cutoff apart, also across the wrap.
struct SnAxisColour {
    int P, n, r, ncol;          // period, extent, remainder, number of colours
};

__host__ __device__ inline SnAxisColour sn_axis_colour(int extent, int cutoff, bool flat)
{
    SnAxisColour a;
    a.n = extent;
    if (flat) { a.P = 1; a.r = 0; a.ncol = 1; return a; }      // axis does not interact (Z==1)
    a.P = cutoff + 1;
    if (extent < a.P) { a.P = extent; a.r = 0; a.ncol = extent; return a; }
    a.r = extent % a.P;
    a.ncol = a.P + a.r;
    return a;
}
__host__ __device__ inline int sn_axis_count(const SnAxisColour &a, int c)
{
    return c < a.P ? (a.n - a.r) / a.P : 1;
}
__host__ __device__ inline int sn_axis_coord(const SnAxisColour &a, int c, int i)
{
    return c < a.P ? c + a.P * i : a.n - a.r + (c - a.P);
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based per-site RNG --------
// replaces the global MT19937 stream (mt19937ar-cok.c) the reference draws from
// in MC_move (montecarlo-core.c:159-161,179) and random_sphere_point (config.c:209-210).
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 sn_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 24-bit uniform on [0,1)
__host__ __device__ inline float sn_u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
// The accept test's uniform: all 32 bits, rounded to the nearest float -- the resolution of the reference's
// genrand_real2 (montecarlo-core.c:179), so acceptance probabilities down to 2^-32 are taken at their rate (a
// 24-bit uniform would cut them off at 6e-8: low temperatures).  Like genrand_real2 it can return 0 (always accept);
// values that round up to 1.0f never accept a dE >= 0, as u -> 1 should.
__host__ __device__ inline float sn_u01_32(uint32_t r) { return (float)r * (1.0f / 4294967296.0f); }

#define SN_CUDA_CHECK(call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return sn_fail(SN_ERR_CUDA, "%s: %s (%s:%d)", #call,           \
                                              cudaGetErrorString(e_), __FILE__, __LINE__);    \
    } while (0)

int sn_fail(int code, const char *fmt, ...);

// neighbour entry for the table-driven (any cutoff) path
struct SnNbEntry {
    int dx, dy, dz, nn;                 // nn = 1 for |r| == 1 (cage-strain neighbours)
    float txx, tyy, tzz, txy, txz, tyz; // (delta_ab - 3 n_a n_b) / d^3
};

struct sn_handle {
    sn_params p;
    SnGeom G;
    float4 *lat = nullptr;              // device, nreplicas * rep_stride
    float *beta = nullptr;              // device, per replica
    float4 *efield = nullptr;           // device, per replica
    unsigned long long *counters = nullptr;   // device, per replica {accept, reject, vacant}
    uint4 *rep_key = nullptr;           // device, per replica Philox key + counter tag
    unsigned long long sweep = 0;       // sweeps done so far (Philox counter word)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};   // sn_mc_sweeps_timed
    cudaEvent_t ev_sweeps = nullptr;          // recorded behind every sn_mc_sweeps (sn_order_after)
    int nnb = 0;
    std::vector<int> nb_dxyz;           // reference order (montecarlo-core.c:47-62)
    std::vector<float> nb_d;
    SnNbEntry *nb_table = nullptr;      // device copy for the table-driven kernels
    int *d_nb_dxyz = nullptr;           // device copy of nb_dxyz for the exact-order audit kernels
    std::vector<float> h_beta;          // host mirrors of the per-replica couplings
    std::vector<float> h_efield;        // 3 per replica
    std::vector<double> h_cage;         // CageStrain per replica, as given (device copy, rounded to float: efield[rep].w)
    bool species = true;                // false when every length is exactly 1 (skips the l_j multiplies)
    std::vector<char> rep_species;      // per replica: some length != 1
    unsigned int *rep_species_dev = nullptr;   // device: raised by the upload kernel, read lazily (sn_resolve_species)
    bool species_dirty = false;
    // couplings (beta, field, cage strain) travel through a ring of pinned slots: the copies are stream-ordered and the
    // calls return at once (a field ramp with one sweep per point must not synchronise the stream at every point)
    float4 *coupling_ring = nullptr; unsigned long long ring_pos = 0;
    unsigned int *rep_species_host = nullptr;  // pinned mirror of the flags, filled behind every upload ...
    cudaEvent_t ev_species = nullptr;          // ... and complete when this event is: sn_resolve_species waits for the UPLOAD only, not for the stream
    bool use_tiled = false;
    bool use_resident = false;          // lattice small enough to live in one CTA's shared memory (sn_sweep_resident.cuh)
    // The tiled kernel works on a second copy of the lattice in the split layout (sn_pidx2 / sn_ld2 / sn_st2).
    // `lat` (canonical) and `lat2` are synchronised lazily: whoever needs one of them converts from the other
    // if it is stale.
    float4 *lat2 = nullptr;
    bool lat_valid = true, lat2_valid = false;
    // slab wiring
    float4 *peer_lat[2] = {nullptr, nullptr};       // lower / upper neighbour's padded lattice
    unsigned int *flags = nullptr;                  // own phase flags [0,1], work counter, tile versions (device; SN_FLAGS_*)
    unsigned int *peer_flags[2] = {nullptr, nullptr};
    size_t nver = 0;                                // entries of the tile-version array behind SN_FLAGS_VER
    unsigned int phase_epoch = 0;
    bool peer_is_ipc[2] = {false, false};
    unsigned char ipc_key[2][64] = {};
    unsigned long long spin_timeout_ns = 60ull * 1000000000ull;   // bound of device-side waits (env SN_SPIN_TIMEOUT_S)
    int num_sms = 148;
    int grid_limit = 0;                 // > 0: cap on the tiled kernel's persistent grid (slab neighbours sharing this device)
    const void *serial_group = nullptr; // handles whose sweeps the caller serialises (sn_order_after) share one group: they never run at once
    void *tmap = nullptr;               // CUtensorMap storage for the tiled kernel (device-constant copy made at launch)
    float *audit_dev = nullptr;         // set for the duration of sn_mc_sweep_audit: the sweep kernels record every attempt
    // scratch
    double *d_scratch = nullptr; size_t scratch_bytes = 0;
    void *staging = nullptr; size_t staging_bytes = 0;
};

// entry points implemented across translation units
int sn_sweep_colour_launch(sn_handle *h, long long nsweeps, long long *launches);
int sn_sweep_tiled_launch(sn_handle *h, long long nsweeps, long long *launches);
int sn_sweep_resident_launch(sn_handle *h, long long nsweeps, long long *launches);
bool sn_tiled_supported(const sn_handle *h, std::string *why);
int sn_tiled_prepare(sn_handle *h);
void sn_tiled_release(sn_handle *h);
int sn_refresh_ghosts(sn_handle *h);
int sn_sync_canonical(sn_handle *h);
int sn_convert_layout(sn_handle *h, bool to_tiled);
int sn_resolve_species(sn_handle *h);
int sn_energy_exact_launch(sn_handle *h, int replica, int precision, int n, const int *d_sites,
                           const float *d_newdip, double *d_out);
int sn_energy_exact_map_launch(sn_handle *h, int replica, int precision, int which, double *d_out);
int sn_scratch(sn_handle *h, size_t bytes, void **out);
int sn_energy_exact_preload();
int sn_check_device_error(sn_handle *h);
