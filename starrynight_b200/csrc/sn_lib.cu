// sn_lib.cu -- C ABI of libstarrynight_b200.so (see include/starrynight_b200.h)
//
// Host runtime around the sm_100a kernels: owns the padded device lattice, the
// per-replica couplings, the Philox key/counter, the stream, and dispatches
// sn_mc_sweeps to the colour-pass kernel (sn_sweep_colour.cuh) or the
// TMA/shared-memory tile kernel (sn_sweep_tiled.cuh).  There is no CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "sn_common.cuh"
#include "sn_field.cuh"
#include "sn_observables.cuh"
#include "sn_sweep_colour.cuh"
#include "sn_sweep_tiled.cuh"
#include "sn_sweep_resident.cuh"

static thread_local char sn_err[512] = "";
static int sn_slab_phase_sync(sn_handle *h, long long *launches);

int sn_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(sn_err, sizeof sn_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *sn_last_error(void) { return sn_err; }
extern "C" const char *sn_version(void) { return "starrynight_b200 0.1 (sm_100a)"; }

extern "C" int sn_device_count(int *n)
{
    if (!n) return sn_fail(SN_ERR_INVALID, "sn_device_count: null");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
    *n = c;
    return SN_OK;
}

extern "C" int sn_default_params(sn_params *p)
{
    if (!p) return sn_fail(SN_ERR_INVALID, "sn_default_params: null");
    memset(p, 0, sizeof *p);
    p->X = 20; p->Y = 20; p->Z = 20;          // config.c:12-14
    p->cutoff = 3;                            // config.c:70
    p->CageStrain = 1.0; p->K = 1.0;          // config.c:66-68
    p->beta = 1.0;                            // config.c:62
    p->ConstrainToX = 0; p->DIM = 3;          // config.c:85,16
    p->nreplicas = 1;
    p->seed = 0xDEADBEEFull + 300;            // main.c:172
    p->device = 0; p->z0 = 0; p->nz = 0;
    p->kernel = SN_KERNEL_AUTO;
    return SN_OK;
}

// Did a device-side wait of this handle run out of time?  (call after a stream synchronisation)
int sn_check_device_error(sn_handle *h)
{
    unsigned int e = 0;
    SN_CUDA_CHECK(cudaMemcpyAsync(&e, h->flags + SN_FLAGS_ERR, sizeof e, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (e) return sn_fail(SN_ERR_CUDA, "a device-side wait timed out after %.1f s (%s): a Z-slab neighbour never ran its part of the sweep "
                                       "(results of this handle are invalid)", h->spin_timeout_ns * 1e-9,
                          e == 2 ? "tile dependencies of the sweep kernel" : "slab handshake");
    return SN_OK;
}

int sn_scratch(sn_handle *h, size_t bytes, void **out)
{
    if (bytes > h->scratch_bytes) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr; h->scratch_bytes = 0;
        SN_CUDA_CHECK(cudaMalloc(&h->d_scratch, bytes));
        h->scratch_bytes = bytes;
    }
    *out = h->d_scratch;
    return SN_OK;
}

// gen_neighbour(), montecarlo-core.c:38-72: order dx -> dy -> dz, 0 < d <= cutoff,
// d = sqrt in float; ZCutOff = 0 when Z == 1.
static int sn_build_neighbours(sn_handle *h)
{
    const int c = h->p.cutoff, zc = h->p.Z == 1 ? 0 : c;
    h->nb_dxyz.clear(); h->nb_d.clear();
    std::vector<SnNbEntry> tab;
    for (int dx = -c; dx <= c; dx++) for (int dy = -c; dy <= c; dy++) for (int dz = -zc; dz <= zc; dz++) {
        if (!dx && !dy && !dz) continue;
        const float d = (float)sqrt((double)((float)dx * dx + dy * dy + dz * dz));
        if (d > (float)c) continue;
        h->nb_dxyz.push_back(dx); h->nb_dxyz.push_back(dy); h->nb_dxyz.push_back(dz);
        h->nb_d.push_back(d);
        const double r2 = (double)dx * dx + (double)dy * dy + (double)dz * dz, dd = sqrt(r2), i3 = 1.0 / (dd * dd * dd);
        SnNbEntry e;
        e.dx = dx; e.dy = dy; e.dz = dz; e.nn = (dx * dx + dy * dy + dz * dz) == 1;
        e.txx = (float)(i3 - 3.0 * dx * dx * i3 / r2); e.tyy = (float)(i3 - 3.0 * dy * dy * i3 / r2);
        e.tzz = (float)(i3 - 3.0 * dz * dz * i3 / r2);
        e.txy = (float)(-3.0 * dx * dy * i3 / r2); e.txz = (float)(-3.0 * dx * dz * i3 / r2); e.tyz = (float)(-3.0 * dy * dz * i3 / r2);
        tab.push_back(e);
    }
    h->nnb = (int)h->nb_d.size();
    SN_CUDA_CHECK(cudaMalloc(&h->nb_table, sizeof(SnNbEntry) * std::max(1, h->nnb)));
    SN_CUDA_CHECK(cudaMemcpy(h->nb_table, tab.data(), sizeof(SnNbEntry) * h->nnb, cudaMemcpyHostToDevice));
    SN_CUDA_CHECK(cudaMalloc(&h->d_nb_dxyz, sizeof(int) * 3 * std::max(1, h->nnb)));
    SN_CUDA_CHECK(cudaMemcpy(h->d_nb_dxyz, h->nb_dxyz.data(), sizeof(int) * 3 * h->nnb, cudaMemcpyHostToDevice));
    return SN_OK;
}

// Slab handles of this process, by device.  Slabs of one lattice wait for each other's tile versions from inside
// their persistent sweep kernels (directly or through a chain of neighbours on other devices), so all slab kernels
// that share a device must be resident at once: one CTA per SM, hence each gets an equal share of the SMs.
static std::mutex sn_slab_registry_mutex;
static std::vector<sn_handle *> sn_slab_registry;

// (registry mutex held) SM share of every registered handle: the SMs of a device are divided among the groups of
// handles that can be resident at once.  Handles the caller serialises with sn_order_after form one group.
static void sn_slab_shares()
{
    for (sn_handle *a : sn_slab_registry) {
        std::vector<const void *> groups;
        for (sn_handle *b : sn_slab_registry)
            if (b->p.device == a->p.device && std::find(groups.begin(), groups.end(), b->serial_group) == groups.end()) groups.push_back(b->serial_group);
        const int n = (int)groups.size();
        a->grid_limit = n > 1 ? std::max(1, a->num_sms / n) : 0;
    }
}

static void sn_slab_register(sn_handle *h, bool add)
{
    std::lock_guard<std::mutex> lock(sn_slab_registry_mutex);
    auto it = std::find(sn_slab_registry.begin(), sn_slab_registry.end(), h);
    if (add && it == sn_slab_registry.end()) sn_slab_registry.push_back(h);
    if (!add && it != sn_slab_registry.end()) sn_slab_registry.erase(it);
    sn_slab_shares();
}

// (E_x, E_y, E_z, CageStrain) of one replica -> device (the sweep kernels read the four together)
constexpr unsigned long long SN_RING_SLOTS = 4096;
// next pinned slot of the coupling ring; when the ring wraps, the copies that used the old contents must have run
static int sn_ring_slot(sn_handle *h, float4 **slot)
{
    if (h->ring_pos && h->ring_pos % SN_RING_SLOTS == 0) SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    *slot = h->coupling_ring + h->ring_pos++ % SN_RING_SLOTS;
    return SN_OK;
}

static int sn_push_couplings(sn_handle *h, int r)
{
    float4 *slot; int rc = sn_ring_slot(h, &slot); if (rc) return rc;
    *slot = make_float4(h->h_efield[3 * r], h->h_efield[3 * r + 1], h->h_efield[3 * r + 2], (float)h->h_cage[r]);
    SN_CUDA_CHECK(cudaMemcpyAsync(h->efield + r, slot, sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    return SN_OK;
}

static void sn_slab_descriptor(const sn_handle *h, unsigned int d[SN_DESC_WORDS])
{
    d[0] = 0x534E3232u; d[1] = (unsigned)h->G.X; d[2] = (unsigned)h->G.Y; d[3] = (unsigned)h->G.nz; d[4] = (unsigned)h->p.nreplicas;
    d[5] = (unsigned)h->p.cutoff; d[6] = h->use_tiled ? 1u : 0u; d[7] = (unsigned)h->G.Z;
}

static int sn_mode(const sn_handle *h) { return h->p.cutoff == 3 ? (h->p.Z == 1 ? 1 : 0) : 2; }

static int sn_preload_kernels(int device);

// Everything sn_create does once the handle exists; any failure returns through sn_create, which destroys the
// partially built handle (stream, device buffers, registry entry) -- no early return leaks.
static int sn_create_body(sn_handle *h, const sn_params *p)
{
    SnGeom &G = h->G;
    G.X = p->X; G.Y = p->Y; G.Z = p->Z; G.z0 = h->p.z0; G.nz = h->p.nz;
    G.g = p->cutoff; G.gz = p->Z == 1 ? 0 : p->cutoff;
    G.PY = G.Y + 2 * G.g; G.PZ = G.nz + 2 * G.gz;
    G.sy = G.PZ; G.sx = (long long)G.PY * G.PZ;
    G.rep_stride = (long long)(G.X + 2 * G.g) * G.sx;
    G.periodic_z = (G.nz == G.Z);
    if (!G.periodic_z) {
        const int P = p->cutoff + 1;
        if (G.nz % P || G.z0 % P || G.Z % P || G.nz < p->cutoff)
            return sn_fail(SN_ERR_UNSUPPORTED, "sn_create: Z-slabs need z0, nz and Z to be multiples of cutoff+1 (z0=%d nz=%d Z=%d)", G.z0, G.nz, G.Z);
    }
    cudaDeviceProp prop;
    SN_CUDA_CHECK(cudaGetDeviceProperties(&prop, p->device));
    h->num_sms = prop.multiProcessorCount;
    { int rc = sn_preload_kernels(p->device); if (rc) return rc; }
    if (const char *t = getenv("SN_SPIN_TIMEOUT_S")) { const double v = atof(t); if (v > 0.0) h->spin_timeout_ns = (unsigned long long)(v * 1e9); }
    SN_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    SN_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_sweeps, cudaEventDisableTiming));
    SN_CUDA_CHECK(cudaMalloc(&h->rep_species_dev, sizeof(unsigned int) * p->nreplicas));
    SN_CUDA_CHECK(cudaMemsetAsync(h->rep_species_dev, 0, sizeof(unsigned int) * p->nreplicas, h->stream));
    SN_CUDA_CHECK(cudaMallocHost(&h->rep_species_host, sizeof(unsigned int) * p->nreplicas));
    SN_CUDA_CHECK(cudaMallocHost(&h->coupling_ring, sizeof(float4) * SN_RING_SLOTS));
    SN_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_species, cudaEventDisableTiming));
    const size_t cells = (size_t)G.rep_stride * p->nreplicas;
    if (cudaMalloc(&h->lat, cells * sizeof(float4)) != cudaSuccess) {
        cudaGetLastError();
        return sn_fail(SN_ERR_NOMEM, "sn_create: cannot allocate %.1f MB for the lattice", cells * 16.0 / 1e6);
    }
    SN_CUDA_CHECK(cudaMemsetAsync(h->lat, 0, cells * sizeof(float4), h->stream));
    SN_CUDA_CHECK(cudaMalloc(&h->beta, sizeof(float) * p->nreplicas));
    SN_CUDA_CHECK(cudaMalloc(&h->efield, sizeof(float4) * p->nreplicas));
    SN_CUDA_CHECK(cudaMalloc(&h->rep_key, sizeof(uint4) * p->nreplicas));
    {
        // default streams: one key (the seed) for the handle, the replica index in the counter
        std::vector<uint4> k(p->nreplicas);
        for (int r = 0; r < p->nreplicas; r++) k[r] = make_uint4((uint32_t)p->seed, (uint32_t)(p->seed >> 32), (uint32_t)r << 8, 0u);
        SN_CUDA_CHECK(cudaMemcpy(h->rep_key, k.data(), sizeof(uint4) * p->nreplicas, cudaMemcpyHostToDevice));
    }
    SN_CUDA_CHECK(cudaMalloc(&h->counters, sizeof(unsigned long long) * 3 * p->nreplicas));
    SN_CUDA_CHECK(cudaMemsetAsync(h->counters, 0, sizeof(unsigned long long) * 3 * p->nreplicas, h->stream));
    {
        // phase flags, the tiled kernel's work counter and its tile-version array (one allocation, so that a
        // slab neighbour reaches all of it through one IPC handle)
        size_t nflags = SN_FLAGS_VER;
        h->nver = (size_t)p->nreplicas * ((G.X + 15) / 16) * ((G.Y + 15) / 16) * ((G.nz + 15) / 16 + 2);       // partial tiles count
        nflags += h->nver;
        SN_CUDA_CHECK(cudaMalloc(&h->flags, sizeof(unsigned int) * nflags));
        SN_CUDA_CHECK(cudaMemsetAsync(h->flags, 0, sizeof(unsigned int) * nflags, h->stream));
    }
    h->h_beta.assign(p->nreplicas, (float)p->beta);
    h->h_cage.assign(p->nreplicas, p->CageStrain);
    h->rep_species.assign(p->nreplicas, 0);
    h->species = false;
    h->h_efield.resize(3 * (size_t)p->nreplicas);
    for (int r = 0; r < p->nreplicas; r++)
        for (int k = 0; k < 3; k++) h->h_efield[3 * r + k] = p->Efield[k];
    SN_CUDA_CHECK(cudaMemcpyAsync(h->beta, h->h_beta.data(), sizeof(float) * p->nreplicas, cudaMemcpyHostToDevice, h->stream));
    for (int r = 0; r < p->nreplicas; r++) { int rc = sn_push_couplings(h, r); if (rc) return rc; }
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    { int rc = sn_build_neighbours(h); if (rc) return rc; }
    SN_CUDA_CHECK(cudaGetLastError());

    std::string why;
    const bool can_tile = sn_tiled_supported(h, &why);
    if ((p->kernel == SN_KERNEL_TILED || p->kernel == SN_KERNEL_TILED_PHASED) && !can_tile)
        return sn_fail(SN_ERR_UNSUPPORTED, "sn_create: tiled kernel unavailable: %s", why.c_str());
    {
        std::string why_r;
        const bool can_reside = sn_resident_supported(h, &why_r);
        if (p->kernel == SN_KERNEL_RESIDENT && !can_reside)
            return sn_fail(SN_ERR_UNSUPPORTED, "sn_create: shared-memory-resident kernel unavailable: %s", why_r.c_str());
        // a lattice that fits one CTA's shared memory stays there (replica batches fill the GPU); the tiles take the rest
        h->use_resident = can_reside && (p->kernel == SN_KERNEL_AUTO || p->kernel == SN_KERNEL_RESIDENT);
        h->use_tiled = can_tile && !h->use_resident && p->kernel != SN_KERNEL_COLOUR && p->kernel != SN_KERNEL_RESIDENT;
    }
    if (h->use_tiled) { int rc = sn_tiled_prepare(h); if (rc) return rc; }
    {
        unsigned int d[SN_DESC_WORDS];
        sn_slab_descriptor(h, d);
        SN_CUDA_CHECK(cudaMemcpy(h->flags + SN_FLAGS_DESC, d, sizeof d, cudaMemcpyHostToDevice));
    }
    if (!G.periodic_z) sn_slab_register(h, true);
    return SN_OK;
}

extern "C" int sn_create(const sn_params *p, sn_handle **out)
{
    if (!p || !out) return sn_fail(SN_ERR_INVALID, "sn_create: null argument");
    *out = nullptr;
    if (p->X < 1 || p->Y < 1 || p->Z < 1) return sn_fail(SN_ERR_INVALID, "sn_create: lattice %dx%dx%d", p->X, p->Y, p->Z);
    if (p->cutoff < 0 || p->cutoff > 6) return sn_fail(SN_ERR_UNSUPPORTED, "sn_create: DipoleCutOff %d outside 0..6", p->cutoff);
    if (p->nreplicas < 1) return sn_fail(SN_ERR_INVALID, "sn_create: nreplicas %d", p->nreplicas);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return sn_fail(SN_ERR_CUDA, "sn_create: no CUDA device (this library has no CPU path)");
    if (p->device < 0 || p->device >= ndev) return sn_fail(SN_ERR_INVALID, "sn_create: device %d of %d", p->device, ndev);
    SN_CUDA_CHECK(cudaSetDevice(p->device));

    sn_handle *h = new sn_handle();
    h->p = *p;
    if (h->p.nz <= 0) { h->p.nz = p->Z; h->p.z0 = 0; }
    h->serial_group = h;
    h->G.periodic_z = 1;                              // until the geometry is known: sn_destroy must not touch the slab registry
    if (h->p.z0 < 0 || h->p.z0 + h->p.nz > p->Z) {
        const int z0 = h->p.z0, nz = h->p.nz;
        delete h;
        return sn_fail(SN_ERR_INVALID, "sn_create: slab [%d,%d) outside Z=%d", z0, z0 + nz, p->Z);
    }
    const int rc = sn_create_body(h, p);
    if (rc) { sn_destroy(h); return rc; }             // sn_destroy leaves sn_last_error() untouched
    *out = h;
    return SN_OK;
}

extern "C" int sn_destroy(sn_handle *h)
{
    if (!h) return SN_OK;
    if (!h->G.periodic_z) sn_slab_register(h, false);
    cudaSetDevice(h->p.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    sn_tiled_release(h);
    for (int s = 0; s < 2; s++) if (h->peer_is_ipc[s]) {
        if (h->peer_lat[s] && !(s == 1 && h->peer_lat[1] == h->peer_lat[0])) cudaIpcCloseMemHandle(h->peer_lat[s]);
        if (h->peer_flags[s] && !(s == 1 && h->peer_flags[1] == h->peer_flags[0])) cudaIpcCloseMemHandle(h->peer_flags[s]);
    }
    cudaFree(h->lat); cudaFree(h->beta); cudaFree(h->efield); cudaFree(h->counters); cudaFree(h->flags); cudaFree(h->rep_key);
    cudaFree(h->nb_table); cudaFree(h->d_nb_dxyz); cudaFree(h->d_scratch); cudaFree(h->staging);
    for (int e = 0; e < 2; e++) if (h->ev[e]) cudaEventDestroy(h->ev[e]);
    if (h->ev_sweeps) cudaEventDestroy(h->ev_sweeps);
    cudaFree(h->rep_species_dev);
    if (h->rep_species_host) cudaFreeHost(h->rep_species_host);
    if (h->coupling_ring) cudaFreeHost(h->coupling_ring);
    if (h->ev_species) cudaEventDestroy(h->ev_species);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SN_OK;
}

extern "C" int sn_neighbour_table(sn_handle *h, int *n, int *dxyz, float *d)
{
    if (!h || !n) return sn_fail(SN_ERR_INVALID, "sn_neighbour_table: null");
    *n = h->nnb;
    if (dxyz) memcpy(dxyz, h->nb_dxyz.data(), sizeof(int) * 3 * h->nnb);
    if (d) memcpy(d, h->nb_d.data(), sizeof(float) * h->nnb);
    return SN_OK;
}

#define SN_CHECK_HANDLE(h, rep)                                                                        \
    do {                                                                                               \
        if (!(h)) return sn_fail(SN_ERR_INVALID, "%s: null handle", __func__);                        \
        if ((rep) < 0 || (rep) >= (h)->p.nreplicas) return sn_fail(SN_ERR_INVALID, "%s: replica %d of %d", __func__, (rep), (h)->p.nreplicas); \
        SN_CUDA_CHECK(cudaSetDevice((h)->p.device));                                                   \
    } while (0)

int sn_refresh_ghosts(sn_handle *h)
{
    const long long n = h->G.rep_stride;
    dim3 grid((unsigned)((n + 255) / 256), h->p.nreplicas);
    sn_refresh_ghosts_kernel<<<grid, 256, 0, h->stream>>>(h->lat, h->G);
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

// ---- host <-> device lattice transfer ----------------------------------------------------------------
// The host block is dense, float[X][Y][nz][4] (the reference's lattice[x][y][z]); on the device the lattice is
// ghost-padded and, for the tiled kernel, de-interleaved in z.  A transfer is ONE contiguous copy between the host
// block and a dense device staging buffer (full PCIe rate whatever the slab height) plus ONE kernel that moves
// between staging and the layout the sweep kernel works on -- writing every periodic image on the way in, and
// noting whether the replica carries species (any length != 1).  No strided copies, no separate ghost refresh, no
// round trip through the canonical array, no host synchronisation.
template <bool TILED>
__global__ void __launch_bounds__(256) sn_scatter_kernel(const float4 *__restrict__ staging0, float4 *__restrict__ dst0, const SnGeom G,
                                                         unsigned int *__restrict__ species_flag0, const long long rep_stride)
{
    // blockIdx.y: replica of the batch (staging dense, the destination rep_stride cells apart)
    const float4 *__restrict__ staging = staging0 + (long long)blockIdx.y * G.X * G.Y * G.nz;
    float4 *__restrict__ dst = dst0 + (long long)blockIdx.y * rep_stride;
    unsigned int *__restrict__ species_flag = species_flag0 + blockIdx.y;
    // one thread per padded cell (ghost shell included); z ghosts of a Z-slab handle belong to the neighbours
    // (sn_pull_ghosts / sn_set_ghost) and are left alone
    const long long cells = (long long)(G.X + 2 * G.g) * G.PY * G.PZ;
    bool nonunit = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
        const int zp = (int)(i % G.PZ) - G.gz, yp = (int)((i / G.PZ) % G.PY) - G.g, xp = (int)(i / ((long long)G.PZ * G.PY)) - G.g;
        int xs = xp % G.X; xs += xs < 0 ? G.X : 0;
        int ys = yp % G.Y; ys += ys < 0 ? G.Y : 0;
        int zs = zp;
        if (zp < 0 || zp >= G.nz) {
            if (!G.periodic_z) continue;
            zs = zp % G.nz; zs += zs < 0 ? G.nz : 0;
        }
        const float4 v = staging[((long long)xs * G.Y + ys) * G.nz + zs];
        if (TILED) sn_st2(dst, G, sn_pidx2(G, xp, yp, zp), v); else dst[sn_pidx(G, xp, yp, zp)] = v;
        nonunit |= (xs == xp && ys == yp && zs == zp && v.w != 1.0f);
    }
    if (__any_sync(0xffffffffu, nonunit) && (threadIdx.x & 31) == 0) *species_flag = 1u;
}

template <bool TILED>
__global__ void __launch_bounds__(256) sn_gather_kernel(const float4 *__restrict__ src0, float4 *__restrict__ staging0, const SnGeom G, const long long rep_stride)
{
    const float4 *__restrict__ src = src0 + (long long)blockIdx.y * rep_stride;
    float4 *__restrict__ staging = staging0 + (long long)blockIdx.y * G.X * G.Y * G.nz;
    const long long n = (long long)G.X * G.Y * G.nz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % G.nz), y = (int)((i / G.nz) % G.Y), x = (int)(i / ((long long)G.nz * G.Y));
        staging[i] = TILED ? sn_ld2(src, G, sn_pidx2(G, x, y, z)) : src[sn_pidx(G, x, y, z)];
    }
}

static int sn_staging(sn_handle *h, float4 **out, int count = 1)
{
    const size_t bytes = (size_t)h->G.X * h->G.Y * h->G.nz * sizeof(float4) * (size_t)count;
    if (bytes > h->staging_bytes) {
        if (h->staging) cudaFree(h->staging);
        h->staging = nullptr; h->staging_bytes = 0;
        if (cudaMalloc(&h->staging, bytes) != cudaSuccess) { cudaGetLastError(); return sn_fail(SN_ERR_NOMEM, "cannot allocate %.1f MB of staging memory", bytes / 1e6); }
        h->staging_bytes = bytes;
    }
    *out = (float4 *)h->staging;
    return SN_OK;
}

// The species flags are raised on the device by the upload and copied to a pinned mirror right behind it; the host looks
// at them only when it next has to choose a kernel specialisation (sweep launch), and then waits for that copy alone --
// not for whatever else has been queued on the stream since (a wait for another handle's sweeps, the slab handshake):
// a stream-wide synchronisation here made the double-buffered end-to-end loop on two GPUs 25 % slower.
int sn_resolve_species(sn_handle *h)
{
    if (!h->species_dirty) return SN_OK;
    SN_CUDA_CHECK(cudaEventSynchronize(h->ev_species));
    h->species = false;
    for (int r = 0; r < h->p.nreplicas; r++) { h->rep_species[r] = h->rep_species_host[r] != 0; h->species = h->species || h->rep_species_host[r] != 0; }
    h->species_dirty = false;
    return SN_OK;
}

// `count` consecutive replicas from one dense host block: ONE copy, ONE scatter launch (replica along grid.y)
extern "C" int sn_set_lattices_async(sn_handle *h, int first, int count, const float *xyzlen)
{
    SN_CHECK_HANDLE(h, first);
    if (!xyzlen) return sn_fail(SN_ERR_INVALID, "sn_set_lattice: null buffer");
    if (count < 1 || first + count > h->p.nreplicas) return sn_fail(SN_ERR_INVALID, "sn_set_lattices: replicas [%d, %d) of %d", first, first + count, h->p.nreplicas);
    const SnGeom &G = h->G;
    float4 *stg; int rc = sn_staging(h, &stg, count);
    if (rc) return rc;
    const size_t bytes = (size_t)G.X * G.Y * G.nz * sizeof(float4) * (size_t)count;
    SN_CUDA_CHECK(cudaMemcpyAsync(stg, xyzlen, bytes, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaMemsetAsync(h->rep_species_dev + first, 0, sizeof(unsigned int) * count, h->stream));
    const long long cells = G.rep_stride;
    const int nbx = (int)std::min<long long>((cells + 255) / 256, std::max<long long>(1, (long long)h->num_sms * 16 / count));
    const dim3 grid((unsigned)nbx, (unsigned)count);
    if (h->use_tiled) {
        // straight into the tiled kernel's copy; other replicas that only live in the canonical array come along first
        if (!h->lat2_valid && h->p.nreplicas > count && (rc = sn_convert_layout(h, true))) return rc;
        sn_scatter_kernel<true><<<grid, 256, 0, h->stream>>>(stg, h->lat2 + (long long)first * sn_rep_stride2(G), G, h->rep_species_dev + first, sn_rep_stride2(G));
        h->lat2_valid = true; h->lat_valid = false;
    } else {
        if ((rc = sn_sync_canonical(h))) return rc;
        sn_scatter_kernel<false><<<grid, 256, 0, h->stream>>>(stg, h->lat + (long long)first * G.rep_stride, G, h->rep_species_dev + first, G.rep_stride);
        h->lat_valid = true; h->lat2_valid = false;
    }
    SN_CUDA_CHECK(cudaGetLastError());
    SN_CUDA_CHECK(cudaMemcpyAsync(h->rep_species_host, h->rep_species_dev, sizeof(unsigned int) * h->p.nreplicas, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaEventRecord(h->ev_species, h->stream));
    h->species_dirty = true;
    return SN_OK;
}

extern "C" int sn_set_lattice_async(sn_handle *h, int replica, const float *xyzlen)
{
    return sn_set_lattices_async(h, replica, 1, xyzlen);
}

extern "C" int sn_set_lattice(sn_handle *h, int replica, const float *xyzlen)
{
    int rc = sn_set_lattice_async(h, replica, xyzlen);
    if (rc) return rc;
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));      // the caller's buffer is free again
    return SN_OK;
}

extern "C" int sn_get_lattices_async(sn_handle *h, int first, int count, float *xyzlen)
{
    SN_CHECK_HANDLE(h, first);
    if (!xyzlen) return sn_fail(SN_ERR_INVALID, "sn_get_lattice: null buffer");
    if (count < 1 || first + count > h->p.nreplicas) return sn_fail(SN_ERR_INVALID, "sn_get_lattices: replicas [%d, %d) of %d", first, first + count, h->p.nreplicas);
    const SnGeom &G = h->G;
    float4 *stg; int rc = sn_staging(h, &stg, count);
    if (rc) return rc;
    const long long n = (long long)G.X * G.Y * G.nz;
    const int nbx = (int)std::min<long long>((n + 255) / 256, std::max<long long>(1, (long long)h->num_sms * 16 / count));
    const dim3 grid((unsigned)nbx, (unsigned)count);
    if (h->use_tiled && h->lat2_valid) sn_gather_kernel<true><<<grid, 256, 0, h->stream>>>(h->lat2 + (long long)first * sn_rep_stride2(G), stg, G, sn_rep_stride2(G));
    else {
        if ((rc = sn_sync_canonical(h))) return rc;
        sn_gather_kernel<false><<<grid, 256, 0, h->stream>>>(h->lat + (long long)first * G.rep_stride, stg, G, G.rep_stride);
    }
    SN_CUDA_CHECK(cudaGetLastError());
    SN_CUDA_CHECK(cudaMemcpyAsync(xyzlen, stg, (size_t)n * sizeof(float4) * (size_t)count, cudaMemcpyDeviceToHost, h->stream));
    return SN_OK;
}

extern "C" int sn_get_lattice_async(sn_handle *h, int replica, float *xyzlen)
{
    return sn_get_lattices_async(h, replica, 1, xyzlen);
}

extern "C" int sn_get_lattice(sn_handle *h, int replica, float *xyzlen)
{
    int rc = sn_get_lattice_async(h, replica, xyzlen);
    if (rc) return rc;
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return sn_check_device_error(h);
}

extern "C" int sn_set_beta(sn_handle *h, int replica, double beta)
{
    SN_CHECK_HANDLE(h, replica);
    h->h_beta[replica] = (float)beta;
    float4 *slot; int rc = sn_ring_slot(h, &slot); if (rc) return rc;
    slot->x = (float)beta;
    SN_CUDA_CHECK(cudaMemcpyAsync(h->beta + replica, &slot->x, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    return SN_OK;
}

extern "C" int sn_set_efield(sn_handle *h, int replica, const float E[3])
{
    SN_CHECK_HANDLE(h, replica);
    if (!E) return sn_fail(SN_ERR_INVALID, "sn_set_efield: null");
    for (int k = 0; k < 3; k++) h->h_efield[3 * replica + k] = E[k];
    return sn_push_couplings(h, replica);
}

extern "C" int sn_set_cagestrain(sn_handle *h, double cagestrain)
{
    SN_CHECK_HANDLE(h, 0);
    h->p.CageStrain = cagestrain;
    for (int r = 0; r < h->p.nreplicas; r++) {
        h->h_cage[r] = cagestrain;
        int rc = sn_push_couplings(h, r);
        if (rc) return rc;
    }
    return SN_OK;
}

extern "C" int sn_set_replica_cagestrain(sn_handle *h, int replica, double cagestrain)
{
    SN_CHECK_HANDLE(h, replica);
    h->h_cage[replica] = cagestrain;
    return sn_push_couplings(h, replica);
}

extern "C" int sn_synchronize(sn_handle *h)
{
    SN_CHECK_HANDLE(h, 0);
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return sn_check_device_error(h);
}

// ---- sweeps -----------------------------------------------------------------
static SnSweepArgs sn_sweep_args(sn_handle *h)
{
    SnSweepArgs a;
    a.lat = h->lat; a.G = h->G;
    a.ax = sn_axis_colour(h->G.X, h->p.cutoff, false);
    a.ay = sn_axis_colour(h->G.Y, h->p.cutoff, false);
    a.az = sn_axis_colour(h->G.nz, h->p.cutoff, h->G.Z == 1);
    a.beta = h->beta; a.efield = h->efield;
    a.K = (float)h->p.K;
    a.constrain = h->p.ConstrainToX; a.dim = h->p.DIM;
    a.counters = h->counters;
    a.rep_key = h->rep_key;
    a.sweep_lo = (uint32_t)h->sweep; a.sweep_hi = (uint32_t)(h->sweep >> 32);
    a.nb = h->nb_table; a.nnb = h->nnb;
    a.peer_lo = h->peer_lat[0]; a.peer_hi = h->peer_lat[1];
    a.audit = h->audit_dev; a.audit_group = 0;
    return a;
}

template <int MODE, bool SPECIES>
static void sn_launch_colour(const SnSweepArgs &a, int nrep, int cx, int cy, int cz, cudaStream_t st)
{
    const long long total = (long long)sn_axis_count(a.ax, cx) * sn_axis_count(a.ay, cy) * sn_axis_count(a.az, cz);
    dim3 grid((unsigned)((total + 127) / 128), nrep);
    sn_colour_pass_kernel<MODE, SPECIES><<<grid, 128, 0, st>>>(a, cx, cy, cz);
}

int sn_sweep_colour_launch(sn_handle *h, long long nsweeps, long long *launches)
{
    const int mode = sn_mode(h);
    { int rc = sn_sync_canonical(h); if (rc) return rc; }
    if (nsweeps > 0) h->lat2_valid = false;
    for (long long s = 0; s < nsweeps; s++) {
        SnSweepArgs a = sn_sweep_args(h);
        for (int cx = 0; cx < a.ax.ncol; cx++) for (int cy = 0; cy < a.ay.ncol; cy++) for (int cz = 0; cz < a.az.ncol; cz++) {
            a.audit_group = (cx * a.ay.ncol + cy) * a.az.ncol + cz;
            if (mode == 0) { if (h->species) sn_launch_colour<0, true>(a, h->p.nreplicas, cx, cy, cz, h->stream); else sn_launch_colour<0, false>(a, h->p.nreplicas, cx, cy, cz, h->stream); }
            else if (mode == 1) { if (h->species) sn_launch_colour<1, true>(a, h->p.nreplicas, cx, cy, cz, h->stream); else sn_launch_colour<1, false>(a, h->p.nreplicas, cx, cy, cz, h->stream); }
            else sn_launch_colour<2, true>(a, h->p.nreplicas, cx, cy, cz, h->stream);
            if (launches) (*launches)++;
            if (!h->G.periodic_z) { int rc = sn_slab_phase_sync(h, launches); if (rc) return rc; }
        }
        h->sweep++;
    }
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

static int sn_sweeps_impl(sn_handle *h, long long nsweeps, long long *launches)
{
    if (nsweeps < 0) return sn_fail(SN_ERR_INVALID, "sn_mc_sweeps: nsweeps %lld", nsweeps);
    if (!h->G.periodic_z && (!h->peer_lat[0] || !h->peer_lat[1]))
        return sn_fail(SN_ERR_INVALID, "sn_mc_sweeps: Z-slab handle has no neighbours attached (sn_ipc_attach / sn_attach_peer)");
    { int rc = sn_resolve_species(h); if (rc) return rc; }
    int rc;
    if (h->use_resident) rc = sn_sweep_resident_launch(h, nsweeps, launches);
    else rc = h->use_tiled ? sn_sweep_tiled_launch(h, nsweeps, launches) : sn_sweep_colour_launch(h, nsweeps, launches);
    if (rc) return rc;
    SN_CUDA_CHECK(cudaEventRecord(h->ev_sweeps, h->stream));     // sn_order_after
    return SN_OK;
}

extern "C" int sn_mc_sweeps(sn_handle *h, long long nsweeps)
{
    SN_CHECK_HANDLE(h, 0);
    return sn_sweeps_impl(h, nsweeps, nullptr);
}

// Make h's later work wait until the sweeps queued so far on `other` are done (device-side, the host does not block).
// Two handles used as a double buffer -- one sweeping while the other's lattice is in flight over PCIe -- keep their
// persistent sweep kernels in a fixed order this way (with Z-slabs on several GPUs every GPU must run them in the
// same order, or the slabs would wait for each other across the two jobs).
extern "C" int sn_order_after(sn_handle *h, sn_handle *other)
{
    SN_CHECK_HANDLE(h, 0);
    if (!other) return sn_fail(SN_ERR_INVALID, "sn_order_after: null");
    SN_CUDA_CHECK(cudaStreamWaitEvent(h->stream, other->ev_sweeps, 0));
    if (h->serial_group != other->serial_group) {
        // from now on the two are one job as far as SM sharing goes: their persistent kernels never run at once,
        // so each may use every SM its group is entitled to (the caller keeps ordering them, on every GPU alike)
        std::lock_guard<std::mutex> lock(sn_slab_registry_mutex);
        const void *from = h->serial_group;
        for (sn_handle *a : sn_slab_registry) if (a->serial_group == from) a->serial_group = other->serial_group;
        h->serial_group = other->serial_group;
        sn_slab_shares();
    }
    return SN_OK;
}

extern "C" int sn_mc_sweeps_timed(sn_handle *h, long long nsweeps, double *ms, long long *launches)
{
    SN_CHECK_HANDLE(h, 0);
    if (!h->ev[0]) { SN_CUDA_CHECK(cudaEventCreate(&h->ev[0])); }     // owned by the handle, released in sn_destroy
    if (!h->ev[1]) { SN_CUDA_CHECK(cudaEventCreate(&h->ev[1])); }
    long long n = 0;
    SN_CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
    int rc = sn_sweeps_impl(h, nsweeps, &n);
    SN_CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));
    SN_CUDA_CHECK(cudaEventSynchronize(h->ev[1]));
    float t = 0.f;
    SN_CUDA_CHECK(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
    if (ms) *ms = t;
    if (launches) *launches = n;
    return rc;
}

// One sweep with every attempt recorded (see the header).  Tiled handles run the AUDIT instantiation of the
// tiled kernel -- same arithmetic, same order; every other handle runs the colour passes (the resident kernel's
// chain is bit-identical to them).
extern "C" int sn_mc_sweep_audit(sn_handle *h, float *records)
{
    SN_CHECK_HANDLE(h, 0);
    if (!records) return sn_fail(SN_ERR_INVALID, "sn_mc_sweep_audit: null buffer");
    if (!h->G.periodic_z && (!h->peer_lat[0] || !h->peer_lat[1]))
        return sn_fail(SN_ERR_INVALID, "sn_mc_sweep_audit: Z-slab handle has no neighbours attached");
    const size_t bytes = sizeof(float) * SN_AUDIT_WORDS * (size_t)h->p.nreplicas * h->G.X * h->G.Y * h->G.nz;
    float *dev = nullptr;
    if (cudaMalloc(&dev, bytes) != cudaSuccess) { cudaGetLastError(); return sn_fail(SN_ERR_NOMEM, "sn_mc_sweep_audit: cannot allocate %.1f MB", bytes / 1e6); }
    int rc = sn_resolve_species(h);
    if (rc) { cudaFree(dev); return rc; }
    cudaError_t e = cudaMemsetAsync(dev, 0, bytes, h->stream);
    if (e == cudaSuccess) {
        h->audit_dev = dev;
        rc = h->use_tiled ? sn_sweep_tiled_launch(h, 1, nullptr) : sn_sweep_colour_launch(h, 1, nullptr);
        h->audit_dev = nullptr;
        if (!rc) e = cudaMemcpyAsync(records, dev, bytes, cudaMemcpyDeviceToHost, h->stream);
    }
    if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    else cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    if (rc) return rc;
    if (e != cudaSuccess) return sn_fail(SN_ERR_CUDA, "sn_mc_sweep_audit: %s", cudaGetErrorString(e));
    return SN_OK;
}

extern "C" int sn_get_counters(sn_handle *h, int replica, unsigned long long *accept, unsigned long long *reject,
                               unsigned long long *vacant)
{
    SN_CHECK_HANDLE(h, replica);
    unsigned long long c[3];
    SN_CUDA_CHECK(cudaMemcpyAsync(c, h->counters + 3 * replica, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    { int rc = sn_check_device_error(h); if (rc) return rc; }
    if (accept) *accept = c[0];
    if (reject) *reject = c[1];
    if (vacant) *vacant = c[2];
    return SN_OK;
}

extern "C" int sn_reset_counters(sn_handle *h)
{
    SN_CHECK_HANDLE(h, 0);
    SN_CUDA_CHECK(cudaMemsetAsync(h->counters, 0, sizeof(unsigned long long) * 3 * h->p.nreplicas, h->stream));
    return SN_OK;
}

// ---- checkpoint / restart -----------------------------------------------------------
// The chain's state is the lattice (sn_get_lattice), the number of sweeps done -- the Philox counter word,
// and the version of every tile in the dataflow kernel -- and the ACCEPT / REJECT counters.
__global__ void sn_fill_u32_kernel(unsigned int *p, long long n, unsigned int v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

extern "C" int sn_set_replica_seed(sn_handle *h, int replica, unsigned long long seed)
{
    SN_CHECK_HANDLE(h, replica);
    const uint4 k = make_uint4((uint32_t)seed, (uint32_t)(seed >> 32), 0u, 0u);     // = replica 0 of a handle created with this seed
    SN_CUDA_CHECK(cudaMemcpyAsync(h->rep_key + replica, &k, sizeof k, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SN_OK;
}

extern "C" int sn_get_sweep_count(sn_handle *h, unsigned long long *sweeps)
{
    SN_CHECK_HANDLE(h, 0);
    if (!sweeps) return sn_fail(SN_ERR_INVALID, "sn_get_sweep_count: null");
    *sweeps = h->sweep;
    return SN_OK;
}

extern "C" int sn_set_sweep_count(sn_handle *h, unsigned long long sweeps)
{
    SN_CHECK_HANDLE(h, 0);
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->sweep = sweeps;
    if (h->nver > 0) {                               // every tile (ghost layers included) has completed `sweeps` sweeps
        sn_fill_u32_kernel<<<64, 256, 0, h->stream>>>(h->flags + SN_FLAGS_VER, (long long)h->nver, (unsigned int)sweeps);
        SN_CUDA_CHECK(cudaGetLastError());
        SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    return SN_OK;
}

extern "C" int sn_set_counters(sn_handle *h, int replica, unsigned long long accept, unsigned long long reject, unsigned long long vacant)
{
    SN_CHECK_HANDLE(h, replica);
    const unsigned long long c[3] = {accept, reject, vacant};
    SN_CUDA_CHECK(cudaMemcpyAsync(h->counters + 3 * replica, c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SN_OK;
}

// ---- energy audit -------------------------------------------------------------
static SnTerms sn_terms(const sn_handle *h, int replica)
{
    SnTerms t;
    t.cage = (float)h->h_cage[replica]; t.K = (float)h->p.K; t.beta = h->h_beta[replica];
    t.E = make_float3(h->h_efield[3 * replica], h->h_efield[3 * replica + 1], h->h_efield[3 * replica + 2]);
    t.constrain = h->p.ConstrainToX; t.dim = h->p.DIM;
    return t;
}

extern "C" int sn_site_energy(sn_handle *h, int replica, int precision, int n, const int *sites, const float *newdip, double *dE)
{
    SN_CHECK_HANDLE(h, replica);
    if (n < 0 || (n > 0 && (!sites || !newdip || !dE))) return sn_fail(SN_ERR_INVALID, "sn_site_energy: bad arguments");
    if (precision < SN_PREC_F32 || precision > SN_PREC_REPLICA) return sn_fail(SN_ERR_INVALID, "sn_site_energy: precision %d", precision);
    if (n == 0) return SN_OK;
    { int rc0 = sn_sync_canonical(h); if (rc0) return rc0; }
    for (int i = 0; i < n; i++)
        if (sites[3 * i] < 0 || sites[3 * i] >= h->G.X || sites[3 * i + 1] < 0 || sites[3 * i + 1] >= h->G.Y || sites[3 * i + 2] < 0 || sites[3 * i + 2] >= h->G.nz)
            return sn_fail(SN_ERR_INVALID, "sn_site_energy: site %d (%d,%d,%d) outside the lattice", i, sites[3 * i], sites[3 * i + 1], sites[3 * i + 2]);
    const size_t bs = (size_t)n * 3 * sizeof(int), bn = (size_t)n * 3 * sizeof(float), bo = (size_t)n * sizeof(double);
    void *s; int rc = sn_scratch(h, bs + bn + bo + 64, &s);
    if (rc) return rc;
    double *d_out = (double *)s;
    int *d_sites = (int *)((char *)s + bo);
    float *d_new = (float *)((char *)s + bo + bs);
    SN_CUDA_CHECK(cudaMemcpyAsync(d_sites, sites, bs, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaMemcpyAsync(d_new, newdip, bn, cudaMemcpyHostToDevice, h->stream));
    if (precision == SN_PREC_F32) {
        const float4 *lat = h->lat + (long long)replica * h->G.rep_stride;
        const SnTerms t = sn_terms(h, replica);
        const int gs = (n + 127) / 128, mode = sn_mode(h);
        if (mode == 0) sn_site_energy_f32_kernel<0, true><<<gs, 128, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t, n, d_sites, d_new, d_out);
        else if (mode == 1) sn_site_energy_f32_kernel<1, true><<<gs, 128, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t, n, d_sites, d_new, d_out);
        else sn_site_energy_f32_kernel<2, true><<<gs, 128, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t, n, d_sites, d_new, d_out);
        SN_CUDA_CHECK(cudaGetLastError());
    } else if ((rc = sn_energy_exact_launch(h, replica, precision, n, d_sites, d_new, d_out))) return rc;
    SN_CUDA_CHECK(cudaMemcpyAsync(dE, d_out, bo, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SN_OK;
}

static int sn_reduce_to_host(sn_handle *h, const double *d_partials, int nblocks, int nv, double *out)
{
    std::vector<double> hp((size_t)nblocks * nv);
    SN_CUDA_CHECK(cudaMemcpyAsync(hp.data(), d_partials, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < nv; k++) { double s = 0.0; for (int b = 0; b < nblocks; b++) s += hp[(size_t)b * nv + k]; out[k] = s; }
    return SN_OK;
}

extern "C" int sn_total_energy(sn_handle *h, int replica, int precision, double out[4])
{
    SN_CHECK_HANDLE(h, replica);
    if (!out) return sn_fail(SN_ERR_INVALID, "sn_total_energy: null");
    if (precision < SN_PREC_F32 || precision > SN_PREC_REPLICA) return sn_fail(SN_ERR_INVALID, "sn_total_energy: precision %d", precision);
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int nblocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
    int rc;
    if ((rc = sn_sync_canonical(h))) return rc;
    if (precision == SN_PREC_F32) {
        void *s; if ((rc = sn_scratch(h, sizeof(double) * 4 * nblocks, &s))) return rc;
        const float4 *lat = h->lat + (long long)replica * h->G.rep_stride;
        const SnTerms t = sn_terms(h, replica);
        const int mode = sn_mode(h);
        if (mode == 0) sn_energy_f32_kernel<0, true><<<nblocks, 256, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t.cage, t.K, t.E, (double *)s);
        else if (mode == 1) sn_energy_f32_kernel<1, true><<<nblocks, 256, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t.cage, t.K, t.E, (double *)s);
        else sn_energy_f32_kernel<2, true><<<nblocks, 256, 0, h->stream>>>(lat, h->G, h->nb_table, h->nnb, t.cage, t.K, t.E, (double *)s);
        SN_CUDA_CHECK(cudaGetLastError());
        if ((rc = sn_reduce_to_host(h, (double *)s, nblocks, 4, out))) return rc;
    } else {
        void *s; if ((rc = sn_scratch(h, sizeof(double) * (n + nblocks), &s))) return rc;
        double *map = (double *)s, *part = map + n;
        for (int k = 0; k < 4; k++) {
            if ((rc = sn_energy_exact_map_launch(h, replica, precision, 1 << k, map))) return rc;
            sn_sum_doubles_kernel<<<nblocks, 256, 0, h->stream>>>(map, n, part);
            SN_CUDA_CHECK(cudaGetLastError());
            if ((rc = sn_reduce_to_host(h, part, nblocks, 1, out + k))) return rc;
        }
    }
    out[0] *= 0.5; out[1] *= 0.5;       // pair terms are counted from both ends
    return SN_OK;
}

// ---- observables --------------------------------------------------------------
static int sn_dipole_sum(sn_handle *h, int replica, double S[3])
{
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int nblocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
    void *s; int rc = sn_scratch(h, sizeof(double) * 3 * nblocks, &s);
    if (rc || (rc = sn_sync_canonical(h))) return rc;
    sn_sum_dipoles_kernel<<<nblocks, 256, 0, h->stream>>>(h->lat + (long long)replica * h->G.rep_stride, h->G, (double *)s);
    SN_CUDA_CHECK(cudaGetLastError());
    return sn_reduce_to_host(h, (double *)s, nblocks, 3, S);
}

extern "C" int sn_polarisation(sn_handle *h, int replica, double P[3])
{
    SN_CHECK_HANDLE(h, replica);
    if (!P) return sn_fail(SN_ERR_INVALID, "sn_polarisation: null");
    int rc = sn_dipole_sum(h, replica, P);
    if (rc) return rc;
    const double n = (double)h->G.X * h->G.Y * h->G.nz;    // analysis.c:60
    for (int k = 0; k < 3; k++) P[k] /= n;
    return SN_OK;
}

extern "C" int sn_state_hash(sn_handle *h, int replica, unsigned long long *hash)
{
    SN_CHECK_HANDLE(h, replica);
    if (!hash) return sn_fail(SN_ERR_INVALID, "sn_state_hash: null");
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int nblocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
    void *s; int rc = sn_scratch(h, 64, &s);
    if (rc || (rc = sn_sync_canonical(h))) return rc;
    SN_CUDA_CHECK(cudaMemsetAsync(s, 0, sizeof(unsigned long long), h->stream));
    sn_state_hash_kernel<<<nblocks, 256, 0, h->stream>>>(h->lat + (long long)replica * h->G.rep_stride, h->G, (unsigned long long *)s);
    SN_CUDA_CHECK(cudaGetLastError());
    SN_CUDA_CHECK(cudaMemcpyAsync(hash, s, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return sn_check_device_error(h);
}

extern "C" int sn_kernel_in_use(sn_handle *h, int *kernel)
{
    SN_CHECK_HANDLE(h, 0);
    if (!kernel) return sn_fail(SN_ERR_INVALID, "sn_kernel_in_use: null");
    *kernel = h->use_resident ? SN_KERNEL_RESIDENT : h->use_tiled ? (h->p.kernel == SN_KERNEL_TILED_PHASED ? SN_KERNEL_TILED_PHASED : SN_KERNEL_TILED)
                                                                  : SN_KERNEL_COLOUR;
    return SN_OK;
}

extern "C" int sn_landau_order(sn_handle *h, int replica, double *landau)
{
    SN_CHECK_HANDLE(h, replica);
    if (!landau) return sn_fail(SN_ERR_INVALID, "sn_landau_order: null");
    double S[3];
    int rc = sn_dipole_sum(h, replica, S);
    if (rc) return rc;
    const double n = (double)h->G.X * h->G.Y * h->G.nz;
    *landau = (S[0] * S[0] + S[1] * S[1] + S[2] * S[2]) / n * n;   // analysis.c:523, as written
    return SN_OK;
}

// Which array holds a replica's sites right now, and where the planes beyond a Z-slab's own live.  Slab neighbours
// are reached through the array their sweep kernel works on (the pointer sn_ipc_attach / sn_attach_peer stored), so a
// slab handle is read in that layout too; every slab must be quiescent (same call sequence on all slabs: the next
// sweep of a neighbour waits for this handle's handshake, which is queued behind the observable kernels).
static int sn_make_view(sn_handle *h, int replica, int reach, const char *who, SnLatView *v)
{
    const SnGeom &G = h->G;
    int rc;
    if (!G.periodic_z) {
        if (!h->peer_lat[0] || !h->peer_lat[1]) return sn_fail(SN_ERR_INVALID, "%s: Z-slab handle has no neighbours attached (sn_ipc_attach / sn_attach_peer)", who);
        if (G.gz > 0 && G.nz < reach) return sn_fail(SN_ERR_UNSUPPORTED, "%s: slab height %d is smaller than the stencil radius %d", who, G.nz, reach);
    }
    if (h->use_tiled && (h->lat2_valid || !G.periodic_z)) {
        if (!h->lat2_valid) { if ((rc = sn_sync_canonical(h)) || (rc = sn_convert_layout(h, true))) return rc; h->lat2_valid = true; }
        const long long rs = sn_rep_stride2(G) * replica;
        v->own = h->lat2 + rs; v->tiled = 1;
        v->lo = G.periodic_z ? v->own : h->peer_lat[0] + rs;
        v->hi = G.periodic_z ? v->own : h->peer_lat[1] + rs;
    } else {
        if ((rc = sn_sync_canonical(h))) return rc;
        const long long rs = G.rep_stride * replica;
        v->own = h->lat + rs; v->tiled = 0;
        v->lo = G.periodic_z ? v->own : h->peer_lat[0] + rs;
        v->hi = G.periodic_z ? v->own : h->peer_lat[1] + rs;
    }
    return SN_OK;
}

static int sn_obs_blocks(const sn_handle *h) { return sno::tiles(h->G.X) * sno::tiles(h->G.Y) * sno::tiles(h->G.nz); }

extern "C" int sn_rdf(sn_handle *h, int replica, double *fe_sum, double *afe_sum, long long *count)
{
    SN_CHECK_HANDLE(h, replica);
    if (!fe_sum || !afe_sum || !count) return sn_fail(SN_ERR_INVALID, "sn_rdf: null");
    const int CUT = sno::RDF_R;                     // analysis.c:540
    SnLatView view; int rc = sn_make_view(h, replica, CUT, "sn_rdf", &view);
    if (rc) return rc;
    static_assert(SN_RDF_BINS == sno::RDF_NBINS, "bin count");
    std::vector<SnRdfOffset> off;
    std::vector<long long> mult(SN_RDF_BINS, 0);    // lattice vectors per r^2 (both signs): the reference's count per site
    for (int dx = -CUT; dx <= CUT; dx++) for (int dy = -CUT; dy <= CUT; dy++) for (int dz = -CUT; dz <= CUT; dz++) {   // dz in [-9, 9] whatever Z is (% wraps it)
        const int r2 = dx * dx + dy * dy + dz * dz;
        if (r2 >= SN_RDF_BINS) continue;            // r^2 == 81 is neither zeroed nor printed by the reference
        mult[r2]++;
        // one of {d, -d}: the kernel walks the upper half space and the origin, the sums of r^2 > 0 are doubled below
        if (!(dx > 0 || (dx == 0 && dy > 0) || (dx == 0 && dy == 0 && dz >= 0))) continue;
        SnRdfOffset o;
        o.delta = (dx * sno::RDF_NY + dy) * sno::RDF_NZ + dz; o.r2 = r2;
        o.dx = dx; o.dy = dy; o.dz = dz;
        off.push_back(o);
    }
    if ((int)off.size() > sno::RDF_NOFF) return sn_fail(SN_ERR_CUDA, "sn_rdf: internal: %d offsets in the half space, table holds %d", (int)off.size(), sno::RDF_NOFF);
    std::stable_sort(off.begin(), off.end(), [](const SnRdfOffset &a, const SnRdfOffset &b) { return a.r2 < b.r2; });
    std::vector<int> first(SN_RDF_BINS + 1, 0);
    for (auto &o : off) first[o.r2 + 1]++;
    for (int b = 0; b < SN_RDF_BINS; b++) first[b + 1] += first[b];
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int nblocks = sn_obs_blocks(h), nv = 2 * SN_RDF_BINS;
    const size_t b_part = sizeof(double) * nv * (size_t)nblocks, b_tot = sizeof(double) * nv;
    void *s; if ((rc = sn_scratch(h, b_part + b_tot + 256, &s))) return rc;
    double *d_part = (double *)s, *d_tot = d_part + (size_t)nv * nblocks;
    // the table is identical for every handle of the device; re-sending the same bytes under a running kernel is harmless
    SN_CUDA_CHECK(cudaMemcpyToSymbolAsync(sn_c_rdf_off, off.data(), off.size() * sizeof(SnRdfOffset), 0, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaMemcpyToSymbolAsync(sn_c_rdf_first, first.data(), first.size() * sizeof(int), 0, cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaFuncSetAttribute(sn_rdf_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sno::RDF_SMEM));
    sn_rdf_tiled_kernel<<<nblocks, sno::THREADS, sno::RDF_SMEM, h->stream>>>(view, h->G, d_part);
    sn_reduce_rows_kernel<<<nv, 256, 0, h->stream>>>(d_part, nblocks, nv, d_tot);
    SN_CUDA_CHECK(cudaGetLastError());
    std::vector<double> tot(nv);
    SN_CUDA_CHECK(cudaMemcpyAsync(tot.data(), d_tot, b_tot, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < SN_RDF_BINS; b++) {
        const double twice = b == 0 ? 1.0 : 2.0;
        fe_sum[b] = twice * tot[2 * b]; afe_sum[b] = twice * tot[2 * b + 1];
        count[b] = mult[b] * n;                                     // analysis.c:578, one count per (site, offset)
    }
    return sn_check_device_error(h);
}

// potential map of the handle's own sites into device scratch (first n doubles of *scratch); extra_bytes are reserved behind it
static int sn_potential_device(sn_handle *h, int replica, size_t extra_bytes, double **d_v, void **extra)
{
    const int MAXR = sno::POT_R;                    // analysis.c:68
    SnLatView view; int rc = sn_make_view(h, replica, MAXR, "sn_potential_map", &view);
    if (rc) return rc;
    std::vector<SnPotOffset> off;
    int pitch, step; sn_obs_layout(sno::POT_N, &pitch, &step);
    for (int dx = -MAXR; dx <= MAXR; dx++) for (int dy = -MAXR; dy <= MAXR; dy++) for (int dz = -MAXR; dz <= MAXR; dz++) {
        if (!dx && !dy && !dz) continue;
        const double d = sqrt((double)(dx * dx + dy * dy + dz * dz));
        if (d > (double)MAXR) continue;
        const double w = 1.0 / (d * d * d);
        SnPotOffset o; o.delta = (dx * sno::POT_N + dy) * pitch + dz; o.pad = 0; o.kx = dx * w; o.ky = dy * w; o.kz = dz * w;
        off.push_back(o);
    }
    if ((int)off.size() > sno::POT_MAXOFF) return sn_fail(SN_ERR_CUDA, "sn_potential_map: internal: %d offsets", (int)off.size());
    const int smem = sno::POT_N * sno::POT_N * pitch * 24 + (int)(off.size() * sizeof(SnPotOffset));
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const size_t b_v = ((sizeof(double) * n + 15) / 16) * 16, b_off = ((off.size() * sizeof(SnPotOffset) + 15) / 16) * 16;
    void *s; if ((rc = sn_scratch(h, b_v + b_off + extra_bytes + 64, &s))) return rc;
    *d_v = (double *)s;
    SnPotOffset *d_off = (SnPotOffset *)((char *)s + b_v);
    if (extra) *extra = (char *)s + b_v + b_off;
    SN_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(SnPotOffset), cudaMemcpyHostToDevice, h->stream));
    SN_CUDA_CHECK(cudaFuncSetAttribute(sn_potential_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    sn_potential_tiled_kernel<<<sn_obs_blocks(h), sno::THREADS, smem, h->stream>>>(view, h->G, d_off, (int)off.size(), pitch, step, *d_v);
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

extern "C" int sn_potential_map(sn_handle *h, int replica, double *V)
{
    SN_CHECK_HANDLE(h, replica);
    if (!V) return sn_fail(SN_ERR_INVALID, "sn_potential_map: null");
    double *d_v; int rc = sn_potential_device(h, replica, 0, &d_v, nullptr);
    if (rc) return rc;
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    SN_CUDA_CHECK(cudaMemcpyAsync(V, d_v, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return sn_check_device_error(h);
}

extern "C" int sn_efield_map(sn_handle *h, int replica, int cutoff, int half_offset, double *Emag)
{
    SN_CHECK_HANDLE(h, replica);
    if (!Emag) return sn_fail(SN_ERR_INVALID, "sn_efield_map: null");
    if (cutoff < 1 || cutoff > 8) return sn_fail(SN_ERR_INVALID, "sn_efield_map: cutoff %d outside 1..8", cutoff);
    // offsets as the reference walks them: integer steps without the origin (analysis.c:407-418), or
    // dx + 0.5 for dx in [-cutoff-1, cutoff-1] (analysis.c:322-334); d <= cutoff
    const int lo = half_offset ? -cutoff - 1 : -cutoff, hi = half_offset ? cutoff - 1 : cutoff;
    const int reach = cutoff + (half_offset ? 1 : 0);
    const bool tiled = reach <= 6;                  // the box of an 8^3 tile with a halo of 6 fits in shared memory
    if (!tiled && !h->G.periodic_z) return sn_fail(SN_ERR_UNSUPPORTED, "sn_efield_map: cutoff %d on a Z-slab handle (radius above 6)", cutoff);
    const int N = sno::T + 2 * reach;
    int pitch, step; sn_obs_layout(N, &pitch, &step);
    std::vector<SnEfOffset> off; std::vector<SnEfOffset2> off2;
    for (int dx = lo; dx <= hi; dx++) for (int dy = lo; dy <= hi; dy++) for (int dz = lo; dz <= hi; dz++) {
        if (!half_offset && !dx && !dy && !dz) continue;
        const double sh = half_offset ? 0.5 : 0.0, rx = dx + sh, ry = dy + sh, rz = dz + sh;
        const double d = sqrt(rx * rx + ry * ry + rz * rz);
        if (d > (double)cutoff) continue;
        SnEfOffset o; o.dx = (short)dx; o.dy = (short)dy; o.dz = (short)dz; o.pad = 0;
        o.nx = rx / d; o.ny = ry / d; o.nz = rz / d; o.w = 1.0 / (d * d * d);
        off.push_back(o);
        SnEfOffset2 q; q.delta = (dx * N + dy) * pitch + dz; q.pad[0] = q.pad[1] = q.pad[2] = 0; q.nx = o.nx; q.ny = o.ny; q.nz = o.nz; q.w = o.w;
        off2.push_back(q);
    }
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const size_t b_v = ((sizeof(double) * n + 15) / 16) * 16, b_off = off.size() * std::max(sizeof(SnEfOffset), sizeof(SnEfOffset2));
    void *s; int rc = sn_scratch(h, b_v + b_off + 64, &s);
    if (rc) return rc;
    double *d_v = (double *)s;
    void *d_off = (char *)s + b_v;
    if (tiled) {
        SnLatView view;
        if ((rc = sn_make_view(h, replica, reach, "sn_efield_map", &view))) return rc;
        const int smem = N * N * pitch * 24 + (int)(off2.size() * sizeof(SnEfOffset2));
        if (smem > 227 * 1024) return sn_fail(SN_ERR_UNSUPPORTED, "sn_efield_map: cutoff %d needs %d bytes of shared memory", cutoff, smem);
        SN_CUDA_CHECK(cudaMemcpyAsync(d_off, off2.data(), off2.size() * sizeof(SnEfOffset2), cudaMemcpyHostToDevice, h->stream));
        SN_CUDA_CHECK(cudaFuncSetAttribute(sn_efield_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        sn_efield_tiled_kernel<<<sn_obs_blocks(h), sno::THREADS, smem, h->stream>>>(view, h->G, (const SnEfOffset2 *)d_off, (int)off2.size(), reach, pitch, step,
                                                                                      half_offset ? 0 : 1, d_v);
    } else {
        if ((rc = sn_sync_canonical(h))) return rc;
        SN_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(SnEfOffset), cudaMemcpyHostToDevice, h->stream));
        const float4 *lat = h->lat + (long long)replica * h->G.rep_stride;
        if (h->G.X >= reach && h->G.Y >= reach && h->G.nz >= reach)
            sn_efield_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(lat, h->G, (const SnEfOffset *)d_off, (int)off.size(), half_offset ? 0 : 1, d_v);
        else
            sn_efield_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(lat, h->G, (const SnEfOffset *)d_off, (int)off.size(), half_offset ? 0 : 1, d_v);
    }
    SN_CUDA_CHECK(cudaGetLastError());
    SN_CUDA_CHECK(cudaMemcpyAsync(Emag, d_v, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return sn_check_device_error(h);
}

// recombination_calculator() (analysis.c:96-170) in two halves, so that the slabs of a decomposed lattice can be merged:
// partial sums over the handle's own sites, then the normalisations over the whole lattice.
extern "C" int sn_recombination_partial(sn_handle *h, int replica, double part[SN_RECOMB_PARTIAL_N])
{
    SN_CHECK_HANDLE(h, replica);
    if (!part) return sn_fail(SN_ERR_INVALID, "sn_recombination_partial: null");
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int nblocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
    double *d_v; void *extra;
    int rc = sn_potential_device(h, replica, sizeof(double) * 8 * nblocks, &d_v, &extra);
    if (rc) return rc;
    const double BETA = 1 / (0.025), potentialeV = 0.165 / 5;                 // analysis.c:104-106
    sn_recombination_kernel<<<nblocks, 256, 0, h->stream>>>(d_v, n, h->G.nz, h->G.z0 == 0 ? 1 : 0, potentialeV * BETA, (double *)extra);
    SN_CUDA_CHECK(cudaGetLastError());
    std::vector<double> hp((size_t)nblocks * 8);
    SN_CUDA_CHECK(cudaMemcpyAsync(hp.data(), extra, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < SN_RECOMB_PARTIAL_N; k++) part[k] = 0.0;
    for (int b = 0; b < nblocks; b++) {
        for (int k = 0; k < 5; k++) part[k] += hp[(size_t)b * 8 + k];
        for (int k = 5; k < 8; k++) part[k] = std::max(part[k], hp[(size_t)b * 8 + k]);
    }
    part[8] = (double)n;
    return sn_check_device_error(h);
}

extern "C" int sn_recombination_finish(int nparts, const double *parts, double out[SN_RECOMB_N])
{
    if (nparts < 1 || !parts || !out) return sn_fail(SN_ERR_INVALID, "sn_recombination_finish: bad arguments");
    double sum[5] = {0, 0, 0, 0, 0}, mx[3] = {0, 0, 0}, N = 0.0;
    for (int p = 0; p < nparts; p++) {
        const double *q = parts + (size_t)p * SN_RECOMB_PARTIAL_N;
        for (int k = 0; k < 5; k++) sum[k] += q[k];
        for (int k = 0; k < 3; k++) mx[k] = std::max(mx[k], q[5 + k]);
        N += q[8];
    }
    const double ZBe = sum[0], ZBh = sum[1], ZFDe = sum[2], ZFDh = sum[3];
    out[0] = ZBe; out[1] = ZBh; out[2] = ZFDe; out[3] = ZFDh;
    out[4] = N * N / (ZBe * ZBh);                                             // R_Boltz (:131; in double, the reference's int product wraps)
    out[5] = N * sum[4] / (ZFDe * ZFDh);                                      // R_FD = N * sum e_i h_i (:169)
    out[6] = sum[2] / ZFDe; out[7] = sum[3] / ZFDh;                            // FD totals (:163-164)
    out[8] = mx[0] / ZFDe; out[9] = mx[1] / ZFDh; out[10] = mx[2] / (ZFDe * ZFDh);   // maxima over z = 0 (:157-159)
    return SN_OK;
}

extern "C" int sn_recombination(sn_handle *h, int replica, double out[SN_RECOMB_N])
{
    SN_CHECK_HANDLE(h, replica);
    if (!out) return sn_fail(SN_ERR_INVALID, "sn_recombination: null");
    if (!h->G.periodic_z) return sn_fail(SN_ERR_UNSUPPORTED, "sn_recombination: a Z-slab handle holds part of the lattice: merge the slabs' "
                                                             "sn_recombination_partial results with sn_recombination_finish");
    double part[SN_RECOMB_PARTIAL_N];
    int rc = sn_recombination_partial(h, replica, part);
    if (rc) return rc;
    return sn_recombination_finish(1, part, out);
}

// ---- Philox known-answer check ---------------------------------------------------
__global__ void sn_philox_kat_kernel(int n, const uint32_t *__restrict__ in, uint32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox4 r = sn_philox4x32_10(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// sn_philox4x32_10 -- the generator every sweep kernel draws from -- evaluated for n (counter[4], key[2]) inputs
// by the host compilation of the function (out_host) and by the device (out_device; pass NULL to skip: no GPU needed).
extern "C" int sn_philox_kat(int n, const unsigned int *counter_key, unsigned int *out_host, unsigned int *out_device)
{
    if (n < 0 || (n > 0 && !counter_key)) return sn_fail(SN_ERR_INVALID, "sn_philox_kat: bad arguments");
    if (out_host)
        for (int i = 0; i < n; i++) {
            const unsigned int *c = counter_key + 6 * i;
            const Philox4 r = sn_philox4x32_10(c[0], c[1], c[2], c[3], c[4], c[5]);
            out_host[4 * i] = r.x; out_host[4 * i + 1] = r.y; out_host[4 * i + 2] = r.z; out_host[4 * i + 3] = r.w;
        }
    if (out_device && n > 0) {
        uint32_t *d_in = nullptr, *d_out = nullptr;
        cudaError_t e = cudaMalloc(&d_in, sizeof(uint32_t) * 6 * n);
        if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(uint32_t) * 4 * n);
        if (e == cudaSuccess) e = cudaMemcpy(d_in, counter_key, sizeof(uint32_t) * 6 * n, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { sn_philox_kat_kernel<<<(n + 127) / 128, 128>>>(n, d_in, d_out); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpy(out_device, d_out, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost);
        cudaFree(d_in); cudaFree(d_out);
        if (e != cudaSuccess) return sn_fail(SN_ERR_CUDA, "sn_philox_kat: %s", cudaGetErrorString(e));
    }
    return SN_OK;
}

// ---- FP32 roofline denominator ---------------------------------------------------
__global__ void __launch_bounds__(256) sn_ffma_peak_kernel(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int sn_bench_fp32_peak(int device, double *tflops)
{
    if (!tflops) return sn_fail(SN_ERR_INVALID, "sn_bench_fp32_peak: null");
    SN_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float *out;
    SN_CUDA_CHECK(cudaMalloc(&out, sizeof(float) * blocks * 256));
    cudaEvent_t e0, e1;
    SN_CUDA_CHECK(cudaEventCreate(&e0));
    SN_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        SN_CUDA_CHECK(cudaEventRecord(e0));
        sn_ffma_peak_kernel<<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
        SN_CUDA_CHECK(cudaEventRecord(e1));
        SN_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        SN_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * 256;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return SN_OK;
}

// ---- FP64 roofline denominator (observables: every pair term is accumulated in FP64) -------------
__global__ void __launch_bounds__(256) sn_dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int sn_bench_fp64_peak(int device, double *tflops)
{
    if (!tflops) return sn_fail(SN_ERR_INVALID, "sn_bench_fp64_peak: null");
    SN_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 1024;
    double *out;
    SN_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * blocks * 256));
    cudaEvent_t e0, e1;
    SN_CUDA_CHECK(cudaEventCreate(&e0));
    SN_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        SN_CUDA_CHECK(cudaEventRecord(e0));
        sn_dfma_peak_kernel<<<blocks, 256>>>(out, iters, 0.999, 0.001);
        SN_CUDA_CHECK(cudaEventRecord(e1));
        SN_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        SN_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * 256;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return SN_OK;
}

#include "sn_slab.cuh"

// CUDA loads kernels lazily, and loading one can need a context-wide synchronisation.  A slab's stream may hold a
// kernel that spins until a neighbouring slab has launched something (handshake, tile versions); if that something
// is launched for the first time in the process while the spinner runs, the load waits for the spinner and the
// spinner for the launch.  So every kernel of the library is loaded up front, once per device.
static int sn_preload_kernels(int device)
{
    static std::mutex m;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lock(m);
    if (std::find(done.begin(), done.end(), device) != done.end()) return SN_OK;
    const void *kernels[] = {
        (const void *)sn_phase_signal_kernel, (const void *)sn_phase_wait_kernel,
        (const void *)sn_pull_ghosts_kernel<true>, (const void *)sn_pull_ghosts_kernel<false>,
        (const void *)sn_scatter_kernel<true>, (const void *)sn_scatter_kernel<false>,
        (const void *)sn_gather_kernel<true>, (const void *)sn_gather_kernel<false>,
        (const void *)sn_convert_layout_kernel, (const void *)sn_refresh_ghosts_kernel, (const void *)sn_fill_u32_kernel,
        (const void *)sn_tiled_kernel<true, false, 3>, (const void *)sn_tiled_kernel<false, false, 3>, (const void *)sn_tiled_kernel<true, false, 2>, (const void *)sn_tiled_kernel<false, false, 2>,
        (const void *)sn_tiled_kernel<true, true, 3>, (const void *)sn_tiled_kernel<false, true, 3>, (const void *)sn_tiled_kernel<true, true, 2>, (const void *)sn_tiled_kernel<false, true, 2>,
        (const void *)sn_colour_pass_kernel<0, true>, (const void *)sn_colour_pass_kernel<0, false>,
        (const void *)sn_colour_pass_kernel<1, true>, (const void *)sn_colour_pass_kernel<1, false>,
        (const void *)sn_colour_pass_kernel<2, true>,
        (const void *)sn_resident_kernel<0, true>, (const void *)sn_resident_kernel<0, false>,
        (const void *)sn_resident_kernel<1, true>, (const void *)sn_resident_kernel<1, false>, (const void *)sn_resident_kernel<2, true>,
        (const void *)sn_sum_dipoles_kernel, (const void *)sn_state_hash_kernel, (const void *)sn_sum_doubles_kernel,
        (const void *)sn_energy_f32_kernel<0, true>, (const void *)sn_energy_f32_kernel<1, true>, (const void *)sn_energy_f32_kernel<2, true>,
        (const void *)sn_site_energy_f32_kernel<0, true>, (const void *)sn_site_energy_f32_kernel<1, true>, (const void *)sn_site_energy_f32_kernel<2, true>,
        (const void *)sn_rdf_tiled_kernel, (const void *)sn_potential_tiled_kernel, (const void *)sn_efield_tiled_kernel, (const void *)sn_reduce_rows_kernel,
        (const void *)sn_efield_kernel<true>, (const void *)sn_efield_kernel<false>, (const void *)sn_recombination_kernel,
    };
    for (const void *k : kernels) {
        cudaFuncAttributes a;
        SN_CUDA_CHECK(cudaFuncGetAttributes(&a, k));
    }
    int rc = sn_energy_exact_preload();
    if (rc) return rc;
    done.push_back(device);
    return SN_OK;
}

