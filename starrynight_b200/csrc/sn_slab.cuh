// sn_slab.cuh -- Z-slab decomposition: ghost-plane bootstrap through the host,
// CUDA-IPC / peer wiring, and the per-phase device-side handshake.
//
// The reference has no distribution at all.  Here every GPU
// owns nz = Z/G planes plus `cutoff` ghost planes on each side.  The sweep
// kernels push every accepted boundary update straight into the neighbours'
// ghost planes with P2P stores over NVLink (sn_store_site); between two phases a
// one-thread kernel publishes "phase e done" to both neighbours and the next
// phase's kernels are held back by a one-thread kernel that waits for the two
// flags.  Everything is stream-ordered; the host never blocks.
//
// Included at the end of sn_lib.cu.
#pragma once

__global__ void sn_phase_signal_kernel(unsigned int *to_lower, unsigned int *to_upper, unsigned int epoch)
{
    __threadfence_system();
    if (to_lower) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_lower), "r"(epoch) : "memory"); }
    if (to_upper) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_upper), "r"(epoch) : "memory"); }
}

__global__ void sn_phase_wait_kernel(unsigned int *flags, unsigned int epoch, unsigned long long timeout_ns)
{
    const unsigned long long t0 = sn_globaltimer_ns();
    for (int s = 0; s < 2; s++) {
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + s) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (sn_globaltimer_ns() - t0 > timeout_ns) {                   // the neighbour never arrived
                printf("starrynight_b200: slab handshake timed out: waiting for epoch %u, neighbour %s is at %u\n", epoch, s ? "above" : "below", v);
                atomicExch(flags + SN_FLAGS_ERR, 1u);
                return;
            }
            __nanosleep(200);
        }
    }
}

// after a phase that may have written into the neighbours' ghost planes
static int sn_slab_phase_sync(sn_handle *h, long long *launches)
{
    h->phase_epoch++;
    // my lower neighbour reads my signal in its slot 1 ("from above"), my upper neighbour in its slot 0
    sn_phase_signal_kernel<<<1, 1, 0, h->stream>>>(h->peer_flags[0] ? h->peer_flags[0] + 1 : nullptr,
                                                   h->peer_flags[1] ? h->peer_flags[1] + 0 : nullptr, h->phase_epoch);
    sn_phase_wait_kernel<<<1, 1, 0, h->stream>>>(h->flags, h->phase_epoch, h->spin_timeout_ns);
    SN_CUDA_CHECK(cudaGetLastError());
    if (launches) *launches += 2;
    return SN_OK;
}

static int sn_copy_planes(sn_handle *h, int replica, int zfirst, float *host, bool to_device)
{
    const SnGeom &G = h->G;
    cudaMemcpy3DParms c;
    memset(&c, 0, sizeof c);
    float4 *dev = h->lat + (long long)replica * G.rep_stride + sn_pidx(G, 0, 0, zfirst);
    cudaPitchedPtr hp = make_cudaPitchedPtr(host, (size_t)G.gz * 16, (size_t)G.gz * 16, G.Y);
    cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)G.PZ * 16, (size_t)G.PZ * 16, G.PY);
    c.srcPtr = to_device ? hp : dp; c.dstPtr = to_device ? dp : hp;
    c.extent = make_cudaExtent((size_t)G.gz * 16, G.Y, G.X);
    c.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    SN_CUDA_CHECK(cudaMemcpy3DAsync(&c, h->stream));
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SN_OK;
}

extern "C" int sn_get_boundary(sn_handle *h, int replica, int side, float *planes)
{
    SN_CHECK_HANDLE(h, replica);
    if (!planes || side < 0 || side > 1) return sn_fail(SN_ERR_INVALID, "sn_get_boundary: bad arguments");
    if (h->G.gz == 0) return sn_fail(SN_ERR_UNSUPPORTED, "sn_get_boundary: lattice has no interacting Z axis");
    { int rc = sn_sync_canonical(h); if (rc) return rc; }
    return sn_copy_planes(h, replica, side == 0 ? 0 : h->G.nz - h->G.gz, planes, false);
}

extern "C" int sn_set_ghost(sn_handle *h, int replica, int side, const float *planes)
{
    SN_CHECK_HANDLE(h, replica);
    if (!planes || side < 0 || side > 1) return sn_fail(SN_ERR_INVALID, "sn_set_ghost: bad arguments");
    if (h->G.periodic_z) return sn_fail(SN_ERR_INVALID, "sn_set_ghost: handle owns the whole Z axis; its ghosts are its own periodic images");
    int rc = sn_sync_canonical(h);
    if (rc) return rc;
    h->lat2_valid = false;
    if ((rc = sn_copy_planes(h, replica, side == 0 ? -h->G.gz : h->G.nz, const_cast<float *>(planes), true))) return rc;
    if ((rc = sn_refresh_ghosts(h))) return rc;      // x / y images of the new planes
    SN_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return SN_OK;
}

// Fill my z ghost planes (x / y ghost columns included) from the neighbours' boundary planes, device to device.
template <bool TILED>
__global__ void __launch_bounds__(256) sn_pull_ghosts_kernel(float4 *__restrict__ mine, const float4 *__restrict__ lo, const float4 *__restrict__ hi,
                                                             const SnGeom G, const long long rep_stride, const int nrep)
{
    const int PX = G.X + 2 * G.g;
    const long long per_rep = (long long)PX * G.PY * 2 * G.gz, total = per_rep * nrep;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int rep = (int)(i / per_rep);
        long long r = i - (long long)rep * per_rep;
        const int dz = (int)(r % G.gz); r /= G.gz;
        const int side = (int)(r & 1); r >>= 1;
        const int yp = (int)(r % G.PY) - G.g, xp = (int)(r / G.PY) - G.g;
        const int zdst = side == 0 ? dz - G.gz : G.nz + dz;              // my ghost plane
        const int zsrc = side == 0 ? G.nz - G.gz + dz : dz;              // the same plane among the neighbour's own
        const float4 *src = (side == 0 ? lo : hi) + (long long)rep * rep_stride;
        float4 *dst = mine + (long long)rep * rep_stride;
        if (TILED) sn_st2(dst, G, sn_pidx2(G, xp, yp, zdst), sn_ld2(src, G, sn_pidx2(G, xp, yp, zsrc)));
        else dst[sn_pidx(G, xp, yp, zdst)] = src[sn_pidx(G, xp, yp, zsrc)];
    }
}

// After every slab of the lattice has been uploaded (sn_set_lattice) and wired (sn_ipc_attach / sn_attach_peer):
// each slab copies its neighbours' boundary planes into its own ghost planes over NVLink.  Bracketed by the
// device-side handshake: nobody reads a neighbour before that neighbour's upload has landed, nobody sweeps (and
// changes its boundary) before its neighbours have read it.  Stream-ordered; call it on every slab.
extern "C" int sn_pull_ghosts(sn_handle *h)
{
    SN_CHECK_HANDLE(h, 0);
    if (h->G.periodic_z) return SN_OK;                 // the handle owns the whole axis: its ghosts are its own images
    if (!h->peer_lat[0] || !h->peer_lat[1]) return sn_fail(SN_ERR_INVALID, "sn_pull_ghosts: no neighbours attached (sn_ipc_attach / sn_attach_peer)");
    int rc;
    if (h->use_tiled) {
        if (!h->lat2_valid) {                          // neighbours read the tiled copy: bring it up to date first
            if ((rc = sn_sync_canonical(h)) || (rc = sn_convert_layout(h, true))) return rc;
            h->lat2_valid = true;
        }
    } else if ((rc = sn_sync_canonical(h))) return rc;
    if ((rc = sn_slab_phase_sync(h, nullptr))) return rc;
    const long long total = (long long)(h->G.X + 2 * h->G.g) * h->G.PY * 2 * h->G.gz * h->p.nreplicas;
    const int nblocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8);
    if (h->use_tiled) {
        sn_pull_ghosts_kernel<true><<<nblocks, 256, 0, h->stream>>>(h->lat2, h->peer_lat[0], h->peer_lat[1], h->G, sn_rep_stride2(h->G), h->p.nreplicas);
        h->lat_valid = false;
    } else {
        sn_pull_ghosts_kernel<false><<<nblocks, 256, 0, h->stream>>>(h->lat, h->peer_lat[0], h->peer_lat[1], h->G, h->G.rep_stride, h->p.nreplicas);
        h->lat2_valid = false;
    }
    SN_CUDA_CHECK(cudaGetLastError());
    return sn_slab_phase_sync(h, nullptr);
}

extern "C" int sn_ipc_export(sn_handle *h, void *lattice_handle64, void *flags_handle64)
{
    SN_CHECK_HANDLE(h, 0);
    if (!lattice_handle64 || !flags_handle64) return sn_fail(SN_ERR_INVALID, "sn_ipc_export: null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t a, b;
    // neighbours push into the array the sweep kernel reads: the de-interleaved copy for the tiled kernel
    SN_CUDA_CHECK(cudaIpcGetMemHandle(&a, h->use_tiled ? h->lat2 : h->lat));
    SN_CUDA_CHECK(cudaIpcGetMemHandle(&b, h->flags));
    memcpy(lattice_handle64, &a, 64); memcpy(flags_handle64, &b, 64);
    return SN_OK;
}

extern "C" int sn_ipc_attach(sn_handle *h, int side, const void *lattice_handle64, const void *flags_handle64)
{
    SN_CHECK_HANDLE(h, 0);
    if (side < 0 || side > 1 || !lattice_handle64 || !flags_handle64) return sn_fail(SN_ERR_INVALID, "sn_ipc_attach: bad arguments");
    if (side == 1 && h->peer_is_ipc[0] && !memcmp(h->ipc_key[0], lattice_handle64, 64)) {
        // two-GPU ring: the lower and the upper neighbour are the same allocation
        h->peer_lat[1] = h->peer_lat[0]; h->peer_flags[1] = h->peer_flags[0]; h->peer_is_ipc[1] = true;
        memcpy(h->ipc_key[1], lattice_handle64, 64);
        return SN_OK;
    }
    cudaIpcMemHandle_t a, b;
    memcpy(&a, lattice_handle64, 64); memcpy(&b, flags_handle64, 64);
    void *pl = nullptr, *pf = nullptr;
    SN_CUDA_CHECK(cudaIpcOpenMemHandle(&pl, a, cudaIpcMemLazyEnablePeerAccess));
    SN_CUDA_CHECK(cudaIpcOpenMemHandle(&pf, b, cudaIpcMemLazyEnablePeerAccess));
    {
        // the kernels index the neighbour's arrays with MY geometry: refuse a neighbour of another shape or kernel
        unsigned int theirs[SN_DESC_WORDS], mine[SN_DESC_WORDS];
        sn_slab_descriptor(h, mine);
        SN_CUDA_CHECK(cudaMemcpy(theirs, (unsigned int *)pf + SN_FLAGS_DESC, sizeof theirs, cudaMemcpyDefault));
        if (memcmp(theirs, mine, sizeof mine)) {
            cudaIpcCloseMemHandle(pl); cudaIpcCloseMemHandle(pf);
            return sn_fail(SN_ERR_INVALID, "sn_ipc_attach: the neighbour is not a slab of the same decomposition (X,Y,nz,replicas,cutoff,kernel = "
                                           "%u,%u,%u,%u,%u,%u here, %u,%u,%u,%u,%u,%u there)", mine[1], mine[2], mine[3], mine[4], mine[5], mine[6],
                           theirs[1], theirs[2], theirs[3], theirs[4], theirs[5], theirs[6]);
        }
    }
    h->peer_lat[side] = (float4 *)pl; h->peer_flags[side] = (unsigned int *)pf; h->peer_is_ipc[side] = true;
    memcpy(h->ipc_key[side], lattice_handle64, 64);
    return SN_OK;
}

extern "C" int sn_attach_peer(sn_handle *h, int side, sn_handle *peer)
{
    SN_CHECK_HANDLE(h, 0);
    if (side < 0 || side > 1 || !peer) return sn_fail(SN_ERR_INVALID, "sn_attach_peer: bad arguments");
    if (peer->G.X != h->G.X || peer->G.Y != h->G.Y || peer->G.nz != h->G.nz || peer->p.nreplicas != h->p.nreplicas)
        return sn_fail(SN_ERR_INVALID, "sn_attach_peer: slabs must have identical shapes");
    if (peer->p.device != h->p.device) {
        int can = 0;
        SN_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, h->p.device, peer->p.device));
        if (!can) return sn_fail(SN_ERR_UNSUPPORTED, "sn_attach_peer: device %d cannot access device %d", h->p.device, peer->p.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->p.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sn_fail(SN_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    if (peer->use_tiled != h->use_tiled) return sn_fail(SN_ERR_INVALID, "sn_attach_peer: slabs must run the same sweep kernel");

    h->peer_lat[side] = h->use_tiled ? peer->lat2 : peer->lat; h->peer_flags[side] = peer->flags; h->peer_is_ipc[side] = false;
    return SN_OK;
}
