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
            if (sn_globaltimer_ns() - t0 > timeout_ns) { atomicExch(flags + SN_FLAGS_ERR, 1u); return; }   // the neighbour never arrived
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
