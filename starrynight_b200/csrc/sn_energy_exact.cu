// sn_energy_exact.cu -- energy audit in the reference's own statement order.
//
// Compiled with -fmad=false so that no multiply-add is contracted: the two
// instantiations below then perform the same IEEE operations, in the same
// order, as site_energy (montecarlo-core.c:76-141):
//   REAL = float   SN_PREC_REPLICA  float terms, double accumulation -- bit-equal
//                                   to the native reference build;
//   REAL = double  SN_PREC_F64      the source under float->double -- the build the
//                                   1e-12 FP64 bar is asserted against.
// The fast FP32 local-field arithmetic (SN_PREC_F32) lives in sn_lib.cu.
#include "sn_common.cuh"

struct SnExactArgs {
    const float4 *lat;          // replica base (padded)
    SnGeom G;
    const int *nb_dxyz;         // reference order, 3 ints per neighbour
    int nnb;
    double cage, K;
    float Ex, Ey, Ez;
};

// config.c:189-199
template <class REAL>
__device__ __forceinline__ REAL sn_dot(const REAL ax, const REAL ay, const REAL az, const REAL bx, const REAL by, const REAL bz)
{
    REAL sum = (REAL)0.0;
    sum += ax * bx;
    sum += ay * by;
    sum += az * bz;
    return sum;
}

// montecarlo-core.c:76-141.  `which` selects the couplings that are switched on:
// bit0 dipole-dipole, bit1 cage strain, bit2 field, bit3 K.
template <class REAL>
__device__ double sn_site_energy_exact(const SnExactArgs &a, int x, int y, int z,
                                       REAL nx, REAL ny, REAL nz, REAL ox, REAL oy, REAL oz, REAL olen, int which)
{
    double dE = 0.0;
    const float4 *site = a.lat + sn_pidx(a.G, x, y, z);
    for (int i = 0; i < a.nnb; i++) {                                           // :91
        const int dx = a.nb_dxyz[3 * i], dy = a.nb_dxyz[3 * i + 1], dz = a.nb_dxyz[3 * i + 2];
        const REAL d = (REAL)sqrt((double)((REAL)dx * dx + dy * dy + dz * dz));  // :54 double sqrt, stored in REAL
        const float4 t4 = site[dx * a.G.sx + dy * a.G.sy + dz];                  // :97 via the ghost shell
        const REAL tx = t4.x, ty = t4.y, tz = t4.z, tl = t4.w;
        const REAL n_x = (REAL)dx / d, n_y = (REAL)dy / d, n_z = (REAL)dz / d;   // :99
        if (which & 1)
            dE += (olen * tl) *                                                  // :102-106
                  ((sn_dot<REAL>(nx, ny, nz, tx, ty, tz) - 3 * sn_dot<REAL>(n_x, n_y, n_z, nx, ny, nz) * sn_dot<REAL>(n_x, n_y, n_z, tx, ty, tz)) -
                   (sn_dot<REAL>(ox, oy, oz, tx, ty, tz) - 3 * sn_dot<REAL>(n_x, n_y, n_z, ox, oy, oz) * sn_dot<REAL>(n_x, n_y, n_z, tx, ty, tz))) /
                  (d * d * d);
        if ((which & 2) && (dx * dx + dy * dy + dz * dz) == 1)                   // :113-115
            dE += -a.cage * sn_dot<REAL>(nx, ny, nz, tx, ty, tz) + a.cage * sn_dot<REAL>(ox, oy, oz, tx, ty, tz);
    }
    if (which & 4) {                                                             // :120-121
        const REAL ex = a.Ex, ey = a.Ey, ez = a.Ez;
        dE += +sn_dot<REAL>(nx, ny, nz, ex, ey, ez) - sn_dot<REAL>(ox, oy, oz, ex, ey, ez);
    }
    if ((which & 8) && a.K > 0.0) {                                              // :124-134
        dE += -a.K * fabs((double)sn_dot<REAL>(nx, ny, nz, (REAL)1.0, (REAL)0.0, (REAL)0.0)) +
               a.K * fabs((double)sn_dot<REAL>(ox, oy, oz, (REAL)1.0, (REAL)0.0, (REAL)0.0));
        dE += -a.K * fabs((double)sn_dot<REAL>(nx, ny, nz, (REAL)0.0, (REAL)1.0, (REAL)0.0)) +
               a.K * fabs((double)sn_dot<REAL>(ox, oy, oz, (REAL)0.0, (REAL)1.0, (REAL)0.0));
    }
    return dE;
}

template <class REAL>
__global__ void sn_site_energy_exact_kernel(const SnExactArgs a, int n, const int *__restrict__ sites,
                                            const float *__restrict__ newdip, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = sites[3 * i], y = sites[3 * i + 1], z = sites[3 * i + 2];
    const float4 o = a.lat[sn_pidx(a.G, x, y, z)];
    out[i] = sn_site_energy_exact<REAL>(a, x, y, z, (REAL)newdip[3 * i], (REAL)newdip[3 * i + 1], (REAL)newdip[3 * i + 2],
                                        (REAL)o.x, (REAL)o.y, (REAL)o.z, (REAL)o.w, 15);
}

// per-site interaction energy e_i = site_energy(new = p_i, old = {0,0,0,len_i}) restricted to `which`
template <class REAL>
__global__ void sn_interaction_map_exact_kernel(const SnExactArgs a, int which, double *__restrict__ out)
{
    const long long n = (long long)a.G.X * a.G.Y * a.G.nz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int z = (int)(i % a.G.nz), y = (int)((i / a.G.nz) % a.G.Y), x = (int)(i / ((long long)a.G.nz * a.G.Y));
    const float4 o = a.lat[sn_pidx(a.G, x, y, z)];
    out[i] = sn_site_energy_exact<REAL>(a, x, y, z, (REAL)o.x, (REAL)o.y, (REAL)o.z,
                                        (REAL)0.0, (REAL)0.0, (REAL)0.0, (REAL)o.w, which);
}

static SnExactArgs sn_exact_args(sn_handle *h, int replica)
{
    SnExactArgs a;
    a.lat = h->lat + (long long)replica * h->G.rep_stride;
    a.G = h->G;
    a.nb_dxyz = h->d_nb_dxyz;
    a.nnb = h->nnb;
    a.cage = h->h_cage[replica]; a.K = h->p.K;
    a.Ex = h->h_efield[3 * replica]; a.Ey = h->h_efield[3 * replica + 1]; a.Ez = h->h_efield[3 * replica + 2];
    return a;
}

int sn_energy_exact_launch(sn_handle *h, int replica, int precision, int n, const int *d_sites,
                           const float *d_newdip, double *d_out)
{
    const SnExactArgs a = sn_exact_args(h, replica);
    const int bs = 128, gs = (n + bs - 1) / bs;
    if (n == 0) return SN_OK;
    if (precision == SN_PREC_F64) sn_site_energy_exact_kernel<double><<<gs, bs, 0, h->stream>>>(a, n, d_sites, d_newdip, d_out);
    else sn_site_energy_exact_kernel<float><<<gs, bs, 0, h->stream>>>(a, n, d_sites, d_newdip, d_out);
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

// which: bitmask as above; d_out has X*Y*nz doubles
int sn_energy_exact_map_launch(sn_handle *h, int replica, int precision, int which, double *d_out)
{
    const SnExactArgs a = sn_exact_args(h, replica);
    const long long n = (long long)h->G.X * h->G.Y * h->G.nz;
    const int bs = 128; const long long gs = (n + bs - 1) / bs;
    if (precision == SN_PREC_F64) sn_interaction_map_exact_kernel<double><<<(unsigned)gs, bs, 0, h->stream>>>(a, which, d_out);
    else sn_interaction_map_exact_kernel<float><<<(unsigned)gs, bs, 0, h->stream>>>(a, which, d_out);
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

// see sn_preload_kernels (sn_lib.cu)
int sn_energy_exact_preload()
{
    const void *kernels[] = {(const void *)sn_site_energy_exact_kernel<double>, (const void *)sn_site_energy_exact_kernel<float>,
                             (const void *)sn_interaction_map_exact_kernel<double>, (const void *)sn_interaction_map_exact_kernel<float>};
    for (const void *k : kernels) {
        cudaFuncAttributes a;
        SN_CUDA_CHECK(cudaFuncGetAttributes(&a, k));
    }
    return SN_OK;
}
