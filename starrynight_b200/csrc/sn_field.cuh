// sn_field.cuh -- the fused FP32 dE of one Metropolis attempt.
//
// site_energy (montecarlo-core.c:76-141) evaluates, per neighbour j,
//     l_i l_j [ (p'.p_j - 3 (n.p')(n.p_j)) - (p.p_j - 3 (n.p)(n.p_j)) ] / d^3
// which is linear in dp = p' - p.  We gather the proposal-independent local field
//     F = sum_j T(r_j) (l_j p_j),   T(r) = (I - 3 n n^T) / d^3          (dipole-dipole)
//     G = sum_{|r_j|=1} p_j                                               (cage strain, :113-115)
// once per attempt and fuse all four terms into
//     dE = l_i dp.F - CageStrain dp.G + dp.E - K (|p'_x|-|p_x| + |p'_y|-|p_y|) [K>0]
// For DipoleCutOff = 3 the 122 (28 when Z==1) tensors T(r) are compile-time
// constants folded into FFMA immediates and zero components cost nothing
// (798 instead of 1098 FFMA in 3-D).  Other cut-offs use the table in SnNbEntry.
#pragma once

#include <type_traits>
#include "sn_common.cuh"

template <int B, int E, class F>
__device__ __forceinline__ void sn_static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        sn_static_for<B + 1, E>(f);
    }
}

// 1/d^3 for the integer r^2 that occur inside a radius-3 sphere
__host__ __device__ constexpr double sn_inv_d3(int r2)
{
    return r2 == 1 ? 1.0 : r2 == 2 ? 0.35355339059327376220 : r2 == 3 ? 0.19245008972987525484 :
           r2 == 4 ? 0.125 : r2 == 5 ? 0.08944271909999158786 : r2 == 6 ? 0.06804138174397716939 :
           r2 == 8 ? 0.04419417382415922028 : r2 == 9 ? 0.03703703703703703704 : 0.0;
}

// T_ab(r) = (delta_ab r^2 - 3 r_a r_b) / (r^2 d^3), rounded once to float.  The integer numerator
// makes vanishing entries exactly zero (e.g. the diagonal of r = (1,1,1)), so they cost no FFMA.
__host__ __device__ constexpr float sn_T(int dx, int dy, int dz, int a, int b)
{
    const int r2 = dx * dx + dy * dy + dz * dz;
    const int ra = a == 0 ? dx : a == 1 ? dy : dz;
    const int rb = b == 0 ? dx : b == 1 ? dy : dz;
    const int num = (a == b ? r2 : 0) - 3 * ra * rb;
    return num == 0 ? 0.0f : (float)((double)num * sn_inv_d3(r2) / r2);
}

// accumulate neighbour m = (p_j, l_j) at compile-time offset (DX,DY,DZ) into F (and G when |r|=1)
template <int DX, int DY, int DZ, bool SPECIES>
__device__ __forceinline__ void sn_accumulate(float3 &F, float3 &G, const float4 m)
{
    constexpr float txx = sn_T(DX, DY, DZ, 0, 0), tyy = sn_T(DX, DY, DZ, 1, 1), tzz = sn_T(DX, DY, DZ, 2, 2);
    constexpr float txy = sn_T(DX, DY, DZ, 0, 1), txz = sn_T(DX, DY, DZ, 0, 2), tyz = sn_T(DX, DY, DZ, 1, 2);
    float ax = m.x, ay = m.y, az = m.z;
    if constexpr (SPECIES) { ax *= m.w; ay *= m.w; az *= m.w; }     // moment l_j p_j (montecarlo-core.c:102)
    if constexpr (txx != 0.0f) F.x = fmaf(txx, ax, F.x);
    if constexpr (tyy != 0.0f) F.y = fmaf(tyy, ay, F.y);
    if constexpr (tzz != 0.0f) F.z = fmaf(tzz, az, F.z);
    if constexpr (txy != 0.0f) { F.x = fmaf(txy, ay, F.x); F.y = fmaf(txy, ax, F.y); }
    if constexpr (txz != 0.0f) { F.x = fmaf(txz, az, F.x); F.z = fmaf(txz, ax, F.z); }
    if constexpr (tyz != 0.0f) { F.y = fmaf(tyz, az, F.y); F.z = fmaf(tyz, ay, F.z); }
    if constexpr (DX * DX + DY * DY + DZ * DZ == 1) { G.x += m.x; G.y += m.y; G.z += m.z; }
}

// Whole cut-off-3 sphere through a loader `load(dx,dy,dz) -> float4`, in the
// reference's table order dx (outer), dy, dz (inner) (montecarlo-core.c:47-49).
template <bool FLAT, bool SPECIES, class Load>
__device__ __forceinline__ void sn_local_field_cut3(Load &&load, float3 &F, float3 &G)
{
    sn_static_for<-3, 4>([&](auto dx) {
        sn_static_for<-3, 4>([&](auto dy) {
            sn_static_for<(FLAT ? 0 : -3), (FLAT ? 1 : 4)>([&](auto dz) {
                constexpr int DX = decltype(dx)::value, DY = decltype(dy)::value, DZ = decltype(dz)::value;
                constexpr int r2 = DX * DX + DY * DY + DZ * DZ;
                if constexpr (r2 > 0 && r2 <= 9)
                    sn_accumulate<DX, DY, DZ, SPECIES>(F, G, load(DX, DY, DZ));
            });
        });
    });
}

// table-driven variant for any cut-off
template <class Load>
__device__ __forceinline__ void sn_local_field_table(const SnNbEntry *__restrict__ nb, int nnb, Load &&load,
                                                     float3 &F, float3 &G)
{
    for (int i = 0; i < nnb; i++) {
        const SnNbEntry e = nb[i];
        const float4 m = load(e.dx, e.dy, e.dz);
        const float ax = m.x * m.w, ay = m.y * m.w, az = m.z * m.w;
        F.x = fmaf(e.txx, ax, fmaf(e.txy, ay, fmaf(e.txz, az, F.x)));
        F.y = fmaf(e.txy, ax, fmaf(e.tyy, ay, fmaf(e.tyz, az, F.y)));
        F.z = fmaf(e.txz, ax, fmaf(e.tyz, ay, fmaf(e.tzz, az, F.z)));
        if (e.nn) { G.x += m.x; G.y += m.y; G.z += m.z; }
    }
}

struct SnTerms {            // couplings shared by every site of a replica
    float cage, K, beta;
    float3 E;
    int constrain, dim;
};

// fused dE for old -> (nx,ny,nz); montecarlo-core.c:102-134 in local-field form
__device__ __forceinline__ float sn_delta_e(const float4 old, const float3 np, const float3 F, const float3 G,
                                            const SnTerms &t)
{
    const float dx = np.x - old.x, dy = np.y - old.y, dz = np.z - old.z;
    float dE = old.w * (dx * F.x + dy * F.y + dz * F.z);
    dE -= t.cage * (dx * G.x + dy * G.y + dz * G.z);
    dE += dx * t.E.x + dy * t.E.y + dz * t.E.z;
    if (t.K > 0.0f) dE -= t.K * ((fabsf(np.x) - fabsf(old.x)) + (fabsf(np.y) - fabsf(old.y)));
    return dE;
}

// Trial orientation from two uniforms (replaces random_sphere_point /
// random_X_point, config.c:203-263): uniform on S^2 via z = 1-2u, phi = 2 pi v
// (Archimedes), on the XY circle when DIM < 3, or one of the six <100> vectors.
__device__ __forceinline__ float3 sn_propose(const SnTerms &t, float u, float v)
{
    float3 p;
    if (t.constrain) {
        const int i = min((int)(u * 6.0f), 5);
        const float s = (i & 1) ? -1.0f : 1.0f;
        p.x = (i >> 1) == 0 ? s : 0.0f; p.y = (i >> 1) == 1 ? s : 0.0f; p.z = (i >> 1) == 2 ? s : 0.0f;
        return p;
    }
    float sn, cs;
    __sincosf(6.283185307179586f * v, &sn, &cs);
    if (t.dim < 3) { p.x = cs; p.y = sn; p.z = 0.0f; return p; }
    const float z = 1.0f - 2.0f * u;
    const float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    p.x = r * cs; p.y = r * sn; p.z = z;
    return p;
}

// Metropolis test, montecarlo-core.c:179: accept iff dE < 0 or exp(-dE beta) > u.
// beta = +inf (T = 0) with dE == 0 gives NaN and rejects, as in the reference.
__device__ __forceinline__ bool sn_accept(float dE, float beta, float u)
{
    return (dE < 0.0f) | (__expf(-dE * beta) > u);      // no short-circuit: keeps the warp converged
}
