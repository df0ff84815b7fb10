// sn_sweep_tiled.cuh -- the fast Metropolis sweep: TMA-staged shared-memory tiles.
//
// Replaces MC_moves -> MC_move -> site_energy (montecarlo-core.c:76-191) for
// DipoleCutOff = 3 (or 2) lattices with X, Y, Z >= 20 and Z a multiple of 4 (Z-slab handles: multiples of 32 planes); the last
// tile of an axis may be partial, an odd number of tiles along an axis gets a third tile colour.
//
// Decomposition
//   * The lattice is cut into 16^3 tiles.  A sweep is 8 "phases", one per tile
//     parity (px,py,pz): active tiles are 32 apart, so the 22^3 read set of one
//     active tile never meets the 16^3 write set of another.
//   * ONE launch runs any number of sweeps as a dataflow over work items
//     (sweep, phase, tile) taken in that order from a global counter.  There is no
//     barrier between phases: every tile carries a version (= sweeps it has
//     completed) and an item starts as soon as its 26 neighbouring tiles have
//     reached the version that precedes it in the global order (neighbours of an
//     earlier phase: this sweep done; of a later phase: previous sweep done).
//     Adjacent tiles are thereby totally ordered -- the result is bit-identical to
//     8 barrier-separated launches per sweep -- distant tiles run freely and the
//     phases overlap at their tails.  Z-slab handles see the neighbouring GPUs'
//     boundary tile layers as ghost entries of the version array, written by the
//     peers over NVLink, so there is no cross-GPU barrier either.
//   * One persistent CTA per SM.  Per tile, one thread issues cp.async.bulk.tensor (TMA) loads of the
//     22 x 22 x 28(z) halo box from the split copy of the lattice (sn_common.cuh: xy as float2, z and the
//     lengths as float arrays, natural z order): one box per array, TMA rows of 224 / 112 B.  The lengths are
//     loaded only when some site has a length != 1.  A thread owns two consecutive z sites that start at an
//     even plane, so it reads two planes of a neighbour column at a time: one LDS.128 (x,y of both planes) and
//     one LDS.64 (both z) -- 12 bytes per site through the shared-memory pipe, no bank conflicts (the 8 lanes
//     of a quarter-warp read 8 consecutive 16-byte pairs, the 16 lanes of a half-warp 16 distinct 8-byte
//     pairs), and every neighbour address is base + compile-time immediate.
//   * Inside a tile the 64 site colours are visited as 16 super-passes (cx,cy).
//     Two lanes own a segment of 4 consecutive z sites of one (x,y) column, i.e.
//     the four colours (cx,cy,0..3), 2 sites each.  None of the 28 neighbour
//     columns around it changes during the super-pass (they belong to other
//     (cx,cy) classes), so their contribution to the local fields is gathered
//     once with a sliding z window (one load serves both sites).  Only the centre
//     column changes: the 4 sites are then decided in sequence, the fields of the
//     later ones corrected in registers for the earlier accepted moves (partner
//     lane and the segment above via shuffles).  Every attempt is a full fresh dE
//     over the cut-off sphere.
//   * Two worker roles of 4 warps: role A gathers the columns that depend on the
//     previous super-pass and runs the chain, role B gathers the rest one super-pass
//     ahead and draws the proposals; a ninth warp schedules (work items,
//     dependencies, TMA, version publishing).  A thread gathers the fields of 2
//     sites; neighbours r and -r share the tensor T(r) = T(-r), so their moments are
//     added first and the tensor applied once (582 instead of 798 FP ops per
//     attempt).  The unrolled streams are shared by 4 warps each, which keeps
//     instruction fetch off the critical path (profiles/experiments/ has what
//     happens otherwise).
//   * The tile's interior is written back once per tile, coalesced, under the next
//     tile's TMA load; sites on a lattice / slab face also go to their ghost images
//     (periodic faces, and the neighbouring GPU's ghost planes over NVLink for a
//     Z-slab handle).
//
// SN_EXP_* macros switch single ingredients off (no loads, no FP, no chain, ...) for the
// bottleneck experiments recorded in profiles/experiments/; they are never defined in a
// product build.
#pragma once

#include <cuda.h>

#include "sn_field.cuh"
#include "sn_sweep_colour.cuh"

#ifndef SN_CTRL_POLL_NS
#define SN_CTRL_POLL_NS 400             // pause between two looks at an unmet dependency (a tile takes ~27 us)
#endif

namespace snt {
constexpr int T = 16;                    // tile edge
constexpr int H = 3;                     // halo = cut-off
constexpr int BX = T + 2 * H;            // 22 columns per axis in the box
constexpr int NP = 14;                   // plane pairs per column of the box (28 planes: z0-4 .. z0+23)
constexpr int XY_BYTES = BX * BX * NP * 16;   // 108416: (x,y) of the box, float4 = two planes
constexpr int Z_BYTES = BX * BX * NP * 8;     // 54208: z (and, in its own box, the lengths), float2 = two planes
constexpr int OFF_Z = XY_BYTES;               // 847 * 128 (TMA destinations are 128-byte aligned)
constexpr int OFF_L = OFF_Z + 54272;          // Z_BYTES rounded up to 128
constexpr int OFF_END = OFF_L + 54272;
constexpr int NCOL = 14;                  // neighbour column pairs {(dx,dy), (-dx,-dy)} inside the cut-off disc
constexpr int SITE_THREADS = 128;        // threads of one role; each owns 2 sites per super-pass
constexpr int OFF_XF = OFF_END;                 // partial fields handed from role B to role A: float2[2 buffers][3][128]
constexpr int OFF_XP = OFF_XF + 2 * 3 * SITE_THREADS * 8;      // proposals drawn by role B: float4[2 buffers][2 sites][128]
constexpr int OFF_BAR = OFF_XP + 2 * 2 * SITE_THREADS * 16;    // mbarrier
constexpr int OFF_CTL = OFF_BAR + 16;                          // control warp hand-over: two SnTileItem slots (48 B each) + two arrival counters
constexpr int SMEM_BYTES = OFF_CTL + 112;
constexpr int WORKERS = 256;                // 2 roles x 4 warps
constexpr int THREADS = WORKERS + 32;      // + the control warp
// Column pairs (snt::col numbering) gathered by role A.  Role B works one super-pass ahead of role A's
// chain, so it may not read a column of the class A is updating: seen from the next class (cx,cy+1)
// those are the columns (0,-1), (0,3), and (-1,-1) when cx advances, i.e. pairs 0, 2 and 6.  Pairs 0 and
// 5 are the in-plane cage-strain neighbours, kept with A so that B hands over dipole fields only.
#ifndef SN_MASK_A_EXTRA
#define SN_MASK_A_EXTRA 0u              // further column pairs moved from role B to role A (balance experiments)
#endif
constexpr unsigned MASK_A = (1u << 0) | (1u << 2) | (1u << 5) | (1u << 6) | (SN_MASK_A_EXTRA);
constexpr unsigned MASK_B = ((1u << NCOL) - 1u) & ~MASK_A;

__host__ __device__ constexpr int half_height(int r2xy) { return 9 - r2xy >= 9 ? 3 : 9 - r2xy >= 4 ? 2 : 9 - r2xy >= 1 ? 1 : 0; }
__host__ __device__ constexpr int residue(int e) { return ((e % 4) + 4) % 4; }
__host__ __device__ constexpr int qshift(int e) { return (e - residue(e)) / 4; }      // floor(e / 4)
}  // namespace snt

__device__ __forceinline__ uint32_t sn_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sn_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// Field accumulators of the thread's two sites: nine independent partial sums per site (one per tensor entry),
// all held as register pairs so that the arithmetic is issued as packed FP32 instructions (Blackwell FADD2 /
// FFMA2: two IEEE FP32 operations per issue slot, each half rounded exactly like the scalar FADD / FFMA, so the
// chain is bit-identical to the scalar instruction stream).  The (x, y) half of a neighbour's moment pairs up
// inside a site; the z half pairs up across the two sites, which see the same tensor T(dx,dy,dz) one plane
// apart.  The kernel is bound by issue slots and shared-memory wavefronts, not by the FMA pipe: a full tensor
// applied to both sites takes 12-13 slots instead of 24.
struct SnAcc2 {
    float2 d[2];                                // per site (xx, yy):  F.x += Txx ax,  F.y += Tyy ay
    float2 o[2];                                // per site (yx, xy):  F.y += Txy ax,  F.x += Txy ay
    float2 r[2];                                // per site (zx, zy):  F.z += Txz ax,  F.z += Tyz ay
    float2 xz, yz, zz;                          // (site 0, site 1):   F.x += Txz az, F.y += Tyz az, F.z += Tzz az
    float2 gxy[2], gz;                          // cage-strain sum: (gx, gy) per site, gz of (site 0, site 1)
};

#ifdef SN_NO_PACKED_FP32      // scalar instruction stream (same arithmetic), kept for A/B timing
__device__ __forceinline__ float2 sn_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sn_fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 sn_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sn_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif

// One neighbour pair r = (DX,DY,DZ) and -r for both sites of the thread (site 1 lies one plane above site 0, so
// its neighbours a1 / b1 are the planes after a0 / b0).  T(r) = T(-r): the two moments are added first and the
// tensor applied once.  ZPAIR: (a0.z, a1.z) and (b0.z, b1.z) are register pairs as loaded (one LDS.64 each),
// so their sum is one FADD2; otherwise two FADDs write the halves of the pair.
template <int DX, int DY, int DZ, bool SPECIES, bool ZPAIR, int CUT>
__device__ __forceinline__ void sn_accumulate_pair2(SnAcc2 &A, const float4 a0, const float4 b0, const float4 a1, const float4 b1)
{
    if constexpr (DX * DX + DY * DY + DZ * DZ > CUT * CUT) return;      // beyond the handle's DipoleCutOff (the tile's halo is always 3 wide)
    else {
    constexpr float txx = sn_T(DX, DY, DZ, 0, 0), tyy = sn_T(DX, DY, DZ, 1, 1), tzz = sn_T(DX, DY, DZ, 2, 2);
    constexpr float txy = sn_T(DX, DY, DZ, 0, 1), txz = sn_T(DX, DY, DZ, 0, 2), tyz = sn_T(DX, DY, DZ, 1, 2);
    float2 axy[2], az;
    if constexpr (SPECIES) {
        axy[0] = make_float2(fmaf(a0.x, a0.w, b0.x * b0.w), fmaf(a0.y, a0.w, b0.y * b0.w));
        axy[1] = make_float2(fmaf(a1.x, a1.w, b1.x * b1.w), fmaf(a1.y, a1.w, b1.y * b1.w));
        az = make_float2(fmaf(a0.z, a0.w, b0.z * b0.w), fmaf(a1.z, a1.w, b1.z * b1.w));
    } else {
        axy[0] = sn_add2(make_float2(a0.x, a0.y), make_float2(b0.x, b0.y));
        axy[1] = sn_add2(make_float2(a1.x, a1.y), make_float2(b1.x, b1.y));
        if constexpr (ZPAIR) az = sn_add2(make_float2(a0.z, a1.z), make_float2(b0.z, b1.z));
        else az = make_float2(a0.z + b0.z, a1.z + b1.z);
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        if constexpr (txx != 0.0f && tyy != 0.0f) A.d[s] = sn_fma2(make_float2(txx, tyy), axy[s], A.d[s]);
        else if constexpr (txx != 0.0f) A.d[s].x = fmaf(txx, axy[s].x, A.d[s].x);
        else if constexpr (tyy != 0.0f) A.d[s].y = fmaf(tyy, axy[s].y, A.d[s].y);
        if constexpr (txy != 0.0f) A.o[s] = sn_fma2(make_float2(txy, txy), axy[s], A.o[s]);
        if constexpr (txz != 0.0f && tyz != 0.0f) A.r[s] = sn_fma2(make_float2(txz, tyz), axy[s], A.r[s]);
        else if constexpr (txz != 0.0f) A.r[s].x = fmaf(txz, axy[s].x, A.r[s].x);
        else if constexpr (tyz != 0.0f) A.r[s].y = fmaf(tyz, axy[s].y, A.r[s].y);
    }
    if constexpr (tzz != 0.0f) A.zz = sn_fma2(make_float2(tzz, tzz), az, A.zz);
    if constexpr (txz != 0.0f) A.xz = sn_fma2(make_float2(txz, txz), az, A.xz);
    if constexpr (tyz != 0.0f) A.yz = sn_fma2(make_float2(tyz, tyz), az, A.yz);
    if constexpr (DX * DX + DY * DY + DZ * DZ == 1) {
        if constexpr (SPECIES) {
            A.gxy[0] = sn_add2(A.gxy[0], sn_add2(make_float2(a0.x, a0.y), make_float2(b0.x, b0.y)));
            A.gxy[1] = sn_add2(A.gxy[1], sn_add2(make_float2(a1.x, a1.y), make_float2(b1.x, b1.y)));
            A.gz = sn_add2(A.gz, make_float2(a0.z + b0.z, a1.z + b1.z));
        } else {
            A.gxy[0] = sn_add2(A.gxy[0], axy[0]); A.gxy[1] = sn_add2(A.gxy[1], axy[1]); A.gz = sn_add2(A.gz, az);
        }
    }
    }
}

namespace snt {
// the 14 neighbour columns of the upper half plane (dx > 0, or dx == 0 and dy > 0); each stands
// for the pair {(dx,dy), (-dx,-dy)}
struct Col { int dx, dy; };
__host__ __device__ constexpr Col col(int idx)
{
    int n = 0;
    for (int dx = 0; dx <= 3; dx++)
        for (int dy = -3; dy <= 3; dy++) {
            if (dx * dx + dy * dy > 9 || !(dx > 0 || dy > 0)) continue;
            if (n == idx) return Col{dx, dy};
            n++;
        }
    return Col{0, 0};
}
}  // namespace snt

// The thread's view of the shared tile: pointers to the plane pair that holds its own two sites, in the (x,y)
// box (float4 = x,y of two planes), the z box and the length box (float2 = two planes).
struct SnTileCol {
    const float4 *xy;
    const float2 *z, *l;
};

// Planes e = -M .. M+1 (relative to the thread's first site) of the column C cells away, as (x, y, z, length):
// they lie in the plane pairs t = -(M+1)/2 .. (M+1)/2 -- one LDS.128 + one LDS.64 (+ one for the lengths) each.
template <int C, int M, bool SPECIES, int N>
__device__ __forceinline__ void sn_tile_load_col(const SnTileCol &tc, float4 (&w)[N])
{
    constexpr int LO = (M + 1) / 2;
    sn_static_for<-LO, LO + 1>([&](auto tcn) {
        constexpr int t = decltype(tcn)::value;
#ifdef SN_EXP_NOLOAD      // experiment: no shared-memory traffic, arithmetic only
        const float4 q = make_float4(C * 0.001f, t * 0.01f, 0.5f, 0.25f);
        const float2 zz = make_float2(C * 0.002f, t * 0.02f), ll = make_float2(1.0f, 1.0f);
#else
        const float4 q = tc.xy[C + t];
        const float2 zz = tc.z[C + t];
        float2 ll = make_float2(1.0f, 1.0f);
        if constexpr (SPECIES) ll = tc.l[C + t];
#endif
        if constexpr (2 * t >= -M && 2 * t <= M + 1) w[2 * t + M] = make_float4(q.x, q.y, zz.x, ll.x);
        if constexpr (2 * t + 1 >= -M && 2 * t + 1 <= M + 1) w[2 * t + 1 + M] = make_float4(q.z, q.w, zz.y, ll.y);
    });
}

template <int IDX, bool SPECIES>
__device__ __forceinline__ void sn_tile_load_pair(const SnTileCol &tc, float4 (&wp)[6], float4 (&wm)[6])
{
    constexpr snt::Col c = snt::col(IDX);
    constexpr int M = snt::half_height(c.dx * c.dx + c.dy * c.dy);
    constexpr int C = (c.dx * snt::BX + c.dy) * snt::NP;
    sn_tile_load_col<C, M, SPECIES>(tc, wp);
    sn_tile_load_col<-C, M, SPECIES>(tc, wm);
}

template <int IDX, bool SPECIES, int CUT>
__device__ __forceinline__ void sn_tile_compute_pair(const float4 (&wp)[6], const float4 (&wm)[6], SnAcc2 &A)
{
    constexpr snt::Col c = snt::col(IDX);
    constexpr int M = snt::half_height(c.dx * c.dx + c.dy * c.dy);
    // site S sees plane e of a column at window index e + M; w[M + 2t], w[M + 2t + 1] come from one plane pair
    sn_static_for<-M, M + 1>([&](auto dzc) {
        constexpr int DZ = decltype(dzc)::value;
#ifdef SN_EXP_NOFP        // experiment: loads only, one add per loaded word
        sn_static_for<0, 2>([&](auto sc) {
            constexpr int S = decltype(sc)::value;
            if constexpr (DZ == 0 || (S == 0 && DZ < 0) || (S == 1 && DZ > 0)) { A.d[S].x += wp[S + DZ + M].x; A.d[S].y += wm[S - DZ + M].y; }
        });
#else
        sn_accumulate_pair2<c.dx, c.dy, DZ, SPECIES, (DZ % 2 == 0), CUT>(A, wp[DZ + M], wm[M - DZ], wp[1 + DZ + M], wm[1 - DZ + M]);
#endif
    });
}

namespace snt {
// next column-pair index >= idx that is in MASK (NCOL if none)
__host__ __device__ constexpr int next_in(unsigned mask, int idx)
{
    while (idx < NCOL && !((mask >> idx) & 1u)) idx++;
    return idx;
}
}  // namespace snt

template <unsigned MASK, int IDX, bool SPECIES, int CUT>
__device__ __forceinline__ void sn_tile_gather_chain(const SnTileCol &tc, SnAcc2 &A, float4 (&c0)[6], float4 (&c1)[6])
{
    // c0/c1 hold pair IDX (already loaded); load the next pair of the mask, then do the arithmetic of this one
    if constexpr (IDX < snt::NCOL) {
        constexpr int NEXT = snt::next_in(MASK, IDX + 1);
        float4 n0[6], n1[6];
        if constexpr (NEXT < snt::NCOL) sn_tile_load_pair<NEXT, SPECIES>(tc, n0, n1);
        sn_tile_compute_pair<IDX, SPECIES, CUT>(c0, c1, A);
        if constexpr (NEXT < snt::NCOL) sn_tile_gather_chain<MASK, NEXT, SPECIES, CUT>(tc, A, n0, n1);
    }
}

// Same, with the loads running two column pairs ahead of the arithmetic (SN_EXP_DEPTH2)
template <unsigned MASK, int IDX, bool SPECIES, int CUT>
__device__ __forceinline__ void sn_tile_gather_chain2(const SnTileCol &tc, SnAcc2 &A, float4 (&c0)[6], float4 (&c1)[6], float4 (&n0)[6], float4 (&n1)[6])
{
    if constexpr (IDX < snt::NCOL) {
        constexpr int N1 = snt::next_in(MASK, IDX + 1);
        constexpr int N2 = N1 < snt::NCOL ? snt::next_in(MASK, N1 + 1) : snt::NCOL;
        float4 m0[6], m1[6];
        if constexpr (N2 < snt::NCOL) sn_tile_load_pair<N2, SPECIES>(tc, m0, m1);
        sn_tile_compute_pair<IDX, SPECIES, CUT>(c0, c1, A);
        if constexpr (N1 < snt::NCOL) sn_tile_gather_chain2<MASK, N1, SPECIES, CUT>(tc, A, n0, n1, m0, m1);
    }
}

// Contribution of the column pairs in MASK -- and of the thread's own column when CENTRE -- to
// the local fields of the thread's 2 consecutive z sites.  tc points at the plane pair of the thread's
// own two sites; a neighbour column is a compile-time immediate away.  Each
// column pair (+c, -c) is loaded once for both sites (sliding z window) and combined with the
// pair symmetry.  The loads of the next pair are issued before the arithmetic of the current one.
template <unsigned MASK, bool CENTRE, bool SPECIES, int CUT>
__device__ __forceinline__ void sn_tile_gather2(const SnTileCol &tc, float3 (&F)[2], float3 (&G)[2], float4 (&old)[2])
{
    SnAcc2 A;
#pragma unroll
    for (int s = 0; s < 2; s++) A.d[s] = A.o[s] = A.r[s] = A.gxy[s] = make_float2(0.f, 0.f);
    A.xz = A.yz = A.zz = A.gz = make_float2(0.f, 0.f);
    constexpr int FIRST = snt::next_in(MASK, 0);
    if constexpr (FIRST < snt::NCOL) {
        float4 c0[6], c1[6];
        sn_tile_load_pair<FIRST, SPECIES>(tc, c0, c1);
#ifdef SN_EXP_DEPTH2
        constexpr int SECOND = snt::next_in(MASK, FIRST + 1);
        float4 n0[6], n1[6];
        if constexpr (SECOND < snt::NCOL) sn_tile_load_pair<SECOND, SPECIES>(tc, n0, n1);
        sn_tile_gather_chain2<MASK, FIRST, SPECIES, CUT>(tc, A, c0, c1, n0, n1);
#else
        sn_tile_gather_chain<MASK, FIRST, SPECIES, CUT>(tc, A, c0, c1);
#endif
    }
    if constexpr (CENTRE) {
        // own column: pairs (0,0,+-dz); also yields the current values of the 2 sites
        float4 w[8];
        sn_tile_load_col<0, 3, SPECIES>(tc, w);
        const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4], w5 = w[5], w6 = w[6], w7 = w[7];
        old[0] = w3; old[1] = w4;
        sn_accumulate_pair2<0, 0, 1, SPECIES, false, CUT>(A, w4, w2, w5, w3);
        sn_accumulate_pair2<0, 0, 2, SPECIES, true, CUT>(A, w5, w1, w6, w2);
        sn_accumulate_pair2<0, 0, 3, SPECIES, false, CUT>(A, w6, w0, w7, w1);
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const float xz = s ? A.xz.y : A.xz.x, yz = s ? A.yz.y : A.yz.x, zz = s ? A.zz.y : A.zz.x;
        F[s] = make_float3(A.d[s].x + A.o[s].y + xz, A.o[s].x + A.d[s].y + yz, A.r[s].x + A.r[s].y + zz);
        G[s] = make_float3(A.gxy[s].x, A.gxy[s].y, s ? A.gz.y : A.gz.x);
    }
}

// Store two consecutive planes (z even, z + 1) of one column of the split copy (layout sn_pidx2: xy and z arrays;
// the lengths never change) and, if they lie within the ghost width of a face, their images: the periodic copies
// in x / y (/ z when the handle owns the whole axis) and the neighbouring GPUs' ghost planes over NVLink.  Extents
// are >= 32 here, so a site has at most one image shift per axis: at most 7 images, written with a handful of
// predicated stores.  The pair granularity also touches the planes 3 and nz - 4, whose "images" land in the
// padding planes nz + 3 and -4 that nobody reads.
__device__ __forceinline__ void sn_store_pair2(float4 *__restrict__ lat2, float4 *__restrict__ peer_lo, float4 *__restrict__ peer_hi,
                                               const SnGeom &G, const long long n2, int x, int y, int z, const float4 xy, const float2 zz)
{
    auto put = [&](float4 *__restrict__ base, int xx, int yy, int zc) {
        const long long c = sn_pidx2(G, xx, yy, zc) >> 1;
        base[c] = xy;
        reinterpret_cast<float2 *>(reinterpret_cast<float *>(base) + 2 * n2)[c] = zz;
    };
    put(lat2, x, y, z);
    const int g = G.g, gz = G.gz;
    const int ix = x < g ? G.X : (x >= G.X - g ? -G.X : 0);
    const int iy = y < g ? G.Y : (y >= G.Y - g ? -G.Y : 0);
    const int iz = z < gz ? G.nz : (z + 1 >= G.nz - gz ? -G.nz : 0);
    if ((ix | iy | iz) == 0) return;
    if (ix) put(lat2, x + ix, y, z);
    if (iy) put(lat2, x, y + iy, z);
    if (ix && iy) put(lat2, x + ix, y + iy, z);
    if (iz) {
        // periodic z: images in this array; Z-slab: the same planes live in the neighbour's ghost shell
#ifdef SN_EXP_NOREMOTE
        float4 *__restrict__ dst = G.periodic_z ? lat2 : nullptr;
#else
        float4 *__restrict__ dst = G.periodic_z ? lat2 : (iz > 0 ? peer_lo : peer_hi);
#endif
        if (dst) {
            put(dst, x, y, z + iz);
            if (ix) put(dst, x + ix, y, z + iz);
            if (iy) put(dst, x, y + iy, z + iz);
            if (ix && iy) put(dst, x + ix, y + iy, z + iz);
        }
    }
}

// canonical padded array <-> split copy, every padded cell (ghosts included)
__global__ void sn_convert_layout_kernel(float4 *__restrict__ lat, float4 *__restrict__ lat2, const SnGeom G, const int to_tiled)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.rep_stride) return;
    const int zp = (int)(i % G.PZ), yp = (int)((i / G.PZ) % G.PY), xp = (int)(i / ((long long)G.PZ * G.PY));
    float4 *a = lat + (long long)blockIdx.y * G.rep_stride + i;
    float4 *b = lat2 + (long long)blockIdx.y * sn_rep_stride2(G);
    const long long c = sn_pidx2(G, xp - G.g, yp - G.g, zp - G.gz);
    if (to_tiled) sn_st2(b, G, c, *a); else *a = sn_ld2(b, G, c);
}

// Tile colours along one axis of n tiles: tiles that are active together must not be neighbours, also across the
// periodic wrap.  Even n: two colours (parity).  Odd n (>= 3): the last tile gets a colour of its own, the others
// alternate -- 3 colours.  A sweep visits the ncx * ncy * ncz colour combinations ("phases", 8 to 27) in order.
__host__ __device__ inline int sn_tc_ncol(int n) { return (n & 1) ? 3 : 2; }
__host__ __device__ inline int sn_tc_count(int n, int c) { return c < 2 ? (n >> 1) : 1; }
__host__ __device__ inline int sn_tc_tile(int n, int c, int i) { return c < 2 ? 2 * i + c : n - 1; }
__host__ __device__ inline int sn_tc_colour(int n, int t) { return ((n & 1) && t == n - 1) ? 2 : (t & 1); }

struct SnTileFlow {
    int tnx, tny, tnz;                  // tiles per axis (own tiles: z ghost layers not counted)
    int ncx, ncy, ncz, np;              // colours per axis, phases per sweep
    int periodic_z;                     // the handle owns the whole z axis (ghost version layers mirror its own end tiles)
    int nrep;
    unsigned int pre[28];               // pre[p] = items of the phases before p within one sweep; pre[np] = items per sweep
    unsigned long long base_sweep;      // sweeps completed before item 0 (the version of every tile at that point)
    unsigned long long n_begin, n_end;  // items of this launch; item n = sweep * pre[np] + pre[phase] + replica * tiles(phase) + tile
    unsigned long long *next;           // work counter (device, zero at launch): item = n_begin + atomicAdd(next, 1)
    unsigned int *ver;                  // [rep][tnx][tny][tnz + 2] tile versions; z index shifted by one ghost layer
    unsigned int *peer_ver_lo, *peer_ver_hi;   // where my bottom / top tile layer is a ghost layer: the slab neighbours' arrays, or `ver` itself
    int sys_scope;                      // ghost versions and planes are written by other GPUs
    float *audit;                       // AUDIT instantiation: one record per attempt (sn_audit_write), else unused
    unsigned int *err;                  // SN_FLAGS_ERR of this handle: raised when a dependency wait runs out of time
    unsigned long long timeout_ns;      // bound of one dependency wait
};

struct SnTileMaps { CUtensorMap xy, z, l; };     // tensor maps of the three arrays of the split copy

struct __align__(16) SnTileItem {
    int x0, y0, z0, rep;                // tile origin (slab-local z) and replica
    int tx, ty, tz, p;                  // tile coordinates and phase ((cx * ncy + cy) * ncz + cz)
    unsigned long long sweep;           // global sweep index = Philox counter word = tile version before the item
    int valid, ready;                   // (scheduler hand-over) item exists; its dependencies were met when polled
};

__host__ __device__ __forceinline__ SnTileItem sn_tile_item(const SnTileFlow &f, unsigned long long n)
{
    SnTileItem it;
    it.valid = 1; it.ready = 0;
    const unsigned long long S = f.pre[f.np];
    const unsigned long long sw = n / S;
    const unsigned int m = (unsigned int)(n - sw * S);
    int p = 0;
    while (p + 1 < f.np && m >= f.pre[p + 1]) p++;
    it.sweep = f.base_sweep + sw;
    it.p = p;
    const int cz = p % f.ncz, cy = (p / f.ncz) % f.ncy, cx = p / (f.ncz * f.ncy);
    const int kx = sn_tc_count(f.tnx, cx), ky = sn_tc_count(f.tny, cy), kz = sn_tc_count(f.tnz, cz);
    const unsigned int P1 = (unsigned int)(kx * ky * kz), pos = m - f.pre[p];
    it.rep = (int)(pos / P1);
    const int r = (int)(pos % P1), iz = r % kz, iy = (r / kz) % ky;
    // The x order is rotated by one tile plane per sweep: the wrap-around neighbour (ix - 1 of ix = 0) would
    // otherwise be the last plane of the previous sweep, i.e. a barrier per sweep; rotated, every dependency
    // of an item lies at least ~a phase back in the order.
    const int ix = (r / (kz * ky) + (int)(it.sweep % (unsigned)kx)) % kx;
    it.tx = sn_tc_tile(f.tnx, cx, ix); it.ty = sn_tc_tile(f.tny, cy, iy); it.tz = sn_tc_tile(f.tnz, cz, iz);
    it.x0 = it.tx * snt::T; it.y0 = it.ty * snt::T; it.z0 = it.tz * snt::T;
    return it;
}

__device__ __forceinline__ long long sn_ver_index(const SnTileFlow &f, int rep, int tx, int ty, int tzg)
{
    return (((long long)rep * f.tnx + tx) * f.tny + ty) * (f.tnz + 2) + tzg;
}

// Warp-collective: have the 26 neighbouring tiles of `it` reached the version that precedes it?
__device__ __forceinline__ bool sn_tile_deps_ready(const SnTileFlow &f, const SnTileItem &it, int lane)
{
    bool ok = true;
    if (lane < 27 && lane != 13) {
        const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
        int ntx = it.tx + dx, nty = it.ty + dy;
        const int ntz = it.tz + dz;                                   // -1 .. tnz: the ends are ghost layers
        ntx = ntx < 0 ? ntx + f.tnx : (ntx >= f.tnx ? ntx - f.tnx : ntx);
        nty = nty < 0 ? nty + f.tny : (nty >= f.tny ? nty - f.tny : nty);
        // colour of the z neighbour: a ghost layer stands for the handle's own end tile (periodic) or for the slab
        // neighbour's boundary layer (slabs have an even number of tile layers, globally and each: parity)
        const int wz = ntz < 0 ? ntz + f.tnz : (ntz >= f.tnz ? ntz - f.tnz : ntz);
        const int qz = f.periodic_z ? sn_tc_colour(f.tnz, wz) : (ntz & 1);
        const int q = (sn_tc_colour(f.tnx, ntx) * f.ncy + sn_tc_colour(f.tny, nty)) * f.ncz + qz;
        const unsigned int need = (unsigned int)it.sweep + (q < it.p ? 1u : 0u);
        const unsigned int *src = f.ver + sn_ver_index(f, it.rep, ntx, nty, ntz + 1);
        unsigned int v;
#ifdef SN_EXP_GPUACQ
        const bool remote_entry = false;
#else
        const bool remote_entry = f.sys_scope && (ntz < 0 || ntz >= f.tnz);
#endif
        if (remote_entry) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        ok = (int)(v - need) >= 0;
    }
    return __all_sync(0xffffffffu, ok);
}

// One thread, after every store of the tile (write-back, ghost images, pushes to the neighbours) has been fenced at CTA
// scope by the workers and observed through the arrival counter.  ONE fence at the scope the readers need, then relaxed
// stores of the version words (fence + relaxed store is a release pattern; a st.release per word would cost a MEMBAR
// each -- with pushes to a neighbour GPU in flight every system-scope MEMBAR waits for an NVLink round trip, which
// made 64-plane slabs 13 % slower than the same tiles on one GPU).
__device__ __forceinline__ void sn_tile_publish(const SnTileFlow &f, const SnTileItem &it)
{
    const unsigned int v = (unsigned int)it.sweep + 1u;
    unsigned int *own = f.ver + sn_ver_index(f, it.rep, it.tx, it.ty, it.tz + 1);
    unsigned int *glo = it.tz == 0 ? f.peer_ver_lo + sn_ver_index(f, it.rep, it.tx, it.ty, f.tnz + 1) : nullptr;
    unsigned int *ghi = it.tz == f.tnz - 1 ? f.peer_ver_hi + sn_ver_index(f, it.rep, it.tx, it.ty, 0) : nullptr;
#ifdef SN_EXP_GPUFENCE
    if (false) {
#else
    if (f.sys_scope && (glo || ghi)) {
#endif
        // a boundary tile: its pushes into the neighbour GPU's ghost planes become visible system-wide before the versions
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(own), "r"(v) : "memory");
        if (glo) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(glo), "r"(v) : "memory");
        if (ghi) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(ghi), "r"(v) : "memory");
    } else {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(own), "r"(v) : "memory");
        if (glo) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(glo), "r"(v) : "memory");
        if (ghi) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ghi), "r"(v) : "memory");
    }
}

// AUDIT: every attempt also leaves a record (proposal, accept uniform, the dE the chain used, decision, group
// ordinal) for sn_mc_sweep_audit; the arithmetic and the order are those of the product instantiation.
template <bool SPECIES, bool AUDIT, int CUT>
__global__ void __launch_bounds__(snt::THREADS, 1)
sn_tiled_kernel(const __grid_constant__ SnTileMaps maps, const SnSweepArgs a, const SnTileFlow fl)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *tile_xy = reinterpret_cast<float4 *>(smem);                     // [22 * 22 columns][14 plane pairs] (x0, y0, x1, y1)
    float2 *tile_z = reinterpret_cast<float2 *>(smem + snt::OFF_Z);         // ... (z0, z1)
    float2 *tile_l = reinterpret_cast<float2 *>(smem + snt::OFF_L);         // ... (l0, l1), SPECIES only
    const uint32_t bar = sn_smem_u32(smem + snt::OFF_BAR);
    // The hand-over slot alternates: the control warp may prepare the next item (slot n + 1) as soon as its dependencies are
    // met, while a slow worker warp is still reading the current one (slot n) behind the previous hand-over barrier.
    SnTileItem *ctl_item = reinterpret_cast<SnTileItem *>(smem + snt::OFF_CTL);
    unsigned int *ctl_wbread = reinterpret_cast<unsigned int *>(smem + snt::OFF_CTL + 96);   // worker warps that reached the write-back
    unsigned int *ctl_arrive = reinterpret_cast<unsigned int *>(smem + snt::OFF_CTL + 100);  // worker warps whose stores are fenced
    int slot = 0;

    // Two roles of 128 threads (4 warps) each -- one warp of each role per scheduler, so that one
    // role's issue slots fill the other's latency gaps:
    //   role A gathers column pairs [0, SPLIT) and the own column, then runs the sequential chain;
    //   role B gathers the remaining pairs, hands its partial fields over through shared memory and
    //          draws the Philox proposals of the next super-pass while A is in the chain.
    // A ninth warp is the control warp: it takes the next work item, polls its dependencies, issues the TMA
    // load as soon as the workers have let go of the shared tile, and publishes the finished tile's version
    // once their stores have landed -- all the global-memory round trips of the scheduling stay off the
    // workers' critical path.
    // thread -> (column i,j ; segment k of 4 z sites ; half h of the segment).  The plane pair of a thread's two
    // sites is 2 + 2k + h: k and h are the low lane bits, so a quarter-warp reads 8 consecutive 16-byte (x,y) pairs
    // and a half-warp (two columns 4 apart in y: 56 = 8 mod 16 pairs) 16 distinct 8-byte z pairs.
    const int tid = threadIdx.x, role = tid >> 7, tl = tid & 127, lane = tid & 31, i = tl >> 5;
    const int k = lane & 3, h = (lane >> 2) & 1, j = lane >> 3;
    const bool ctrl = tid >= snt::WORKERS;
    float2 *xF = reinterpret_cast<float2 *>(smem + snt::OFF_XF);
    float4 *xP = reinterpret_cast<float4 *>(smem + snt::OFF_XP);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *ctl_arrive = 0; *ctl_wbread = 0;
    }
    __syncthreads();

    uint32_t parity = 0;
    const SnGeom &G = a.G;

    const int pair0 = 2 + 2 * k + h;          // plane pair of the thread's two sites inside the 14 pairs of a column

    // TMA load of a tile into shared memory (one box per array), completion on the mbarrier (one thread)
    auto issue_tile_load = [&](const SnTileItem &it) {
        // shared memory was last touched through the generic proxy; order it before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        constexpr int BYTES = snt::XY_BYTES + snt::Z_BYTES + (SPECIES ? snt::Z_BYTES : 0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BYTES) : "memory");
        // tensor = (floats of one z row, padded y, padded x, replica); the 28-plane window of a tile at z0 starts at
        // row position z0 (plane z0 - 4); halo x0-3 / y0-3 -> padded x0 / y0 for cut-off 3; with a smaller cut-off the ghost shell
        // is narrower than the halo and the box starts outside the array (zero fill; those cells carry zero tensors anyway)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(sn_smem_u32(smem)), "l"(&maps.xy), "r"(2 * it.z0), "r"(it.y0 + G.g - snt::H), "r"(it.x0 + G.g - snt::H), "r"(it.rep), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(sn_smem_u32(smem + snt::OFF_Z)), "l"(&maps.z), "r"(it.z0), "r"(it.y0 + G.g - snt::H), "r"(it.x0 + G.g - snt::H), "r"(it.rep), "r"(bar) : "memory");
        if constexpr (SPECIES)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(sn_smem_u32(smem + snt::OFF_L)), "l"(&maps.l), "r"(it.z0), "r"(it.y0 + G.g - snt::H), "r"(it.x0 + G.g - snt::H), "r"(it.rep), "r"(bar) : "memory");
    };
    // Block-wide rendezvous of the control warp and the workers.  They meet from different places in the code, so this
    // is a named barrier with an explicit thread count (bar.sync 2, THREADS), not __syncthreads().
    auto cta_sync = [&]() { asm volatile("barrier.sync 2, %0;" ::"n"(snt::THREADS) : "memory"); };
    // ---- control warp -----------------------------------------------------------------------------
    auto deps_met = [&]() {
        // The halo was written by other CTAs / GPUs through the generic proxy and observed by this warp's
        // acquire loads; the TMA reads it through the async proxy.
        asm volatile("fence.proxy.async.global;" ::: "memory");
    };
    auto take = [&]() -> SnTileItem {
        unsigned long long n = lane == 0 ? fl.n_begin + atomicAdd(fl.next, 1ULL) : 0ULL;
        n = __shfl_sync(0xffffffffu, n, 0);
        const unsigned int failed = *reinterpret_cast<volatile unsigned int *>(fl.err);     // a wait timed out somewhere: drain
        SnTileItem it;
        it.valid = 0; it.ready = 0;
        if (n < fl.n_end && !failed) it = sn_tile_item(fl, n);
        return it;
    };
    // Blocking wait for an item's dependencies, bounded: if a neighbour (another CTA, or a slab neighbour's kernel on
    // another GPU that was never launched) does not get there in time, raise the handle's error flag and go on -- the
    // tile's result is meaningless, every control warp stops taking items, and the host call that synchronises fails.
    auto wait_deps = [&](const SnTileItem &it) {
        const unsigned long long t0 = sn_globaltimer_ns();
        while (!sn_tile_deps_ready(fl, it, lane)) {
            __nanosleep(100);
            if (sn_globaltimer_ns() - t0 > fl.timeout_ns || *reinterpret_cast<volatile unsigned int *>(fl.err)) {
                if (lane == 0) atomicExch(fl.err, 2u);
                break;
            }
        }
    };
    if (ctrl) {
        SnTileItem cur = take();
        if (cur.valid) {
            wait_deps(cur);
            deps_met();
            if (lane == 0) issue_tile_load(cur);
        }
        if (lane == 0) ctl_item[0] = cur;
        __syncwarp();
        cta_sync();
        for (unsigned int done = 8; cur.valid; done += 8) {
            // while the workers compute `cur`: next item, first look at its dependencies
            SnTileItem nxt = take();
            // Poll only while the next item's dependencies are unmet and no worker warp has reached the write-back, with a
            // pause between looks: the control warp shares its scheduler with two worker warps, and every instruction it
            // issues is a slot they do not get (the v11 profile had 19 % of all executed instructions in this loop).  Once
            // the dependencies are met it blocks in the hand-over barrier, which costs nothing.
            bool ready = !nxt.valid;
            while (!ready && *reinterpret_cast<volatile unsigned int *>(ctl_wbread) + 8u - done == 0u) {
                ready = sn_tile_deps_ready(fl, nxt, lane);
                if (!ready) __nanosleep(SN_CTRL_POLL_NS);
            }
            ready = ready && nxt.valid;
            if (ready) deps_met();
            nxt.ready = ready;
            slot ^= 1;
            if (lane == 0) ctl_item[slot] = nxt;
            __syncwarp();
            cta_sync();                                   // hand-over: every worker has read its part of the tile
            if (ready && lane == 0) issue_tile_load(nxt);
            while ((int)(*reinterpret_cast<volatile unsigned int *>(ctl_arrive) - done) < 0) __nanosleep(40);
            __syncwarp();
            // one more look at the next item before the (possibly slow) publication: if its dependencies have arrived in
            // the meantime, its load goes first (it cannot depend on `cur` then, whose version is still the old one)
            if (nxt.valid && !ready && sn_tile_deps_ready(fl, nxt, lane)) {
                ready = true;
                deps_met();
                if (lane == 0) issue_tile_load(nxt);
            }
            __syncwarp();
            if (lane == 0) sn_tile_publish(fl, cur);
            if (nxt.valid && !ready) {
                wait_deps(nxt);
                deps_met();
                if (lane == 0) issue_tile_load(nxt);
            }
            cur = nxt;
        }
        return;
    }

    // ---- workers -----------------------------------------------------------------------------------
    auto workers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(snt::WORKERS) : "memory"); };
    cta_sync();
    SnTileItem item = ctl_item[0];

    while (item.valid) {
        const int x0 = item.x0, y0 = item.y0, z0 = item.z0, rep = item.rep;
        const uint32_t sweep_lo = (uint32_t)item.sweep, sweep_hi = (uint32_t)(item.sweep >> 32);

        SnTerms tm;
        tm.K = a.K; tm.beta = a.beta[rep];
        { const float4 E = a.efield[rep]; tm.E = make_float3(E.x, E.y, E.z); tm.cage = E.w; }
        tm.constrain = a.constrain; tm.dim = a.dim;
        const uint4 rkey = a.rep_key[rep];
        // a.lat / a.peer_* are the split copies here (layout sn_pidx2)
        const long long rs2 = sn_rep_stride2(G);
        float4 *glat = a.lat + (long long)rep * rs2;
        float4 *plo = a.peer_lo ? a.peer_lo + (long long)rep * rs2 : nullptr;
        float4 *phi = a.peer_hi ? a.peer_hi + (long long)rep * rs2 : nullptr;
        // (a partial tile -- extents need not be multiples of 16 -- ends at the lattice face: its cells beyond the face are
        // zero-filled by the TMA unit, never attempted and never written back)
        const bool face_tile = x0 == 0 || x0 + snt::T >= G.X || y0 == 0 || y0 + snt::T >= G.Y || z0 == 0 || z0 + snt::T >= G.nz;

        // Trial orientations for the thread's 2 sites in super-pass sp from ONE Philox4x32-10 call keyed by
        // (global site of the first one, replica, sweep): per site 64 random bits = 32 (accept) + 20 + 20, the low 8
        // bits of the accept word doubling as the last bits of the azimuth (they only reach the accept uniform
        // when it is below 2^-8, as resolution beyond 2^-24).
        auto draw = [&](int sp) __attribute__((always_inline)) {
            const int gx = x0 + (sp >> 2) + 4 * i, gy = y0 + (sp & 3) + 4 * j, gz = z0 + 4 * k + 2 * h;
            const unsigned long long gsite = ((unsigned long long)gx * G.Y + gy) * G.Z + (G.z0 + gz);
            const Philox4 r = sn_philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32) ^ rkey.z, sweep_lo, sweep_hi, rkey.x, rkey.y);
            const uint32_t wa[2] = {r.x, r.z}, wb[2] = {r.y, r.w};
            float4 *dst = xP + (sp & 1) * 2 * snt::SITE_THREADS + tl;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const float u = (float)(wb[s] >> 12) * (1.0f / 1048576.0f);
                const float v = (float)(((wb[s] & 0xfffu) << 8) | (wa[s] & 0xffu)) * (1.0f / 1048576.0f);
                const float3 np = sn_propose(tm, u, v);
                dst[s * snt::SITE_THREADS] = make_float4(np.x, np.y, np.z, sn_u01_32(wa[s]));
            }
        };
        if (role == 1) draw(0);                          // overlaps the TMA flight

        sn_mbar_wait(bar, parity);
        parity ^= 1;

        // role B's share of the fields of super-pass sp, left in xF[sp & 1]
        auto gather_b = [&](int sp) __attribute__((always_inline)) {
            const int cell = ((snt::H + (sp >> 2) + 4 * i) * snt::BX + (snt::H + (sp & 3) + 4 * j)) * snt::NP + pair0;
            const SnTileCol tc{tile_xy + cell, tile_z + cell, tile_l + cell};
            float3 F[2], Gc[2];
            float4 old[2];
#ifdef SN_EXP_NOGB
            sn_tile_gather2<0u, false, SPECIES, CUT>(tc, F, Gc, old);
#else
            sn_tile_gather2<snt::MASK_B, false, SPECIES, CUT>(tc, F, Gc, old);
#endif
            float2 *dst = xF + (sp & 1) * 3 * snt::SITE_THREADS + tl;
            dst[0 * snt::SITE_THREADS] = make_float2(F[0].x, F[0].y);
            dst[1 * snt::SITE_THREADS] = make_float2(F[0].z, F[1].x);
            dst[2 * snt::SITE_THREADS] = make_float2(F[1].y, F[1].z);
        };
        if (role == 1) gather_b(0);
        workers_sync();

        int n_acc = 0, n_rej = 0, n_vac = 0;
#pragma unroll 1
        for (int sp = 0; sp < 16; sp++) {
            if (role == 0) {
                const int cx = sp >> 2, cy = sp & 3;
                const int gx = x0 + cx + 4 * i, gy = y0 + cy + 4 * j, gz = z0 + 4 * k + 2 * h;
                const bool live = gx < G.X && gy < G.Y && gz < G.nz;         // false only in the cells of a partial tile beyond the lattice
                const int cell = ((snt::H + cx + 4 * i) * snt::BX + (snt::H + cy + 4 * j)) * snt::NP + pair0;
                const SnTileCol tc{tile_xy + cell, tile_z + cell, tile_l + cell};
                float3 F[2], Gc[2];
                float4 old[2];
#ifdef SN_EXP_NOGA
                sn_tile_gather2<0u, true, SPECIES, CUT>(tc, F, Gc, old);
#else
                sn_tile_gather2<snt::MASK_A, true, SPECIES, CUT>(tc, F, Gc, old);
#endif
                {
                    const float2 *src = xF + (sp & 1) * 3 * snt::SITE_THREADS + tl;
                    const float2 v0 = src[0 * snt::SITE_THREADS], v1 = src[1 * snt::SITE_THREADS], v2 = src[2 * snt::SITE_THREADS];
                    F[0].x += v0.x; F[0].y += v0.y; F[0].z += v1.x; F[1].x += v1.y; F[1].y += v2.x; F[1].z += v2.y;
                }
                float3 np[2]; float ua[2];
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const float4 pr = xP[((sp & 1) * 2 + s) * snt::SITE_THREADS + tl];
                    np[s] = make_float3(pr.x, pr.y, pr.z); ua[s] = pr.w;
                }
                // The 4 sites of a segment interact: decide them in z order.  Step t belongs to the lane with
                // h == t/2 (local site t&1); its accepted change dp is broadcast to the partner lane and to the
                // segment below (which sees it as "the segment above"), and enters the dE of the later sites as
                //   l_i dp_i . T(0,0,dz) (l_j dp_j),  T(0,0,dz) = diag(1,1,-2)/|dz|^3,   plus the cage term for |dz| = 1.
                // Everything that does not depend on earlier decisions (dE0, q, cg) is computed up front, so the
                // sequential part is one short dot product, the accept test and six independent shuffles per step.
                float3 dpp[2], q[2], cg[2];
                float dE0[2];
                bool vac[2];
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const float4 o = old[s];
                    vac[s] = o.w == 0.0f;                                                      // montecarlo-core.c:163
                    dpp[s] = make_float3(np[s].x - o.x, np[s].y - o.y, np[s].z - o.z);
                    dE0[s] = sn_delta_e(o, np[s], F[s], Gc[s], tm);
                    q[s] = make_float3(o.w * dpp[s].x, o.w * dpp[s].y, -2.0f * o.w * dpp[s].z);
                    cg[s] = make_float3(-tm.cage * dpp[s].x, -tm.cage * dpp[s].y, -tm.cage * dpp[s].z);
                }
                float3 dmS[4], dpS[4], dmU[4], dpU[4];
                __syncwarp();      // every lane has read its own column (which holds its neighbours' segments) before any lane's chain writes
#ifdef SN_EXP_NOCHAIN
#pragma unroll
                for (int t4 = 0; t4 < 0; t4++) {
#else
#pragma unroll
                for (int t4 = 0; t4 < 4; t4++) {
#endif
                    const int s = t4 & 1;
                    float dE = dE0[s];
#pragma unroll
                    for (int t2 = 0; t2 < t4; t2++) {
                        const int d1 = t4 - t2, d2 = 4 + t2 - t4;      // distance to the earlier site below / in the segment above
                        const float w1 = d1 > CUT ? 0.0f : d1 == 1 ? 1.0f : d1 == 2 ? 0.125f : (1.0f / 27.0f);     // beyond the cut-off: no interaction
                        const float w2 = d2 > CUT ? 0.0f : d2 == 1 ? 1.0f : d2 == 2 ? 0.125f : (1.0f / 27.0f);
                        dE = fmaf(w1, q[s].x * dmS[t2].x + q[s].y * dmS[t2].y + q[s].z * dmS[t2].z, dE);
                        dE = fmaf(w2, q[s].x * dmU[t2].x + q[s].y * dmU[t2].y + q[s].z * dmU[t2].z, dE);
                        if (d1 == 1) dE += cg[s].x * dpS[t2].x + cg[s].y * dpS[t2].y + cg[s].z * dpS[t2].z;
                        if (d2 == 1) dE += cg[s].x * dpU[t2].x + cg[s].y * dpU[t2].y + cg[s].z * dpU[t2].z;
                    }
                    const bool mine = (h == (t4 >> 1)) & live;
#ifdef SN_EXP_NOEXP
                    const bool acc = mine & !vac[s] & (dE < 0.0f);
#else
                    const bool acc = mine & !vac[s] & sn_accept(dE, tm.beta, ua[s]);           // montecarlo-core.c:179 (no short-circuit: no divergence)
#endif
                    const float3 dp = acc ? dpp[s] : make_float3(0.f, 0.f, 0.f);
#ifndef SN_EXP_NOSTS
                    if (acc) {
                        reinterpret_cast<float2 *>(tile_xy)[2 * cell + s] = make_float2(np[s].x, np[s].y);
                        reinterpret_cast<float *>(tile_z)[2 * cell + s] = np[s].z;
                    }
#endif
                    n_acc += acc; n_rej += (mine & !acc & !vac[s]); n_vac += (mine & vac[s]);
                    if constexpr (AUDIT) {
                        if (mine) sn_audit_write(fl.audit, G, rep, gx, gy, gz + s, np[s], ua[s], dE, acc, vac[s], (item.p * 16 + sp) * 4 + t4);
                    }
                    if (t4 < 3) {
                        const int srcS = (lane & 27) | ((t4 >> 1) << 2);            // owner of step t4 in this segment
                        const int srcU = (((lane & 27) + 1) & 27) | ((t4 >> 1) << 2);   // ... in the segment above (k + 1; unused for k == 3)
#ifdef SN_EXP_NOSHFL
                        dpS[t4] = dp; dpU[t4] = make_float3(dp.y, dp.z, dp.x); (void)srcS; (void)srcU;
#else
                        dpS[t4].x = __shfl_sync(0xffffffffu, dp.x, srcS);
                        dpS[t4].y = __shfl_sync(0xffffffffu, dp.y, srcS);
                        dpS[t4].z = __shfl_sync(0xffffffffu, dp.z, srcS);
                        dpU[t4].x = __shfl_sync(0xffffffffu, dp.x, srcU);
                        dpU[t4].y = __shfl_sync(0xffffffffu, dp.y, srcU);
                        dpU[t4].z = __shfl_sync(0xffffffffu, dp.z, srcU);
#endif
                        if constexpr (SPECIES) {
                            const float lS = __shfl_sync(0xffffffffu, old[s].w, srcS), lU = __shfl_sync(0xffffffffu, old[s].w, srcU);
                            dmS[t4] = make_float3(lS * dpS[t4].x, lS * dpS[t4].y, lS * dpS[t4].z);
                            dmU[t4] = make_float3(lU * dpU[t4].x, lU * dpU[t4].y, lU * dpU[t4].z);
                        } else { dmS[t4] = dpS[t4]; dmU[t4] = dpU[t4]; }
                        if (k == 3) { dpU[t4] = make_float3(0.f, 0.f, 0.f); dmU[t4] = make_float3(0.f, 0.f, 0.f); }   // above lies the static halo
                    }
                }
            } else if (sp < 15) {
                // one super-pass ahead of the chain: none of role B's columns belongs to the class being updated
                // Role B draws first and gathers second, role A gathers first and runs the chain second: the two roles'
                // shared-memory phases alternate instead of colliding at the start of the super-pass (+13 % at 512^3).
#ifdef SN_EXP_GATHER_FIRST
                gather_b(sp + 1);
                draw(sp + 1);
#else
                draw(sp + 1);
                gather_b(sp + 1);
#endif
            }
            workers_sync();
        }
        // Write the tile's 16^3 interior back: shared memory -> registers, then (once everybody has read) the
        // TMA load of the next tile is started and the registers are stored to global memory underneath it.
        // 8 lanes cover one (x,y) row of 16 planes: 128 contiguous bytes of the xy array, 64 of the z array.  Rows
        // on a lattice / slab face also go to their ghost images (periodic copies, or the neighbouring GPU's ghost
        // planes over NVLink).
        SnTileItem nxt;
        {
            float4 wbx[8];
            float2 wbz[8];
            const int pr = tid & 7, row0 = tid >> 3;                       // plane pair z0 + 2 pr, 32 rows per pass
#pragma unroll
            for (int pss = 0; pss < 8; pss++) {
                const int row = pss * 32 + row0, lx = row >> 4, ly = row & 15;
                const int cell = ((lx + snt::H) * snt::BX + (ly + snt::H)) * snt::NP + 2 + pr;
                wbx[pss] = tile_xy[cell];
                wbz[pss] = tile_z[cell];
            }
            __syncwarp();
            if (lane == 0) atomicAdd(ctl_wbread, 1u);     // tells the control warp to stop polling and come to the hand-over
            __syncwarp();
            cta_sync();                                   // hand-over: the control warp starts the next TMA load
            slot ^= 1;
            nxt = ctl_item[slot];
            float4 *gxy = glat;                                                            // pairs of the xy array
            float2 *gz2 = reinterpret_cast<float2 *>(reinterpret_cast<float *>(glat) + 2 * rs2);   // pairs of the z array
            const long long gbase = sn_pidx2(G, x0, y0, z0 + 2 * pr) >> 1;
            const long long sy2 = sn_pz2(G) >> 1, sx2 = sy2 * G.PY;
#pragma unroll
            for (int pss = 0; pss < 8; pss++) {
                const int row = pss * 32 + row0, lx = row >> 4, ly = row & 15;
                if (face_tile) { if (x0 + lx < G.X && y0 + ly < G.Y && z0 + 2 * pr < G.nz) sn_store_pair2(glat, plo, phi, G, rs2, x0 + lx, y0 + ly, z0 + 2 * pr, wbx[pss], wbz[pss]); }
                else { gxy[gbase + lx * sx2 + ly * sy2] = wbx[pss]; gz2[gbase + lx * sx2 + ly * sy2] = wbz[pss]; }
            }
        }
        // Release the tile's stores to the control warp at CTA scope (the lanes' stores are ordered before lane
        // 0's fence by the warp barrier); the control warp sees the 8 arrivals, fences at GPU scope -- system
        // scope for a tile that pushed into a neighbour GPU -- and publishes the tile's new version.  By the
        // cumulativity of the PTX memory model that chain orders every store of the tile before the version,
        // without a GPU-scope fence (MEMBAR.SC.GPU + L1 invalidate, 6 % of the kernel) in every worker thread.
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(ctl_arrive, 1u); }
        if (role == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
                n_rej += __shfl_xor_sync(0xffffffffu, n_rej, o);
                n_vac += __shfl_xor_sync(0xffffffffu, n_vac, o);
            }
            if (lane == 0) {
                unsigned long long *c = a.counters + 3 * rep;
                if (n_acc) atomicAdd(c + 0, (unsigned long long)n_acc);
                if (n_rej) atomicAdd(c + 1, (unsigned long long)n_rej);
                if (n_vac) atomicAdd(c + 2, (unsigned long long)n_vac);
            }
        }
        item = nxt;
    }
}

// ---- host side -------------------------------------------------------------------
typedef CUresult (*SnEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool sn_tiled_supported(const sn_handle *h, std::string *why)
{
    const SnGeom &G = h->G;
    const char *msg = nullptr;
    if (h->p.cutoff != 3 && h->p.cutoff != 2) msg = "DipoleCutOff must be 2 or 3";
    else if (G.Z == 1) msg = "lattice is flat (Z == 1)";
    // two tiles per axis at least, and a tile's box must not hold images of the tile's own sites: the lower halo of the
    // first tile shows the sites X-3 .. X-1, which must lie beyond its 16 own columns (X >= 19)
    else if (G.X < 20 || G.Y < 20 || G.nz < 20) msg = "X, Y and Z must be at least 20";
    else if (G.nz % 4) msg = "Z must be a multiple of 4 (a thread's two sites and a segment of 4 must not straddle the lattice's end)";
    else if (!G.periodic_z && (G.nz % 32 || G.z0 % 32 || G.Z % 32)) msg = "Z-slabs: Z, slab height and slab origin must be multiples of 32";
    if (msg) { if (why) *why = msg; return false; }
    return true;
}

int sn_tiled_prepare(sn_handle *h)
{
    const SnGeom &G = h->G;
    SnEncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    SN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr));
    if (!enc || qr != cudaDriverEntryPointSuccess) return sn_fail(SN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    // the split copy (layout sn_pidx2) and the tensor maps of its three arrays:
    // dims = (floats of one z row, padded y, padded x, replica); one TMA row = 28 planes = 224 B (xy) / 112 B (z, lengths)
    const long long rs2 = sn_rep_stride2(G);
    const int PZ2 = sn_pz2(G);
    const size_t bytes2 = (size_t)rs2 * h->p.nreplicas * sizeof(float4);
    if (cudaMalloc(&h->lat2, bytes2) != cudaSuccess) { cudaGetLastError(); return sn_fail(SN_ERR_NOMEM, "sn_create: cannot allocate %.1f MB for the tiled copy", bytes2 / 1e6); }
    SN_CUDA_CHECK(cudaMemsetAsync(h->lat2, 0, bytes2, h->stream));
    h->lat2_valid = false;
    SnTileMaps *tm = new SnTileMaps;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int which = 0; which < 3; which++) {
        const int fpc = which == 0 ? 2 : 1;                           // floats per cell in this array
        float *base = reinterpret_cast<float *>(h->lat2) + (which == 0 ? 0 : which == 1 ? 2 * rs2 : 3 * rs2);
        const cuuint64_t gdim[4] = {(cuuint64_t)PZ2 * fpc, (cuuint64_t)G.PY, (cuuint64_t)(G.X + 2 * G.g), (cuuint64_t)h->p.nreplicas};
        const cuuint64_t gstr[3] = {(cuuint64_t)PZ2 * fpc * 4, (cuuint64_t)PZ2 * fpc * 4 * G.PY, (cuuint64_t)rs2 * 16};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * snt::NP * fpc), snt::BX, snt::BX, 1};
        CUtensorMap *m = which == 0 ? &tm->xy : which == 1 ? &tm->z : &tm->l;
        const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete tm; return sn_fail(SN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r); }
    }
    h->tmap = tm;
    for (const void *k : {(const void *)sn_tiled_kernel<true, false, 3>, (const void *)sn_tiled_kernel<false, false, 3>, (const void *)sn_tiled_kernel<true, true, 3>,
                          (const void *)sn_tiled_kernel<false, true, 3>, (const void *)sn_tiled_kernel<true, false, 2>, (const void *)sn_tiled_kernel<false, false, 2>,
                          (const void *)sn_tiled_kernel<true, true, 2>, (const void *)sn_tiled_kernel<false, true, 2>})
        SN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, snt::SMEM_BYTES));
    return SN_OK;
}

void sn_tiled_release(sn_handle *h)
{
    delete reinterpret_cast<SnTileMaps *>(h->tmap);
    h->tmap = nullptr;
    cudaFree(h->lat2);
    h->lat2 = nullptr;
}

int sn_convert_layout(sn_handle *h, bool to_tiled)
{
    dim3 grid((unsigned)((h->G.rep_stride + 255) / 256), h->p.nreplicas);
    sn_convert_layout_kernel<<<grid, 256, 0, h->stream>>>(h->lat, h->lat2, h->G, to_tiled ? 1 : 0);
    SN_CUDA_CHECK(cudaGetLastError());
    return SN_OK;
}

// Bring the canonical array up to date (before anything that reads or partially writes it).
int sn_sync_canonical(sn_handle *h)
{
    if (h->lat_valid) return SN_OK;
    int rc = sn_convert_layout(h, false);
    if (rc) return rc;
    h->lat_valid = true;
    return SN_OK;
}

static SnSweepArgs sn_sweep_args(sn_handle *h);
static int sn_slab_phase_sync(sn_handle *h, long long *launches);

// tile grid, colours and the items-per-phase prefix sums of a lattice (the last tile of an axis may be partial)
static void sn_tile_flow_shape(SnTileFlow &f, int X, int Y, int nz, int nrep, int periodic_z)
{
    f.tnx = (X + snt::T - 1) / snt::T; f.tny = (Y + snt::T - 1) / snt::T; f.tnz = (nz + snt::T - 1) / snt::T; f.nrep = nrep;
    f.ncx = sn_tc_ncol(f.tnx); f.ncy = sn_tc_ncol(f.tny); f.ncz = sn_tc_ncol(f.tnz); f.np = f.ncx * f.ncy * f.ncz;
    f.periodic_z = periodic_z;
    f.pre[0] = 0;
    for (int p = 0; p < f.np; p++) {
        const int cz = p % f.ncz, cy = (p / f.ncz) % f.ncy, cx = p / (f.ncz * f.ncy);
        f.pre[p + 1] = f.pre[p] + (unsigned int)(sn_tc_count(f.tnx, cx) * sn_tc_count(f.tny, cy) * sn_tc_count(f.tnz, cz) * f.nrep);
    }
    for (int p = f.np + 1; p < 28; p++) f.pre[p] = f.pre[f.np];
}

// Host-side view of the tiled kernel's work order (no GPU involved): the items of sweep `sweep` of a periodic X x Y x Z
// lattice in the order the persistent CTAs take them, items[n] = {replica, tx, ty, tz, phase}.  Diagnostic / test aid.
extern "C" int sn_tile_schedule(int X, int Y, int Z, int nreplicas, unsigned long long sweep, int *n_items, int *items, int max_items)
{
    if (!n_items || X < 20 || Y < 20 || Z < 20 || Z % 4 || nreplicas < 1) return sn_fail(SN_ERR_INVALID, "sn_tile_schedule: bad arguments");
    SnTileFlow f;
    sn_tile_flow_shape(f, X, Y, Z, nreplicas, 1);
    f.base_sweep = 0;
    const unsigned long long S = f.pre[f.np];
    *n_items = (int)S;
    if (!items) return SN_OK;
    if ((unsigned long long)max_items < S) return sn_fail(SN_ERR_INVALID, "sn_tile_schedule: %llu items, room for %d", S, max_items);
    for (unsigned long long n = 0; n < S; n++) {
        const SnTileItem it = sn_tile_item(f, sweep * S + n);
        int *o = items + 5 * n;
        o[0] = it.rep; o[1] = it.tx; o[2] = it.ty; o[3] = it.tz; o[4] = it.p;
    }
    return SN_OK;
}

int sn_sweep_tiled_launch(sn_handle *h, long long nsweeps, long long *launches)
{
    const SnGeom &G = h->G;
    const SnTileMaps &tm = *reinterpret_cast<const SnTileMaps *>(h->tmap);
    if (nsweeps <= 0) return SN_OK;
    if (!h->lat2_valid) {                            // first tiled sweep after the canonical array changed
        int rc = sn_sync_canonical(h);
        if (rc || (rc = sn_convert_layout(h, true))) return rc;
        h->lat2_valid = true;
    }
    h->lat_valid = false;
    // Z-slabs: neighbours push into this copy, so nobody may start before everybody's copy is in place
    if (!G.periodic_z) { int rc = sn_slab_phase_sync(h, launches); if (rc) return rc; }

    SnSweepArgs a = sn_sweep_args(h);
    a.lat = h->lat2;                                  // the kernel works on the split copies (own and neighbours')
    SnTileFlow f;
    sn_tile_flow_shape(f, G.X, G.Y, G.nz, h->p.nreplicas, G.periodic_z);
    f.base_sweep = h->sweep;
    f.next = reinterpret_cast<unsigned long long *>(h->flags + SN_FLAGS_NEXT);
    f.ver = h->flags + SN_FLAGS_VER;
    f.peer_ver_lo = G.periodic_z ? f.ver : h->peer_flags[0] + SN_FLAGS_VER;
    f.peer_ver_hi = G.periodic_z ? f.ver : h->peer_flags[1] + SN_FLAGS_VER;
    f.sys_scope = !G.periodic_z;
    f.audit = h->audit_dev;
    f.err = h->flags + SN_FLAGS_ERR;
    f.timeout_ns = h->spin_timeout_ns;
    const unsigned long long S = f.pre[f.np];                                         // items per sweep
    auto launch = [&](unsigned long long n0, unsigned long long n1) -> int {
        f.n_begin = n0; f.n_end = n1;
        SN_CUDA_CHECK(cudaMemsetAsync(f.next, 0, sizeof(unsigned long long), h->stream));
        const int sms = h->grid_limit > 0 ? std::min(h->grid_limit, h->num_sms) : h->num_sms;
        const int grid = (int)std::min<unsigned long long>(n1 - n0, (unsigned long long)sms);
        auto go = [&](auto kern) { kern<<<grid, snt::THREADS, snt::SMEM_BYTES, h->stream>>>(tm, a, f); };
        if (h->p.cutoff == 3) {
            if (f.audit) { if (h->species) go(sn_tiled_kernel<true, true, 3>); else go(sn_tiled_kernel<false, true, 3>); }
            else if (h->species) go(sn_tiled_kernel<true, false, 3>);
            else go(sn_tiled_kernel<false, false, 3>);
        } else {
            if (f.audit) { if (h->species) go(sn_tiled_kernel<true, true, 2>); else go(sn_tiled_kernel<false, true, 2>); }
            else if (h->species) go(sn_tiled_kernel<true, false, 2>);
            else go(sn_tiled_kernel<false, false, 2>);
        }
        if (launches) (*launches)++;
        return SN_OK;
    };
    if (h->p.kernel == SN_KERNEL_TILED_PHASED) {
        // validation mode: one launch per tile-colour phase (stream order = barrier between phases)
        for (unsigned long long sw = 0; sw < (unsigned long long)nsweeps; sw++)
            for (int p = 0; p < f.np; p++) { int rc = launch(sw * S + f.pre[p], sw * S + f.pre[p + 1]); if (rc) return rc; }
    } else {
        int rc = launch(0, (unsigned long long)nsweeps * S);
        if (rc) return rc;
    }
    h->sweep += (unsigned long long)nsweeps;
    SN_CUDA_CHECK(cudaGetLastError());
    // Z-slabs: my ghost planes are complete only once both neighbours have finished too
    if (!G.periodic_z) { int rc = sn_slab_phase_sync(h, launches); if (rc) return rc; }
    return SN_OK;
}
