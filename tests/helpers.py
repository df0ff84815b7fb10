"""Shared helpers for the test-suite."""
import glob
import os

import numpy as np

from oracle import oracle_api as oa

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = ["species3d", "flat2d", "odd_cut2", "cut4_constrain", "tiny", "dim2"]


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    X, Y, Z, cut, constrain, dim = [int(v) for v in g["params"]]
    cage, K, ex, ey, ez, beta = [float(v) for v in g["couplings"]]
    p = oa.make_params(X, Y, Z, cut, cage, K, (ex, ey, ez), beta, constrain, dim, 300)
    return g, p


def sim_for(sn, p, lat=None, **kw):
    """A Simulation configured like oracle params `p`."""
    sim = sn.Simulation(p.X, p.Y, p.Z, DipoleCutOff=p.cutoff, CageStrain=p.CageStrain, K=p.K,
                        Efield=tuple(p.Efield), beta=p.beta, ConstrainToX=bool(p.ConstrainToX), DIM=p.DIM, **kw)
    if lat is not None:
        sim.set_lattice(lat)
    return sim


def random_moves(p, n, seed=0):
    rng = np.random.default_rng(seed)
    sites = np.stack([rng.integers(0, p.X, n), rng.integers(0, p.Y, n), rng.integers(0, p.Z, n)], 1).astype(np.int32)
    nd = rng.normal(size=(n, 3))
    nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    return sites, nd.astype(np.float32)


def term_scale(p, lat, sites, newdip):
    """sum_j |term_j| for each trial move: the scale the FP32 1e-5 bar is relative to
    (the native reference is itself only float-accurate in these units)."""
    lat = np.asarray(lat, np.float64)
    dxyz, d = oa.Oracle("f32").neighbours(p)
    out = np.zeros(len(sites))
    for i, ((x, y, z), nd) in enumerate(zip(sites, np.asarray(newdip, np.float64))):
        old = lat[x, y, z]
        dp = nd - old[:3]
        s = 0.0
        for (dx, dy, dz), dd in zip(dxyz, d):
            t = lat[(x + dx) % p.X, (y + dy) % p.Y, (z + dz) % p.Z]
            n = np.array([dx, dy, dz]) / dd
            s += abs(old[3] * t[3]) * (abs(dp @ t[:3]) + 3 * abs((n @ dp) * (n @ t[:3]))) / dd ** 3
            if dx * dx + dy * dy + dz * dz == 1:
                s += abs(p.CageStrain * (dp @ t[:3]))
        s += abs(dp @ np.array(p.Efield[:])) + abs(p.K) * 4
        out[i] = s
    return out


# ---- Philox4x32-10 (Salmon et al., SC'11), written from the paper's round function, independently of the library ----
PHILOX_KAT = [   # Random123 kat_vectors: counter[4], key[2] -> output[4]
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised numpy Philox4x32-10; returns four uint32 arrays."""
    m, s = np.uint64(0xFFFFFFFF), np.uint64(32)
    c0, c1, c2, c3, k0, k1 = [np.asarray(v, np.uint64) & m for v in (c0, c1, c2, c3, k0, k1)]
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> s) ^ c1 ^ k0, p1 & m, (p0 >> s) ^ c3 ^ k1, p0 & m
        k0, k1 = (k0 + W0) & m, (k1 + W1) & m
    return [v.astype(np.uint32) for v in (c0, c1, c2, c3)]


def term_scale_batch(p, lat, sites, newdip, dxyz, d):
    """Vectorised term_scale: sum_j |term_j| of site_energy (montecarlo-core.c:99-134) for a batch of trial moves."""
    lat = np.asarray(lat, np.float64)
    sites = np.asarray(sites, np.int64)
    old = lat[sites[:, 0], sites[:, 1], sites[:, 2]]                 # (n, 4)
    dp = np.asarray(newdip, np.float64) - old[:, :3]                  # (n, 3)
    nx = (sites[:, None, 0] + dxyz[None, :, 0]) % p.X
    ny = (sites[:, None, 1] + dxyz[None, :, 1]) % p.Y
    nz = (sites[:, None, 2] + dxyz[None, :, 2]) % p.Z
    t = lat[nx, ny, nz]                                               # (n, nb, 4)
    nhat = dxyz.astype(np.float64) / d[:, None]                       # (nb, 3)
    dpt = np.abs(np.einsum("nc,nbc->nb", dp, t[..., :3]))
    ndp = np.abs(dp @ nhat.T)                                         # (n, nb)
    npt = np.abs(np.einsum("bc,nbc->nb", nhat, t[..., :3]))
    s = (np.abs(old[:, 3:4] * t[..., 3]) * (dpt + 3 * ndp * npt) / d[None, :] ** 3).sum(1)
    nn = (dxyz ** 2).sum(1) == 1
    s += np.abs(p.CageStrain) * dpt[:, nn].sum(1)
    s += np.abs(dp @ np.array(p.Efield[:])) + abs(p.K) * 4
    return s
