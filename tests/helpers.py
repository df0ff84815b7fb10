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
