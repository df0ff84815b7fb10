"""One small run of a sweep kernel for compute-sanitizer (scripts/sanitize.sh): usage sanitize_case.py tiled|resident|colour"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn

which = sys.argv[1]
X, Y, Z, kern = {"tiled": (64, 32, 32, sn.SN_KERNEL_TILED), "resident": (20, 20, 28, sn.SN_KERNEL_RESIDENT),
                 "colour": (16, 12, 16, sn.SN_KERNEL_COLOUR), "partial": (40, 35, 44, sn.SN_KERNEL_TILED),       # every axis ends in a cut tile
                 "cut2": (32, 48, 32, sn.SN_KERNEL_TILED)}[which]                                                  # DipoleCutOff 2: box starts outside the array
cutoff = 2 if which == "cut2" else 3
rng = np.random.default_rng(3)
lat = np.zeros((X, Y, Z, 4), np.float32)
v = rng.standard_normal((X, Y, Z, 3)).astype(np.float32)
lat[..., :3] = v / np.linalg.norm(v, axis=-1, keepdims=True)
lat[..., 3] = rng.choice(np.array([1.0, 0.5, 0.0], np.float32), size=(X, Y, Z), p=[0.7, 0.2, 0.1])
with sn.Simulation(X, Y, Z, DipoleCutOff=cutoff, CageStrain=1.0, Efield=(0.02, 0, 0), nreplicas=2, seed=11, kernel=kern) as sim:
    for r in range(2):
        sim.set_lattice(lat, r)
    sim.MC_sweeps(2)
    acc, rej, vac = sim.counters()
    out = sim.get_lattice()
    assert acc + rej + vac == 2 * X * Y * Z
    assert np.array_equal(out[..., 3], lat[..., 3])
    print(which, "ok: accept ratio %.3f" % (acc / (acc + rej)), "hash %016x" % sim.state_hash())
