"""Scratch: one resident-kernel launch for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
X, Y, Z, R, S = 100, 100, 1, 148, 50
rng = np.random.default_rng(1)
lat = np.zeros((X, Y, Z, 4), np.float32)
v = rng.standard_normal((X, Y, Z, 3), dtype=np.float32); v /= np.linalg.norm(v, axis=-1, keepdims=True)
lat[..., :3] = v; lat[..., 3] = 1
sim = sn.Simulation(X, Y, Z, nreplicas=R, Efield=(0.02, 0, 0))
for r in range(R):
    sim.set_lattice(lat, r)
sim.MC_sweeps(2)
sim.MC_sweeps(S)
print(sim.counters())
