"""Scratch: one tiled launch for ncu. usage: prof_one.py XxYxZ sweeps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
X, Y, Z = [int(v) for v in sys.argv[1].split("x")]
rng = np.random.default_rng(1)
lat = np.zeros((X, Y, Z, 4), np.float32)
v = rng.standard_normal((X, Y, Z, 3), dtype=np.float32); v /= np.linalg.norm(v, axis=-1, keepdims=True)
lat[..., :3] = v; lat[..., 3] = 1
sim = sn.Simulation(X, Y, Z, kernel=sn.SN_KERNEL_TILED)
sim.set_lattice(lat)
sim.MC_sweeps(1)
sim.MC_sweeps(int(sys.argv[2]))
print(sim.counters())
