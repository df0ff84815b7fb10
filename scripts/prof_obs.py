"""One call of each tiled observable kernel on an N^3 lattice (for ncu): usage prof_obs.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
X = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(1)
lat = np.zeros((X, X, X, 4), np.float32)
v = rng.standard_normal((X, X, X, 3), dtype=np.float32)
lat[..., :3] = v / np.linalg.norm(v, axis=-1, keepdims=True)
lat[..., 3] = 1
with sn.Simulation(X, X, X) as sim:
    sim.set_lattice(lat)
    sim.radial_order_parameter()
    sim.dipole_potential()
    sim.dipole_electricfield(4, False)
print("done")
