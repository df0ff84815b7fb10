import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
from oracle import oracle_api as oa
for (X, Y, Z) in [(10, 9, 12), (13, 9, 11), (13, 16, 16), (16, 9, 16), (16, 16, 11), (16, 16, 16), (12, 9, 11), (24, 8, 8), (13, 8, 8), (5, 4, 3), (17, 8, 1)]:
    p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, (0, 0, 0), 1.0)
    lat = oa.random_lattice(X, Y, Z, seed=X, lengths=(1.0, 0.5), prevalence=(0.7, 0.3))
    ref = oa.Oracle("f64").potential_map(p, lat).reshape(X, Y, Z)
    with sn.Simulation(X, Y, Z) as sim:
        sim.set_lattice(lat)
        back = sim.get_lattice()
        V = sim.dipole_potential()
        V2 = sim.dipole_potential()
    err = np.abs(V - ref)
    bad = np.argwhere(err > 1e-9)
    print((X, Y, Z), "upload ok", np.array_equal(back, lat), "max err", err.max(), "repeat equal", np.array_equal(V, V2), "nbad", len(bad), "first bad", bad[:4].tolist(), flush=True)
