"""Scratch: wall time of the lattice-wide observables at a given size (device-resident lattice)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
X = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(1)
lat = np.zeros((X, X, X, 4), np.float32)
for x0 in range(0, X, 64):
    v = rng.standard_normal((min(64, X - x0), X, X, 3), dtype=np.float32); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    lat[x0:x0 + 64, ..., :3] = v
lat[..., 3] = 1
sim = sn.Simulation(X, X, X)
sim.set_lattice(lat)
def t(name, f, *a):
    f(*a); t0 = time.perf_counter(); r = f(*a); dt = time.perf_counter() - t0
    print(f"{X}^3 {name:28s} {dt*1e3:9.2f} ms", flush=True); return r
t("polarisation", sim.polarisation)
t("landau_order", sim.landau_order)
t("total_energy F32", sim.total_energy, sn.SN_PREC_F32)
t("total_energy F64", sim.total_energy, sn.SN_PREC_F64)
t("radial_order_parameter", sim.radial_order_parameter)
t("dipole_potential (incl. D2H)", sim.dipole_potential)
t("dipole_electricfield", sim.dipole_electricfield, 4, False)
t("recombination", sim.recombination)
