"""Scratch GPU check: smoke + first timings of the sweep kernels."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.smoke()
import starrynight_b200 as sn
from oracle import oracle_api as oa
for (X, Y, Z, reps) in [(20, 20, 28, 1), (64, 64, 64, 1), (128, 128, 128, 1), (256, 256, 256, 1), (100, 100, 1, 64)]:
    lat = oa.random_lattice(X, Y, Z, seed=1)
    sim = sn.Simulation(X, Y, Z, nreplicas=reps)
    for r in range(reps):
        sim.set_lattice(lat, r)
    sim.MC_sweeps_timed(2)
    ms, n = sim.MC_sweeps_timed(5)
    acc, rej, vac = sim.counters()
    print(f"{X}x{Y}x{Z} x{reps}: {ms/5:.3f} ms/sweep, {X*Y*Z*reps*5/ms*1e3:.3e} attempts/s, launches {n}, accept {acc/(acc+rej):.3f}", flush=True)
    sim.close()
