"""Scratch GPU check: timings of the sweep kernels."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
from oracle import oracle_api as oa
cases = [(64, 64, 64, 1), (128, 128, 128, 1), (256, 256, 256, 1), (512, 512, 64, 1), (512, 512, 512, 1)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) + (1,) for a in sys.argv[1:]]
for (X, Y, Z, reps) in cases:
    rng = np.random.default_rng(1)
    lat = np.zeros((X, Y, Z, 4), np.float32)
    v = rng.standard_normal((X, Y, Z, 3), dtype=np.float32)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    lat[..., :3] = v; lat[..., 3] = 1
    for kern in (sn.SN_KERNEL_TILED, sn.SN_KERNEL_COLOUR):
        if kern == sn.SN_KERNEL_COLOUR and X * Y * Z > 256 ** 3: continue
        sim = sn.Simulation(X, Y, Z, nreplicas=reps, kernel=kern)
        for r in range(reps):
            sim.set_lattice(lat, r)
        sim.MC_sweeps_timed(2)
        ms, n = sim.MC_sweeps_timed(5)
        acc, rej, vac = sim.counters()
        e = sim.total_energy(sn.SN_PREC_F32).sum() / (X * Y * Z)
        print(f"{X}x{Y}x{Z} x{reps} kernel={kern}: {ms/5:.3f} ms/sweep, {X*Y*Z*reps*5/ms*1e3:.3e} attempts/s, launches {n}, accept {acc/max(1,acc+rej):.4f}, E/N {e:.5f}", flush=True)
        sim.close()
