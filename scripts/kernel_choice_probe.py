"""Scratch: tiled vs colour-pass kernel on mid-size single lattices (which should SN_KERNEL_AUTO pick?)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
rng = np.random.default_rng(1)
for n, reps in ((32, 1), (48, 1), (64, 1), (80, 1), (96, 1), (128, 1), (64, 4), (32, 16)):
    lat = np.zeros((n, n, n, 4), np.float32)
    v = rng.standard_normal((n, n, n, 3), dtype=np.float32); lat[..., :3] = v / np.linalg.norm(v, axis=-1, keepdims=True); lat[..., 3] = 1
    row = []
    for kern, name in ((sn.SN_KERNEL_TILED, "tiled"), (sn.SN_KERNEL_COLOUR, "colour")):
        with sn.Simulation(n, n, n, nreplicas=reps, kernel=kern) as sim:
            for r in range(reps): sim.set_lattice(lat, r)
            sweeps = 20 if kern == sn.SN_KERNEL_TILED else 5
            sim.MC_sweeps_timed(sweeps)
            ms = min(sim.MC_sweeps_timed(sweeps)[0] for _ in range(3))
            row.append("%s %.3e" % (name, n ** 3 * reps * sweeps / ms * 1e3))
    print(f"{n}^3 x{reps}: tiles per phase {(n // 16) ** 3 // 8 * reps if n % 32 == 0 else '-'}: " + ", ".join(row), flush=True)
