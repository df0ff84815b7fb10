"""Throughput of the sweep on the other BASELINE.json configurations (parity-test cases, not bench lines):
device-resident attempts/s, CUDA events on the library's stream.  Writes gpurun_out/configs.json.
  C2  100x100x1, T = 300 K, Efield.x = 0.02, 64 replicas          (colour-pass kernel, 28 neighbours)
  C3  64^3, CageStrain 1, T = 0..500 K step 25 as 21 replicas     (tiled kernel)
  C4  128^3, Dipoles [1.0, 0.5, 0.0] Prevalence [0.6, 0.3, 0.1]   (tiled kernel, species path)
  C5b 1024x1024x128: the per-GPU slab of a 1024^3 run on 8 GPUs, run here as a periodic box"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import starrynight_b200 as sn
from oracle import oracle_api as oa

def unit_lattice(X, Y, Z, seed, lengths=(1.0,), prevalence=(1.0,)):
    rng = np.random.default_rng(seed)
    lat = np.zeros((X, Y, Z, 4), np.float32)
    v = rng.standard_normal((X, Y, Z, 3), dtype=np.float32)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    lat[..., :3] = v
    lat[..., 3] = rng.choice(np.asarray(lengths, np.float32), size=(X, Y, Z), p=np.asarray(prevalence) / np.sum(prevalence))
    return lat

def run(name, X, Y, Z, reps, sweeps, temps=None, **kw):
    lat_kw = {k: kw.pop(k) for k in ("lengths", "prevalence") if k in kw}
    sim = sn.Simulation(X, Y, Z, nreplicas=reps, **kw)
    lat = unit_lattice(X, Y, Z, 1, **lat_kw)
    for r in range(reps):
        sim.set_lattice(lat, r)
        if temps is not None:
            sim.set_T(temps[r], r)
    sim.MC_sweeps_timed(max(1, sweeps // 4))
    best, launches = 1e30, 0
    for _ in range(3):
        ms, launches = sim.MC_sweeps_timed(sweeps)
        best = min(best, ms)
    acc = [sim.counters(r) for r in range(reps)]
    n = X * Y * Z * reps
    nb = 28 if Z == 1 else 122
    flop = 20 * nb + 6 * (4 if Z == 1 else 6) + 24
    rate = n * sweeps / best * 1e3
    row = {"config": name, "lattice": [X, Y, Z], "replicas": reps, "sweeps_timed": sweeps, "ms_per_sweep": best / sweeps,
           "attempts_per_s": rate, "algorithmic_tflops": rate * flop / 1e12, "flop_per_attempt": flop, "launches": launches,
           "accept_ratio_first_last": [acc[0][0] / max(1, acc[0][0] + acc[0][1]), acc[-1][0] / max(1, acc[-1][0] + acc[-1][1])]}
    print(json.dumps(row), flush=True)
    sim.close()
    return row

rows = []
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
rows.append(run("C2 100x100x1 Efield.x=0.02, 64 replicas", 100, 100, 1, 64, 200, Efield=(0.02, 0, 0)))
rows.append(run("C2 100x100x1 Efield.x=0.02, 1024 replicas", 100, 100, 1, 1024, 50, Efield=(0.02, 0, 0)))
rows.append(run("C2 100x100x1 Efield.x=0.02, single lattice", 100, 100, 1, 1, 400, Efield=(0.02, 0, 0)))
rows.append(run("C2 100x100x1, 64 replicas, colour-pass kernel (one launch per colour)", 100, 100, 1, 64, 100, Efield=(0.02, 0, 0), kernel=sn.SN_KERNEL_COLOUR))
rows.append(run("C1 20x20x28 (make test lattice), 148 replicas", 20, 20, 28, 148, 40))
rows.append(run("C1 20x20x28 (make test lattice), single lattice", 20, 20, 28, 1, 100))
if QUICK:
    raise SystemExit(0)
rows.append(run("C3 64^3 CageStrain=1, T=0..500 K step 25 (21 replicas)", 64, 64, 64, 21, 40, temps=list(range(0, 501, 25))))
rows.append(run("C4 128^3 solid solution (lengths 1.0/0.5/0.0 at 0.6/0.3/0.1)", 128, 128, 128, 1, 40, lengths=(1.0, 0.5, 0.0), prevalence=(0.6, 0.3, 0.1), Efield=(0.05, 0, 0)))
rows.append(run("C4 128^3 solid solution, 8 field replicas", 128, 128, 128, 8, 20, lengths=(1.0, 0.5, 0.0), prevalence=(0.6, 0.3, 0.1), Efield=(0.05, 0, 0)))
rows.append(run("C5b 1024x1024x128 (per-GPU share of 1024^3 on 8 GPUs)", 1024, 1024, 128, 1, 5))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/configs.json", "w"), indent=1)
