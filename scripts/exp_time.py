"""Scratch: time the tiled sweep of several library variants (SN_B200_LIB) in one GPU call.
usage: exp_time.py XxYxZ sweeps lib1.so lib2.so ...   (each variant runs in a subprocess)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import starrynight_b200 as sn
    X, Y, Z = [int(v) for v in sys.argv[2].split("x")]
    sweeps = int(sys.argv[3])
    rng = np.random.default_rng(1)
    lat = np.zeros((X, Y, Z, 4), np.float32)
    v = rng.standard_normal((X, Y, Z, 3), dtype=np.float32)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    lat[..., :3] = v; lat[..., 3] = 1
    sim = sn.Simulation(X, Y, Z, kernel=sn.SN_KERNEL_TILED)
    sim.set_lattice(lat)
    sim.MC_sweeps_timed(sweeps)
    best = 1e9
    for _ in range(3):
        ms, n = sim.MC_sweeps_timed(sweeps)
        best = min(best, ms)
    acc, rej, vac = sim.counters()
    print(f"{os.path.basename(os.environ.get('SN_B200_LIB', 'default')):24s} {X}x{Y}x{Z}: {best/sweeps:.3f} ms/sweep, {X*Y*Z*sweeps/best*1e3:.4e} attempts/s, accept {acc/max(1,acc+rej):.4f}", flush=True)
    sim.close()
else:
    shape, sweeps, libs = sys.argv[1], sys.argv[2], sys.argv[3:]
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["SN_B200_LIB"] = os.path.join(ROOT, lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", shape, sweeps], env=env, timeout=300)
