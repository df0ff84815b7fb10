"""Scratch: where the end-to-end step spends its time (512^3, pinned host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starrynight_b200 as sn
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
host = torch.zeros((size, size, size, 4), dtype=torch.float32).pin_memory()
host[..., 0] = 1.0; host[..., 3] = 1.0
out = torch.empty_like(host).pin_memory()
sim = sn.Simulation(size, size, size)
def t(f, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(*a); sim.lib.sn_synchronize(sim.h); return (time.perf_counter() - t0) * 1e3
for it in range(3):
    a = t(sim.set_lattice_ptr, host.data_ptr()); b = t(sim.MC_sweeps, 20); c = t(sim.get_lattice_ptr, out.data_ptr())
    print(f"set_lattice {a:.1f} ms  sweeps(20) {b:.1f} ms  get_lattice {c:.1f} ms   ({host.numel()*4/1e9/(a*1e-3):.1f} / {host.numel()*4/1e9/(c*1e-3):.1f} GB/s)", flush=True)
d = torch.empty_like(host, device="cuda")
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize(); a = time.perf_counter() - t0
    t0 = time.perf_counter(); out.copy_(d, non_blocking=True); torch.cuda.synchronize(); b = time.perf_counter() - t0
    print(f"plain contiguous H2D {host.numel()*4/1e9/a:.1f} GB/s, D2H {host.numel()*4/1e9/b:.1f} GB/s")
