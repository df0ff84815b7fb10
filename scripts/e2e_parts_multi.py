"""Scratch: where the end-to-end step of a Z-slab run spends its time (torchrun, one rank per GPU).
Every phase is bracketed by a barrier, so the numbers are the slowest rank's with all ranks doing the same thing at once."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import starrynight_b200 as sn
from starrynight_b200 import slab as sn_slab
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = 512; nz = X // world
host = torch.zeros((X, X, nz, 4), dtype=torch.float32).pin_memory(); host[..., 0] = 1.0; host[..., 3] = 1.0
out = torch.empty_like(host).pin_memory()
sim = sn.Simulation(X, X, X, device=local, z0=rank * nz if world > 1 else 0, nz=nz if world > 1 else 0)
if world > 1:
    sn_slab.wire_ipc(sim, dist, world, rank)
def bar():
    dist.barrier(); torch.cuda.synchronize()
def t(f, *a):
    bar(); t0 = time.perf_counter(); f(*a); sim.synchronize(); dt = time.perf_counter() - t0
    v = torch.tensor([dt], device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MAX); return float(v.item()) * 1e3
gb = host.numel() * 4 / 1e9
for it in range(3):
    a = t(sim.set_lattice_ptr, host.data_ptr()); p = t(sim.pull_ghosts); b = t(sim.MC_sweeps, 20); c = t(sim.get_lattice_ptr, out.data_ptr())
    if rank == 0:
        print(f"N={world}: set_lattice {a:.1f} ms ({gb/(a*1e-3):.1f} GB/s per rank)  pull_ghosts {p:.1f}  sweeps(20) {b:.1f}  get_lattice {c:.1f} ms ({gb/(c*1e-3):.1f} GB/s per rank)", flush=True)
d = torch.empty_like(host, device="cuda")
for it in range(2):
    bar(); t0 = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize(); a = time.perf_counter() - t0
    bar(); t0 = time.perf_counter(); out.copy_(d, non_blocking=True); torch.cuda.synchronize(); b = time.perf_counter() - t0
    print(f"rank {rank}: plain contiguous H2D {gb/a:.1f} GB/s, D2H {gb/b:.1f} GB/s (all ranks at once)", flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m | head -20; numactl -H 2>/dev/null | head -5; lscpu | grep -i 'numa\\|socket\\|^CPU(s)'")
sim.close(); dist.barrier(); dist.destroy_process_group()
