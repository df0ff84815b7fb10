"""Scratch: which part of the double-buffered end-to-end step fails to overlap on Z-slabs?  (torchrun, one rank per GPU)
Runs the bench's pipelined loop with single ingredients left out (timing only)."""
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
outs = [torch.empty_like(host).pin_memory() for _ in range(2)]
def mk():
    s = sn.Simulation(X, X, X, device=local, z0=rank * nz if world > 1 else 0, nz=nz if world > 1 else 0)
    if world > 1:
        sn_slab.wire_ipc(s, dist, world, rank)
    return s
sims = [mk(), mk()]
for s in sims:
    s.set_lattice_ptr(host.data_ptr()); s.pull_ghosts()
def bar():
    dist.barrier(); torch.cuda.synchronize()
def loop(nsteps, up=True, down=True, pull=True, sweeps=20, pull_first=False):
    bar(); t0 = time.perf_counter()
    for i in range(nsteps):
        s, o = sims[i % 2], sims[(i - 1) % 2]
        s.synchronize()
        if up: s.set_lattice_async(host.data_ptr(), 0)
        if pull and pull_first: s.pull_ghosts()
        s.order_after(o)
        if pull and not pull_first: s.pull_ghosts()
        if sweeps: s.MC_sweeps(sweeps)
        if down: s.get_lattice_async(outs[i % 2].data_ptr(), 0)
    for s in sims: s.synchronize()
    bar(); dt = time.perf_counter() - t0
    v = torch.tensor([dt], device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MAX); return float(v.item()) / nsteps * 1e3
loop(4)
for name, kw in [("full", {}), ("pull_ghosts before order_after", dict(pull_first=True)), ("no download", dict(down=False)), ("no upload", dict(up=False)), ("no upload, no download", dict(up=False, down=False)),
                 ("no pull_ghosts", dict(pull=False)), ("copies only (no sweeps)", dict(sweeps=0)), ("upload only", dict(sweeps=0, down=False)), ("download only", dict(sweeps=0, up=False))]:
    ms = loop(12, **kw)
    if rank == 0: print(f"N={world}: {name:28s} {ms:7.2f} ms per step", flush=True)
for s in sims: s.close()
dist.barrier(); dist.destroy_process_group()
