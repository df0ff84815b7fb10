"""Z-slab decomposition helpers (host side).

The reference has no distribution; large lattices are cut
into contiguous Z-slabs, one per GPU / rank, periodic ring in Z.  These helpers
hold the bookkeeping that is independent of the device: which planes a rank owns,
who its ring neighbours are, and the bootstrap exchange of ghost planes through
``torch.distributed`` (any backend; the CPU tests run it over gloo).  During
sweeps the kernels push boundary updates GPU-to-GPU themselves
(sn_ipc_attach / sn_attach_peer); this path only fills the ghost planes once
after the lattice has been uploaded.
"""
from __future__ import annotations

import numpy as np


def slab_range(Z: int, world: int, rank: int, multiple: int = 4):
    """(z0, nz) of `rank`'s slab.  All slabs are equal and a multiple of `multiple`
    planes (cutoff+1 for the colour kernel, 32 for the tiled kernel)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of {world}")
    if Z % world or (Z // world) % multiple:
        raise ValueError(f"Z={Z} cannot be cut into {world} slabs that are multiples of {multiple} planes")
    nz = Z // world
    return rank * nz, nz


def ring_neighbours(world: int, rank: int):
    """(lower, upper): the ranks owning the planes below z0 and above z0+nz-1."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_ghosts(sim, dist, world: int, rank: int, replica: int = 0):
    """Fill `sim`'s ghost planes from the ring neighbours' boundary planes.

    `sim` needs get_boundary(side, replica) -> float32[X][Y][g][4] and
    set_ghost(side, planes, replica); `dist` is torch.distributed (initialised).
    My lowest planes become the lower neighbour's upper ghost, my highest planes the
    upper neighbour's lower ghost."""
    import torch
    if world == 1:
        return                                        # the handle owns the whole axis: its ghosts are its own periodic images
    lo, hi = ring_neighbours(world, rank)
    mine_low = torch.from_numpy(np.ascontiguousarray(sim.get_boundary(0, replica)))
    mine_high = torch.from_numpy(np.ascontiguousarray(sim.get_boundary(1, replica)))
    got = [torch.empty_like(mine_low) for _ in range(2 * world)]
    # all_gather keeps the exchange deadlock-free on every backend (2 small tensors per rank)
    dist.all_gather(got[:world], mine_low)
    dist.all_gather(got[world:], mine_high)
    sim.set_ghost(0, got[world + lo].numpy(), replica)     # below me: the lower neighbour's highest planes
    sim.set_ghost(1, got[hi].numpy(), replica)             # above me: the upper neighbour's lowest planes


def wire_ipc(sim, dist, world: int, rank: int):
    """Exchange CUDA IPC handles around the ring and attach both neighbours."""
    handles = [None] * world
    dist.all_gather_object(handles, sim.ipc_export())
    lo, hi = ring_neighbours(world, rank)
    sim.ipc_attach(0, *handles[lo])
    sim.ipc_attach(1, *handles[hi])


def merge_observables(sim, dist, world: int, replica: int = 0, precision: int = 0):
    """Lattice-wide observables of a slab-decomposed lattice from the slabs' own reductions: ONE small FP64
    all_reduce (the only collective on this path; the sweep itself exchanges nothing through it).

    Returns dict(polarisation[3] = (1/N) sum p, energy[4] = pair, cage, field, K terms, accept, reject, vacant).
    `sim` needs polarisation(replica) -> mean over its slab, total_energy(precision, replica) -> its slab's share
    (pair terms across a seam are counted half on each side), counters(replica) and nsites."""
    import torch
    P = np.asarray(sim.polarisation(replica), np.float64) * sim.nsites
    E = np.asarray(sim.total_energy(precision, replica), np.float64)
    c = np.asarray(sim.counters(replica), np.float64)            # exact below 2^53 attempts
    t = torch.from_numpy(np.concatenate([P, E, c, [float(sim.nsites)]]))
    if world > 1:
        tt = t.cuda() if dist.get_backend() == "nccl" else t      # NCCL reduces device tensors only
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        t = tt.cpu()
    v = t.numpy()
    return dict(polarisation=v[0:3] / v[10], energy=v[3:7].copy(), accept=int(v[7]), reject=int(v[8]), vacant=int(v[9]), nsites=int(v[10]))
