"""CPU tests (gloo, world_size 2 and 3) of the Z-slab host logic: slab ownership, ring
neighbours and the ghost-plane bootstrap exchange.  The device is replaced by a numpy
stand-in with the same get_boundary / set_ghost surface as starrynight_b200.Simulation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starrynight_b200 import slab


class FakeSlab:
    """numpy model of a slab handle: own planes plus g ghost planes on each side."""

    def __init__(self, full, z0, nz, g=3):
        self.g, self.nz = g, nz
        self.own = full[:, :, z0:z0 + nz].copy()
        self.ghost = [None, None]

    def get_boundary(self, side, replica=0):
        return self.own[:, :, :self.g] if side == 0 else self.own[:, :, self.nz - self.g:]

    def set_ghost(self, side, planes, replica=0):
        self.ghost[side] = np.array(planes, copy=True)

    # observables of the own planes, same surface as Simulation
    @property
    def nsites(self):
        return self.own.shape[0] * self.own.shape[1] * self.nz

    def polarisation(self, replica=0):
        return self.own[..., :3].astype(np.float64).reshape(-1, 3).mean(0)

    def total_energy(self, precision=0, replica=0):
        return np.array([self.own[..., 0].astype(np.float64).sum(), 1.0, 2.0, 0.0])

    def counters(self, replica=0):
        return (10 * self.nz, 20 * self.nz, self.nz)


def test_slab_range_and_ring():
    assert slab.slab_range(512, 8, 3, 32) == (192, 64)
    assert slab.slab_range(64, 1, 0, 32) == (0, 64)
    assert slab.ring_neighbours(8, 0) == (7, 1)
    assert slab.ring_neighbours(2, 1) == (0, 0)
    with pytest.raises(ValueError):
        slab.slab_range(512, 3, 0, 32)
    with pytest.raises(ValueError):
        slab.slab_range(64, 4, 0, 32)
    with pytest.raises(ValueError):
        slab.slab_range(64, 2, 2, 4)
    # slabs tile the axis exactly once
    for world in (1, 2, 4, 8):
        cover = np.zeros(256, int)
        for r in range(world):
            z0, nz = slab.slab_range(256, world, r, 32)
            cover[z0:z0 + nz] += 1
        assert np.all(cover == 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Z, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full = rng.standard_normal((6, 5, Z, 4)).astype(np.float32)       # same on every rank
    z0, nz = slab.slab_range(Z, world, rank, 4)
    sim = FakeSlab(full, z0, nz)
    slab.exchange_ghosts(sim, dist, world, rank)
    below = np.take(full, [(z0 - 3 + i) % Z for i in range(3)], axis=2)
    above = np.take(full, [(z0 + nz + i) % Z for i in range(3)], axis=2)
    ok = np.array_equal(sim.ghost[0], below) and np.array_equal(sim.ghost[1], above)
    m = slab.merge_observables(sim, dist, world)                       # one FP64 all_reduce
    ok = ok and np.allclose(m["polarisation"], full[..., :3].astype(np.float64).reshape(-1, 3).mean(0), rtol=1e-13, atol=1e-15)
    ok = ok and np.isclose(m["energy"][0], full[..., 0].astype(np.float64).sum(), rtol=1e-13) and m["energy"][1] == world
    ok = ok and (m["accept"], m["reject"], m["vacant"], m["nsites"]) == (10 * Z, 20 * Z, Z, 6 * 5 * Z)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,Z", [(2, 16), (3, 24), (2, 8)])
def test_ghost_exchange_over_gloo(world, Z):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Z, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) == 1


def test_single_rank_exchange_is_a_no_op():
    """world == 1: the handle owns the whole Z axis, its ghosts are its own periodic images (sn_set_ghost refuses
    such a handle), so there is nothing to exchange."""
    rng = np.random.default_rng(1)
    full = rng.standard_normal((4, 4, 12, 4)).astype(np.float32)
    sim = FakeSlab(full, 0, 12)
    before = [None if g is None else g.copy() for g in sim.ghost]
    slab.exchange_ghosts(sim, None, 1, 0)
    assert all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(before, sim.ghost))
