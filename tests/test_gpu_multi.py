"""GPU tests of the Z-slab path.  Slab handles in one process, wired with sn_attach_peer: boundary updates
travel as stores from inside the sweep kernels into the neighbour slab's ghost planes (P2P over NVLink when the
slabs sit on different GPUs), phases are ordered by device-side flags / tile versions.  Because Philox counters are
keyed by the GLOBAL site and the phase / colour order is global, the decomposed chain must be bit-identical to the
single-GPU chain.  On a box with >= 2 GPUs the slabs are spread over the GPUs; on a one-GPU box the same tests run
with every slab on device 0 (the slabs share its SMs) so that the decomposition logic is always exercised; the tests
that are only meaningful across GPUs are skipped there."""
import numpy as np
import pytest

from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


@pytest.fixture(scope="module")
def devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n >= 2 else [0]


def _run_split(sn, lat, nslab, kernel, sweeps, devices, per_call=1, pull=True):
    X, Y, Z = lat.shape[:3]
    nz = Z // nslab
    sims = [sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, device=devices[r % len(devices)], z0=r * nz, nz=nz, kernel=kernel)
            for r in range(nslab)]
    for r, s in enumerate(sims):
        s.set_lattice(lat[:, :, r * nz:(r + 1) * nz])
    for r, s in enumerate(sims):
        lo, hi = sims[(r - 1) % nslab], sims[(r + 1) % nslab]
        if not pull:                                   # bootstrap through the host (sn_get_boundary / sn_set_ghost)
            s.set_ghost(0, lo.get_boundary(1))
            s.set_ghost(1, hi.get_boundary(0))
        s.attach_peer(0, lo)
        s.attach_peer(1, hi)
    if pull:                                           # device to device (sn_pull_ghosts), every slab
        for s in sims:
            s.pull_ghosts()
    for _ in range(sweeps // per_call):     # few sweeps per call per slab: the launch queues never fill
        for s in sims:
            s.MC_sweeps(per_call)
    out = np.concatenate([s.get_lattice() for s in sims], axis=2)
    _run_split.last_hash = sum(s.state_hash() for s in sims) % (1 << 64)
    counters = np.sum([s.counters() for s in sims], axis=0)
    energy = np.sum([s.total_energy(sn.SN_PREC_F64) for s in sims], axis=0)
    for s in sims:
        s.close()
    return out, counters, energy


@pytest.mark.parametrize("shape,kernel", [((64, 32, 64), "tiled"), ((16, 12, 16), "colour"), ((32, 32, 128), "tiled"), ((40, 52, 64), "tiled")])
def test_two_slabs_match_one_gpu_bit_for_bit(sn, devices, shape, kernel):
    X, Y, Z = shape
    kid = sn.SN_KERNEL_TILED if kernel == "tiled" else sn.SN_KERNEL_COLOUR
    lat = oa.random_lattice(X, Y, Z, seed=21, lengths=(1.0, 0.5, 0.0), prevalence=(0.8, 0.15, 0.05))
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, kernel=kid) as one:
        one.set_lattice(lat)
        one.MC_sweeps(3)
        ref = one.get_lattice()
        ref_c = np.array(one.counters())
        ref_e = one.total_energy(sn.SN_PREC_F64)
        ref_h = one.state_hash()
    out, counters, energy = _run_split(sn, lat, 2, kid, 3, devices=devices)
    assert np.array_equal(out, ref), "slab-decomposed chain differs from the single-GPU chain"
    out_h, _, _ = _run_split(sn, lat, 2, kid, 3, devices=devices, pull=False)
    assert np.array_equal(out_h, ref), "ghost planes set through the host give a different chain"
    assert _run_split.last_hash == ref_h, "the slabs' state hashes do not add up to the single-GPU hash"
    assert np.array_equal(counters, ref_c)
    assert np.allclose(energy, ref_e, rtol=1e-12, atol=1e-9)


def test_four_slabs_on_two_gpus(sn, devices):
    """More slabs than GPUs (two per device): exercises ring wiring beyond the 2-GPU special case."""
    X, Y, Z = 32, 32, 128
    lat = oa.random_lattice(X, Y, Z, seed=22)
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, kernel=sn.SN_KERNEL_TILED) as one:
        one.set_lattice(lat)
        one.MC_sweeps(2)
        ref = one.get_lattice()
    out, _, _ = _run_split(sn, lat, 4, sn.SN_KERNEL_TILED, 2, devices=devices[:2])
    assert np.array_equal(out, ref)


def test_eight_slabs(sn, devices):
    """The 8-way decomposition of the headline run (512^3 over 8 GPUs) in miniature: 8 slabs of 32 planes, spread over
    however many GPUs the box has; lattice, counters and state hash equal the single-GPU chain."""
    X, Y, Z = 64, 32, 256
    lat = oa.random_lattice(X, Y, Z, seed=25, lengths=(1.0, 0.5, 0.0), prevalence=(0.8, 0.15, 0.05))
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, kernel=sn.SN_KERNEL_TILED) as one:
        one.set_lattice(lat)
        one.MC_sweeps(4)
        ref = one.get_lattice()
        ref_c = np.array(one.counters())
        ref_h = one.state_hash()
    out, counters, _ = _run_split(sn, lat, 8, sn.SN_KERNEL_TILED, 4, devices=devices, per_call=2)
    assert np.array_equal(out, ref)
    assert np.array_equal(counters, ref_c)
    assert _run_split.last_hash == ref_h


def test_slabs_run_many_sweeps_per_launch(sn, devices):
    """Several sweeps in one dataflow launch per slab: the two GPUs are ordered only by the tile versions
    they publish to each other over NVLink (no launch boundary, no barrier), and the chain is still the
    single-GPU one bit for bit."""
    X, Y, Z = 64, 64, 128
    lat = oa.random_lattice(X, Y, Z, seed=23, lengths=(1.0, 0.5, 0.0), prevalence=(0.8, 0.15, 0.05))
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, kernel=sn.SN_KERNEL_TILED_PHASED) as one:
        one.set_lattice(lat)
        one.MC_sweeps(6)
        ref = one.get_lattice()
        ref_c = np.array(one.counters())
        ref_e = one.total_energy(sn.SN_PREC_F64)
    out, counters, energy = _run_split(sn, lat, 2, sn.SN_KERNEL_TILED, 6, devices=devices[:2], per_call=3)
    assert np.array_equal(out, ref)
    assert np.array_equal(counters, ref_c)
    assert np.allclose(energy, ref_e, rtol=1e-12, atol=1e-9)


def test_two_large_slabs_per_device_share_the_sms(sn, devices):
    """Two slabs of one lattice on the SAME device, each with more tiles than SMs: their persistent kernels wait for
    each other's tile versions, so both must be resident -- sn_attach_peer halves their grids.  Still the 1-GPU chain."""
    X, Y, Z = 256, 256, 128
    lat = oa.random_lattice(X, Y, Z, seed=24)
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=99, kernel=sn.SN_KERNEL_TILED) as one:
        one.set_lattice(lat)
        one.MC_sweeps(2)
        ref = one.get_lattice()
    out, _, _ = _run_split(sn, lat, 4, sn.SN_KERNEL_TILED, 2, devices=devices[:2], per_call=2)
    assert np.array_equal(out, ref)


def test_a_slab_whose_neighbour_never_runs_fails_instead_of_hanging(sn, devices, monkeypatch):
    """Device-side waits are bounded (SN_SPIN_TIMEOUT_S): sweeping only one of two slabs must come back with an error
    from the next synchronising call, not hang the GPU."""
    monkeypatch.setenv("SN_SPIN_TIMEOUT_S", "1.5")
    X, Y, Z = 32, 32, 64
    lat = oa.random_lattice(X, Y, Z, seed=26)
    sims = [sn.Simulation(X, Y, Z, seed=1, device=devices[r % len(devices)], z0=32 * r, nz=32, kernel=sn.SN_KERNEL_TILED) for r in range(2)]
    try:
        for r, s in enumerate(sims):
            s.set_lattice(lat[:, :, 32 * r:32 * r + 32])
        for r, s in enumerate(sims):
            s.set_ghost(0, sims[1 - r].get_boundary(1)); s.set_ghost(1, sims[1 - r].get_boundary(0))
            s.attach_peer(0, sims[1 - r]); s.attach_peer(1, sims[1 - r])
        sims[0].MC_sweeps(1)                              # slab 1 never sweeps
        with pytest.raises(sn.SnError, match="timed out"):
            sims[0].synchronize()
    finally:
        for s in sims:
            s.close()
