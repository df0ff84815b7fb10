"""GPU parity tests for the observables of starrynight-analysis.c, through the C ABI."""
import numpy as np
import pytest

from oracle import oracle_api as oa
from tests.helpers import CASE_NAMES, load_case, sim_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


@pytest.mark.parametrize("name", CASE_NAMES)
def test_observables_against_golden(sn, name):
    g, p = load_case(name)
    lat = g["lattice"]
    n = p.X * p.Y * p.Z
    with sim_for(sn, p, lat) as sim:
        P = sim.polarisation()
        assert P[0] == pytest.approx(float(g["polarisation_f64"]), rel=1e-12, abs=1e-15)       # analysis.c:48-62
        assert np.allclose(P, lat[..., :3].astype(np.float64).reshape(-1, 3).sum(0) / n, rtol=1e-12, atol=1e-15)
        assert sim.landau_order() == pytest.approx(float(g["landau_f64"]), rel=1e-12)           # analysis.c:506-526
        assert sim.landau_order() == pytest.approx(float(g["landau_f32"]), rel=2e-5)            # float accumulators there
        if "potential_f64" in g:
            V = sim.dipole_potential().ravel()                                                   # analysis.c:65-94
            ref = g["potential_f64"]
            assert np.max(np.abs(V - ref)) < 1e-12 * np.max(np.abs(ref)) * 50
            assert np.max(np.abs(V - g["potential_f32"])) < 2e-5 * np.max(np.abs(ref))
        if "rdf_f64" in g:
            fe, afe, cnt = sim.radial_order_parameter()                                          # analysis.c:528-598
            ofe, oafe, ocnt = oa.Oracle("f64").rdf(p, lat)
            assert np.array_equal(cnt, ocnt.astype(np.int64))
            assert np.allclose(fe, ofe, rtol=1e-11, atol=1e-9)
            assert np.allclose(afe, oafe, rtol=1e-11, atol=1e-9)
            rows = g["rdf_f64"]                     # what the reference printed: r2 r FE AFE count T
            nz = np.nonzero(cnt)[0]
            assert np.array_equal(rows[:, 0].astype(int), nz)
            assert np.allclose(rows[:, 2], fe[nz] / cnt[nz], atol=6e-7)
            assert np.allclose(rows[:, 3], afe[nz] / cnt[nz], atol=6e-7)


def test_flat_lattice_potential_counts_the_plane_13_times(sn):
    """dipole_potential always loops dz in [-6,6]; on Z == 1 every dz lands on the same plane
    through %Z (analysis.c:73-91).  The kernel keeps that behaviour."""
    p = oa.make_params(14, 12, 1, 3, 1.0, 0.0, (0, 0, 0), 1.0)
    lat = oa.random_lattice(14, 12, 1, seed=8)
    with sim_for(sn, p, lat) as sim:
        V = sim.dipole_potential().ravel()
    ref = oa.Oracle("f64").potential_map(p, lat)
    assert np.max(np.abs(V - ref)) < 1e-11


def test_large_lattice_properties(sn):
    """Size-independent properties at a size the CPU oracle cannot finish: ferroelectric state."""
    X = 96
    lat = np.zeros((X, X, X, 4), np.float32)
    lat[..., 0] = 1.0
    lat[..., 3] = 1.0
    with sn.Simulation(X, X, X, CageStrain=1.0, K=0.0) as sim:
        sim.set_lattice(lat)
        assert np.allclose(sim.polarisation(), [1.0, 0.0, 0.0])
        fe, afe, cnt = sim.radial_order_parameter()
        nz = cnt > 0
        assert np.array_equal(fe[nz], cnt[nz].astype(np.float64))        # FE correlation exactly 1 at every r
        assert cnt[1] == 6 * X ** 3 and cnt[74] == 120 * X ** 3          # int64 counts: beyond int32 at 512^3
        e = sim.total_energy(sn.SN_PREC_F64)
        assert e[1] == pytest.approx(-3.0 * X ** 3, rel=1e-13)


@pytest.mark.parametrize("name", ["species3d", "flat2d", "odd_cut2", "cut4_constrain", "dim2"])
def test_efield_maps_and_recombination_against_golden(sn, name):
    """sn_efield_map / sn_recombination (FP64 kernels) against the reference build's vectors:
    1e-12 of the float->double build, float accuracy of the native one."""
    from tests.test_oracle import parse_recombination_log
    from tests.helpers import GOLDEN
    g, p = load_case(name)
    x = dict(np.load(f"{GOLDEN}/analysis_extra.npz"))
    lat = g["lattice"]
    with sim_for(sn, p, lat) as sim:
        for key, cut, half in (("efield", 4, False), ("efieldoffset", 2, True)):
            E = sim.dipole_electricfield(cut, half).ravel()                  # analysis.c:310-465
            ref = x[f"{name}_{key}_f64"]
            assert np.max(np.abs(E - ref)) < 1e-12 * np.max(np.abs(ref)) * 50
            assert np.max(np.abs(E - x[f"{name}_{key}_f32"])) < 2e-5 * np.max(np.abs(ref))
        got = sim.recombination()                                            # analysis.c:96-171
    ref = parse_recombination_log(bytes(x[f"{name}_recombination_f64"]).decode())
    assert np.allclose(got[:8], ref, rtol=1.5e-6)                            # the log holds 7 digits
    full = oa.Oracle("f64").recombination(p, lat)
    assert np.allclose(got, full, rtol=1e-11)


@pytest.fixture(scope="module")
def devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n >= 2 else [0]


@pytest.mark.parametrize("shape,nslab,kernel", [((64, 32, 64), 2, "tiled"), ((20, 16, 36), 3, "colour"), ((32, 32, 128), 4, "tiled")])
def test_slab_native_observables_equal_the_whole_lattice(sn, devices, shape, nslab, kernel):
    """north_star parts (2)+(3): radial_order_parameter (analysis.c:528-598), dipole_potential (:65-94), the E-field maps
    (:310-465) and the recombination sums (:96-170) evaluated on a Z-slab decomposed lattice -- each slab over its own
    sites, planes beyond the slab read from the neighbouring slabs' device memory -- against one handle holding the
    whole lattice, after sweeps (so the slabs' data is what the sweep kernels and their halo pushes left behind)."""
    X, Y, Z = shape
    kid = sn.SN_KERNEL_TILED if kernel == "tiled" else sn.SN_KERNEL_COLOUR
    lat = oa.random_lattice(X, Y, Z, seed=33, lengths=(1.0, 0.5, 0.0), prevalence=(0.7, 0.2, 0.1))
    with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=9, kernel=kid) as one:
        one.set_lattice(lat)
        one.MC_sweeps(2)
        fe, afe, cnt = one.radial_order_parameter()
        V = one.dipole_potential()
        E4 = one.dipole_electricfield(4, False)
        E2 = one.dipole_electricfield(2, True)
        rec = one.recombination()
        P = one.polarisation()
    nz = Z // nslab
    sims = [sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.05, 0, 0), seed=9, device=devices[r % len(devices)], z0=r * nz, nz=nz, kernel=kid)
            for r in range(nslab)]
    try:
        for r, s in enumerate(sims):
            s.set_lattice(lat[:, :, r * nz:(r + 1) * nz])
        for r, s in enumerate(sims):
            s.attach_peer(0, sims[(r - 1) % nslab]); s.attach_peer(1, sims[(r + 1) % nslab])
        for s in sims:
            s.pull_ghosts()
        for s in sims:
            s.MC_sweeps(2)
        parts = [s.radial_order_parameter() for s in sims]
        assert np.array_equal(sum(q[2] for q in parts), cnt)
        assert np.allclose(sum(q[0] for q in parts), fe, rtol=1e-12, atol=1e-9)
        assert np.allclose(sum(q[1] for q in parts), afe, rtol=1e-12, atol=1e-9)
        assert np.array_equal(np.concatenate([s.dipole_potential() for s in sims], axis=2), V)      # same terms in the same order
        assert np.array_equal(np.concatenate([s.dipole_electricfield(4, False) for s in sims], axis=2), E4)
        assert np.array_equal(np.concatenate([s.dipole_electricfield(2, True) for s in sims], axis=2), E2)
        got = sn.recombination_finish([s.recombination_partial() for s in sims])
        assert np.allclose(got, rec, rtol=1e-12)
        assert np.allclose(sum(np.asarray(s.polarisation()) * s.nsites for s in sims) / (X * Y * Z), P, rtol=1e-12, atol=1e-15)
        with pytest.raises(sn.SnError, match="sn_recombination_partial"):
            sims[0].recombination()
    finally:
        for s in sims:
            s.close()


def test_observable_kernels_cover_ragged_and_tiny_lattices(sn):
    """Extents that are not multiples of the 8^3 observable tile, down to the stencil radius, and a flat lattice (every dz
    lands on the one plane): against the f64 oracle.  (Below the radius the reference indexes out of bounds --
    `(X + x + dx) % X` is negative in C for dx < -X - x, analysis.c:88,566 -- so there is nothing to be equal to; the
    kernels keep wrapping periodically.)"""
    for (X, Y, Z) in [(13, 9, 11), (10, 17, 9), (17, 9, 1)]:
        p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, (0, 0, 0), 1.0)
        lat = oa.random_lattice(X, Y, Z, seed=X, lengths=(1.0, 0.5), prevalence=(0.7, 0.3))
        o = oa.Oracle("f64")
        with sim_for(sn, p, lat) as sim:
            V = sim.dipole_potential().ravel()
            ref = o.potential_map(p, lat)
            assert np.max(np.abs(V - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
            fe, afe, cnt = sim.radial_order_parameter()
            ofe, oafe, ocnt = o.rdf(p, lat)
            assert np.array_equal(cnt, ocnt.astype(np.int64))
            assert np.allclose(fe, ofe, rtol=1e-11, atol=1e-9) and np.allclose(afe, oafe, rtol=1e-11, atol=1e-9)
            E = sim.dipole_electricfield(4, False).ravel()
            assert np.allclose(E, o.efield_map(p, lat, 4, False).ravel(), rtol=1e-11, atol=1e-12)
