"""GPU parity tests for site_energy / lattice energy (montecarlo-core.c:76-141),
called through the C ABI.  Bars (BASELINE.json north_star):
  SN_PREC_REPLICA  bit-equal to the native reference / oracle f32
  SN_PREC_F64      1e-12 relative, against the reference built with float->double
  SN_PREC_F32      1e-5 relative to sum_j |term_j| (the scale the native reference
                   itself is accurate to)
"""
import numpy as np
import pytest

from oracle import oracle_api as oa
from tests.helpers import CASE_NAMES, load_case, random_moves, sim_for, term_scale

pytestmark = pytest.mark.gpu

REL64 = 1e-12
REL32 = 1e-5


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


@pytest.mark.parametrize("name", CASE_NAMES)
def test_site_energy_against_golden(sn, name):
    g, p = load_case(name)
    with sim_for(sn, p, g["lattice"]) as sim:
        dxyz, d = sim.neighbours()
        assert np.array_equal(dxyz, g["nb_dxyz_f32"]) and np.array_equal(d.astype(np.float64), g["nb_d_f32"])
        rep = sim.site_energy(g["sites"], g["newdip"], sn.SN_PREC_REPLICA)
        assert np.array_equal(rep, g["dE_f32"]), "REPLICA mode must be bit-equal to the native reference"
        f64 = sim.site_energy(g["sites"], g["newdip"], sn.SN_PREC_F64)
        ref = g["dE_f64"]
        scale = term_scale(p, g["lattice"], g["sites"], g["newdip"])
        assert np.max(np.abs(f64 - ref) / np.maximum(np.abs(ref), 1e-6 * scale)) < REL64
        f32 = sim.site_energy(g["sites"], g["newdip"], sn.SN_PREC_F32)
        assert np.max(np.abs(f32 - ref) / scale) < REL32


@pytest.mark.parametrize("name", CASE_NAMES)
def test_total_energy_against_golden(sn, name):
    g, p = load_case(name)
    with sim_for(sn, p, g["lattice"]) as sim:
        e64 = sim.total_energy(sn.SN_PREC_F64)
        ref = g["total_f64"]
        assert np.allclose(e64, ref, rtol=REL64, atol=1e-12 * np.abs(ref).sum())
        erep = sim.total_energy(sn.SN_PREC_REPLICA)
        assert np.allclose(erep, g["total_f32"], rtol=1e-12, atol=1e-12 * np.abs(ref).sum())
        e32 = sim.total_energy(sn.SN_PREC_F32)
        n = p.X * p.Y * p.Z
        assert np.allclose(e32, ref, rtol=REL32, atol=REL32 * n * 1e-1)


@pytest.mark.parametrize("shape,cut", [((16, 16, 16), 3), ((20, 20, 28), 3), ((13, 17, 11), 3), ((100, 100, 1), 3),
                                        ((32, 32, 32), 3), ((12, 12, 12), 1), ((15, 14, 13), 5), ((3, 3, 3), 3),
                                        ((40, 8, 8), 6), ((5, 5, 1), 3)])
def test_site_energy_against_oracle_shapes(sn, shape, cut):
    """Seeded inputs at more shapes: odd extents (uneven colouring, ghost shell wrap), the
    `make test` geometry, the 2-D 100x100 case, extents below the cut-off, other cut-offs."""
    X, Y, Z = shape
    E = tuple(float(np.float32(v)) for v in (0.02, 0.015, -0.01))
    p = oa.make_params(X, Y, Z, cut, 1.1, 0.6, E, beta=1.0)
    lat = oa.random_lattice(X, Y, Z, seed=X * 31 + Z, lengths=(1.0, 0.4, 0.0), prevalence=(0.7, 0.2, 0.1))
    sites, nd = random_moves(p, 200, seed=3)
    with sim_for(sn, p, lat) as sim:
        assert np.array_equal(sim.site_energy(sites, nd, sn.SN_PREC_REPLICA), oa.Oracle("f32").site_energy(p, lat, sites, nd))
        ref = oa.Oracle("f64").site_energy(p, lat, sites, nd)
        scale = term_scale(p, lat, sites, nd)
        f64 = sim.site_energy(sites, nd, sn.SN_PREC_F64)
        assert np.max(np.abs(f64 - ref) / np.maximum(np.abs(ref), 1e-6 * scale)) < REL64
        f32 = sim.site_energy(sites, nd, sn.SN_PREC_F32)
        assert np.max(np.abs(f32 - ref) / scale) < REL32


def test_structured_lattices_known_answers(sn):
    """Closed-form initial states (lattice.c:39-105) as exact known answers."""
    g = dict(np.load(__import__("os").path.join(__import__("tests.helpers", fromlist=["GOLDEN"]).GOLDEN, "initial_lattices.npz")))
    p = oa.make_params(20, 20, 28, 3, 1.0, 0.0, (0, 0, 0), 1.0)
    sites, nd = random_moves(p, 128, seed=9)
    for kind in ("ferroelectric", "antiferro_wall", "ferro_wall", "antiferro_slip", "buckled"):
        lat = g[kind]
        with sim_for(sn, p, lat) as sim:
            assert np.array_equal(sim.site_energy(sites, nd, sn.SN_PREC_REPLICA), oa.Oracle("f32").site_energy(p, lat, sites, nd)), kind
            ref = oa.Oracle("f64").total_energy(p, lat)
            assert np.allclose(sim.total_energy(sn.SN_PREC_F64), ref, rtol=REL64, atol=1e-9), kind
    # ferroelectric along x: a fully aligned lattice has zero dipole sum over the cubic shells
    # inside the cut-off sphere (sum_j (1 - 3 n_x^2)/d^3 = 0 shell by shell), so E_dd = 0 and
    # E_cage = -1/2 * CageStrain * 6 * N
    with sim_for(sn, p, g["ferroelectric"]) as sim:
        e = sim.total_energy(sn.SN_PREC_F64)
        assert abs(e[0]) < 1e-9 * 11200
        assert e[1] == pytest.approx(-0.5 * 1.0 * 6 * 11200, rel=1e-14)


def test_energy_difference_is_site_energy(sn):
    """H is defined so that site_energy is its exact single-site difference."""
    E = tuple(float(np.float32(v)) for v in (0.05, 0.0, -0.02))
    p = oa.make_params(12, 12, 12, 3, 1.5, 0.8, E, beta=1.0)
    lat = oa.random_lattice(12, 12, 12, seed=2, lengths=(1.0, 0.5), prevalence=(0.7, 0.3))
    sites, nd = random_moves(p, 8, seed=1)
    with sim_for(sn, p, lat) as sim:
        dE = sim.site_energy(sites, nd, sn.SN_PREC_F64)
        e0 = sim.total_energy(sn.SN_PREC_F64).sum()
        for i in range(len(sites)):
            trial = lat.copy()
            x, y, z = sites[i]
            trial[x, y, z, :3] = nd[i]
            sim.set_lattice(trial)
            e1 = sim.total_energy(sn.SN_PREC_F64).sum()
            assert e1 - e0 == pytest.approx(dE[i], abs=1e-9)
        sim.set_lattice(lat)


def test_replicas_and_couplings(sn):
    """Per-replica beta / Efield and the global CageStrain reach the kernels."""
    p = oa.make_params(12, 12, 12, 3, 1.0, 0.0, (0, 0, 0), 1.0)
    lats = [oa.random_lattice(12, 12, 12, seed=s) for s in range(3)]
    sites, nd = random_moves(p, 64, seed=5)
    with sim_for(sn, p, nreplicas=3) as sim:
        for r, lat in enumerate(lats):
            sim.set_lattice(lat, r)
        sim.set_efield((0.25, 0.0, 0.5), replica=1)
        sim.set_cagestrain(2.0)
        for r, lat in enumerate(lats):
            q = oa.make_params(12, 12, 12, 3, 2.0, 0.0, (0.25, 0.0, 0.5) if r == 1 else (0, 0, 0), 1.0)
            assert np.array_equal(sim.site_energy(sites, nd, sn.SN_PREC_REPLICA, replica=r), oa.Oracle("f32").site_energy(q, lat, sites, nd))


def test_errors_are_reported(sn):
    with pytest.raises(sn.SnError):
        sn.Simulation(0, 4, 4)
    with pytest.raises(sn.SnError):
        sn.Simulation(8, 8, 8, DipoleCutOff=9)
    with sn.Simulation(8, 8, 8) as sim:
        with pytest.raises(sn.SnError):
            sim.site_energy([[9, 0, 0]], [[1, 0, 0]])
        with pytest.raises(sn.SnError):
            sim.set_lattice(np.zeros((4, 4, 4, 4), np.float32))
        with pytest.raises(sn.SnError):
            sim.get_lattice(replica=3)
