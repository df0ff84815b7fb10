"""GPU tests of the Metropolis sweep (MC_moves / MC_move, montecarlo-core.c:143-191).

The colour-sweep order and the Philox stream necessarily give a different Markov
chain from the reference's serial random-site MT19937 chain, so parity is
(i) structural invariants, (ii) exactness of the energies the chain is built on
(test_gpu_energy.py), and (iii) agreement of equilibrium observables with the
reference chain (the CPU oracle, bit-equal to the reference) within statistical
error bars from independent seeds."""
import numpy as np
import pytest

from oracle import oracle_api as oa
from tests.helpers import sim_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


KERNELS = ["colour", "auto"]


def kernel_id(sn, k):
    return {"colour": sn.SN_KERNEL_COLOUR, "auto": sn.SN_KERNEL_AUTO}[k]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape", [(32, 32, 32), (20, 20, 28), (13, 17, 11), (24, 24, 1), (64, 32, 32)])
def test_sweep_invariants(sn, shape, kernel):
    X, Y, Z = shape
    lat = oa.random_lattice(X, Y, Z, seed=5, lengths=(1.0, 0.5, 0.0), prevalence=(0.6, 0.3, 0.1))
    p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, (0.02, 0, 0), 1.0)
    with sim_for(sn, p, lat, kernel=kernel_id(sn, kernel)) as sim:
        sim.MC_sweeps(3)
        acc, rej, vac = sim.counters()
        out = sim.get_lattice()
    n = X * Y * Z
    nvac = int((lat[..., 3] == 0).sum())
    assert vac == 3 * nvac                          # vacancies are skipped, never counted (montecarlo-core.c:163)
    assert acc + rej == 3 * (n - nvac)
    assert 0.05 < acc / (acc + rej) < 0.95
    assert np.array_equal(out[..., 3], lat[..., 3])                    # length / species never changes (:173,:184)
    assert np.array_equal(out[lat[..., 3] == 0], lat[lat[..., 3] == 0])
    live = lat[..., 3] != 0
    assert np.max(np.abs(np.linalg.norm(out[..., :3][live], axis=-1) - 1.0)) < 2e-6
    assert (out[..., :3][live] != lat[..., :3][live]).any(axis=-1).mean() > 0.3


@pytest.mark.parametrize("kernel", KERNELS)
def test_sweep_is_deterministic_and_seed_dependent(sn, kernel):
    lat = oa.random_lattice(32, 32, 32, seed=6)
    outs = []
    for seed in (1, 1, 2):
        with sn.Simulation(32, 32, 32, seed=seed, kernel=kernel_id(sn, kernel)) as sim:
            sim.set_lattice(lat)
            sim.MC_sweeps(2)
            sim.MC_sweeps(1)
            outs.append(sim.get_lattice())
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("kernel", KERNELS)
def test_ghost_shell_stays_consistent(sn, kernel):
    """After sweeps the energy computed through the ghost shell must equal the energy of the
    downloaded lattice re-uploaded from scratch (ghosts rebuilt): every boundary update wrote its images."""
    lat = oa.random_lattice(32, 32, 32, seed=7)
    with sn.Simulation(32, 32, 32, CageStrain=1.0, kernel=kernel_id(sn, kernel)) as sim:
        sim.set_lattice(lat)
        sim.MC_sweeps(5)
        e_live = sim.total_energy(sn.SN_PREC_F64)
        out = sim.get_lattice()
        sim.set_lattice(out)
        e_fresh = sim.total_energy(sn.SN_PREC_F64)
    assert np.array_equal(e_live, e_fresh)
    p = oa.make_params(32, 32, 32, 3, 1.0, 0.0, (0, 0, 0), 1.0)
    assert np.allclose(e_live, oa.Oracle("f64").total_energy(p, out), rtol=1e-11, atol=1e-8)


def test_zero_temperature_only_goes_downhill(sn):
    """T = 0 => beta = +inf: accept iff dE < 0 (main.c:215, montecarlo-core.c:179)."""
    lat = oa.random_lattice(16, 16, 16, seed=8)
    with sn.Simulation(16, 16, 16, CageStrain=1.0, beta=float("inf")) as sim:
        sim.set_lattice(lat)
        e = [sim.total_energy(sn.SN_PREC_F64).sum()]
        for _ in range(6):
            sim.MC_sweeps(1)
            e.append(sim.total_energy(sn.SN_PREC_F64).sum())
    assert all(b <= a + 1e-6 for a, b in zip(e, e[1:]))
    assert e[-1] < e[0] - 100


def test_constrain_to_x_proposals(sn):
    """ConstrainToX: every accepted orientation is one of the six <100> vectors (config.c:230-263)."""
    lat = oa.random_lattice(16, 16, 16, seed=9)
    with sn.Simulation(16, 16, 16, ConstrainToX=True, beta=0.2) as sim:
        sim.set_lattice(lat)
        sim.MC_sweeps(30)
        out = sim.get_lattice()[..., :3].reshape(-1, 3)
    moved = (out != lat[..., :3].reshape(-1, 3)).any(1)
    assert moved.mean() > 0.9
    v = out[moved]
    assert np.all(np.sort(np.abs(v), axis=1) == np.array([0, 0, 1], np.float32))
    frac = [(np.argmax(np.abs(v), 1) == k).mean() for k in range(3)]
    assert all(abs(f - 1 / 3) < 0.05 for f in frac)


def test_dim2_proposals_stay_in_plane(sn):
    lat = oa.random_lattice(16, 16, 16, seed=10)
    with sn.Simulation(16, 16, 16, DIM=2, beta=0.2) as sim:
        sim.set_lattice(lat)
        sim.MC_sweeps(30)
        out = sim.get_lattice()[..., :3].reshape(-1, 3)
    moved = (out != lat[..., :3].reshape(-1, 3)).any(1)
    assert np.all(out[moved][:, 2] == 0)


def test_infinite_temperature_proposals_are_uniform_on_the_sphere(sn):
    """beta = 0 accepts everything: the lattice becomes a sample of the proposal distribution."""
    lat = oa.random_lattice(32, 32, 32, seed=11)
    with sn.Simulation(32, 32, 32, beta=0.0) as sim:
        sim.set_lattice(lat)
        sim.MC_sweeps(1)
        acc, rej, vac = sim.counters()
        v = sim.get_lattice()[..., :3].reshape(-1, 3).astype(np.float64)
    assert rej == 0 and acc == 32 ** 3
    n = len(v)
    assert np.all(np.abs(v.mean(0)) < 5 / np.sqrt(3 * n))                   # <p> = 0, var 1/3 per component
    assert np.all(np.abs((v ** 2).mean(0) - 1 / 3) < 5 * np.sqrt(4 / 45 / n))
    hist, _ = np.histogram(v[:, 2], bins=20, range=(-1, 1))                 # z uniform on [-1,1] (Archimedes)
    assert np.all(np.abs(hist - n / 20) < 6 * np.sqrt(n / 20))


def _reference_chain_stats(p, lat0, seeds, eqm, nsamp, stride):
    """<E>/N, <P_x>, acceptance from the reference's serial chain (oracle f32, bit-equal to the reference)."""
    o = oa.Oracle("f32")
    n = p.X * p.Y * p.Z
    rows = []
    for s in seeds:
        lat = np.ascontiguousarray(lat0, np.float32)
        mt = o.mt(s)
        o.mc_moves(p, lat, mt, eqm * n)
        es, ps, acc, rej = [], [], 0, 0
        for _ in range(nsamp):
            a, r = o.mc_moves(p, lat, mt, stride * n)
            acc += a; rej += r
            es.append(o.total_energy(p, lat).sum() / n)
            ps.append(o.polarisation(p, lat))
        rows.append((np.mean(es), np.mean(ps), acc / (acc + rej)))
    return np.array(rows)


@pytest.mark.parametrize("T,Ex", [(150, 0.3), (300, 0.3), (600, 0.5)])
def test_equilibrium_observables_match_reference_chain(sn, T, Ex):
    """Energy per site, polarisation along an applied field and acceptance ratio at equilibrium:
    colour-sweep Philox chain on the GPU vs the reference's serial chain, independent seeds on both
    sides, agreement within 4.5 combined standard errors."""
    X = 10
    beta = 1.0 / (T / 300.0)
    p = oa.make_params(X, X, X, 3, 1.0, 0.0, (float(np.float32(Ex)), 0.0, 0.0), beta)
    lat0 = oa.random_lattice(X, X, X, seed=12)
    eqm, nsamp, stride = 120, 20, 6
    ref = _reference_chain_stats(p, lat0, seeds=range(100, 106), eqm=eqm, nsamp=nsamp, stride=stride)
    R = 16
    with sim_for(sn, p, nreplicas=R, seed=4242) as sim:
        for r in range(R):
            sim.set_lattice(lat0, r)
        sim.MC_sweeps(eqm)
        sim.reset_counters()
        es, ps = np.zeros((R, nsamp)), np.zeros((R, nsamp))
        for k in range(nsamp):
            sim.MC_sweeps(stride)
            for r in range(R):
                es[r, k] = sim.total_energy(sn.SN_PREC_F32, r).sum() / X ** 3
                ps[r, k] = sim.polarisation(r)[0]
        ratios = []
        for r in range(R):
            a, rj, _ = sim.counters(r)
            ratios.append(a / (a + rj))
    gpu = np.stack([es.mean(1), ps.mean(1), np.array(ratios)], 1)
    for col, what in enumerate(("energy per site", "polarisation", "acceptance ratio")):
        m_ref, m_gpu = ref[:, col].mean(), gpu[:, col].mean()
        se = np.sqrt(ref[:, col].var(ddof=1) / len(ref) + gpu[:, col].var(ddof=1) / len(gpu))
        assert abs(m_ref - m_gpu) < 4.5 * se + 2e-4, f"T={T}: {what} reference {m_ref:.5f} vs GPU {m_gpu:.5f} (se {se:.5f})"


@pytest.mark.parametrize("shape,reps,calls", [((32, 32, 32), 1, (3,)), ((64, 64, 64), 1, (2, 3)), ((96, 64, 32), 2, (4,)),
                                              ((160, 96, 64), 1, (3, 1)), ((256, 256, 64), 1, (2,)),
                                              ((48, 48, 48), 2, (2, 1)), ((80, 48, 32), 1, (3,)), ((112, 32, 144), 1, (2,)),
                                              ((100, 44, 36), 1, (2, 1)), ((33, 50, 40), 2, (3,)), ((20, 21, 28), 1, (3,)), ((64, 30, 20), 2, (2,))])
def test_dataflow_launch_equals_barrier_separated_phases(sn, shape, reps, calls):
    """The tiled kernel runs whole sweeps in one launch, ordering adjacent tiles through per-tile version
    counters instead of a barrier per tile-parity phase.  The chain must be bit-identical to the same
    kernel launched once per phase (stream order = barrier), for any number of sweeps per call."""
    X, Y, Z = shape
    lats = [oa.random_lattice(X, Y, Z, seed=30 + r, lengths=(1.0, 0.5, 0.0), prevalence=(0.7, 0.2, 0.1)) for r in range(reps)]
    res = []
    for kern in (sn.SN_KERNEL_TILED, sn.SN_KERNEL_TILED_PHASED):
        with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.03, 0, 0), nreplicas=reps, seed=77, kernel=kern) as sim:
            for r in range(reps):
                sim.set_lattice(lats[r], r)
            for c in calls:
                sim.MC_sweeps(c)
            res.append(([sim.get_lattice(r) for r in range(reps)], [sim.counters(r) for r in range(reps)]))
    for r in range(reps):
        assert np.array_equal(res[0][0][r], res[1][0][r]), f"replica {r}: dataflow chain differs from the phased chain"
        assert res[0][1][r] == res[1][1][r]
        assert not np.array_equal(res[0][0][r][..., :3], lats[r][..., :3])


@pytest.mark.parametrize("shape", [(20, 20, 28), (32, 32, 64), (24, 24, 1)])
def test_seeded_replicas_reproduce_independent_runs(sn, shape):
    """sn_set_replica_seed: a batch of replicas (a temperature sweep) draws exactly the numbers of as many
    separate handles -- the reference's parallel mode is one process per T seeded 0xDEADBEEF + T
    (main.c:172, Makefile:49-63)."""
    X, Y, Z = shape
    temps = [75, 300, 450]
    lats = [oa.random_lattice(X, Y, Z, seed=40 + r, lengths=(1.0, 0.5), prevalence=(0.8, 0.2)) for r in range(3)]
    with sn.Simulation(X, Y, Z, CageStrain=1.0, nreplicas=3, seed=0xDEADBEEF + temps[0]) as batch:
        for r, T in enumerate(temps):
            batch.set_lattice(lats[r], r)
            batch.set_T(T, r)
            batch.set_replica_seed(0xDEADBEEF + T, r)
        batch.MC_sweeps(2)
        batch.MC_sweeps(3)
        got = [(batch.get_lattice(r), batch.counters(r)) for r in range(3)]
    for r, T in enumerate(temps):
        with sn.Simulation(X, Y, Z, CageStrain=1.0, seed=0xDEADBEEF + T) as one:
            one.set_lattice(lats[r])
            one.set_T(T)
            one.MC_sweeps(5)
            assert np.array_equal(one.get_lattice(), got[r][0]), f"replica {r} (T={T}) differs from its separate run"
            assert one.counters() == got[r][1]


RESIDENT_CASES = [
    # X, Y, Z, cutoff, K, ConstrainToX, DIM, replicas
    (20, 20, 28, 3, 0.0, False, 3, 2),          # the reference's `make test` lattice
    (100, 100, 1, 3, 0.0, False, 3, 3),         # the 2-D figures case
    (13, 17, 11, 3, 0.6, False, 3, 1),          # odd extents: extra colours on every axis
    (9, 11, 10, 2, 0.0, False, 3, 2),           # table-driven cut-offs
    (9, 10, 9, 4, 0.3, True, 3, 1),
    (12, 12, 12, 3, 0.0, False, 2, 1),
    (24, 24, 1, 3, 0.0, False, 3, 1),
    (8, 8, 8, 3, 0.0, False, 3, 200),           # more replicas than SMs: CTAs loop over replicas
    (4, 5, 3, 3, 0.0, False, 3, 2),             # extents barely above the cut-off
]


@pytest.mark.parametrize("case", RESIDENT_CASES)
def test_resident_kernel_equals_colour_passes_bit_for_bit(sn, case):
    """The shared-memory-resident kernel runs the same colour order, Philox streams and arithmetic as one
    launch per colour over global memory: identical lattices and counters, any sweeps per call."""
    X, Y, Z, cut, K, constrain, dim, reps = case
    lats = [oa.random_lattice(X, Y, Z, seed=60 + r % 5, lengths=(1.0, 0.5, 0.0), prevalence=(0.7, 0.2, 0.1)) for r in range(reps)]
    res = []
    for kern in (sn.SN_KERNEL_RESIDENT, sn.SN_KERNEL_COLOUR):
        with sn.Simulation(X, Y, Z, DipoleCutOff=cut, CageStrain=1.0, K=K, Efield=(0.02, 0.01, 0), ConstrainToX=constrain, DIM=dim,
                           nreplicas=reps, seed=31, kernel=kern) as sim:
            for r in range(reps):
                sim.set_lattice(lats[r], r)
                sim.set_T(100 + 37 * (r % 11), r)
            sim.MC_sweeps(3)
            sim.MC_sweeps(1)
            e = sim.total_energy(sn.SN_PREC_F64, reps - 1)          # through the ghost shell the kernel left behind
            res.append(([sim.get_lattice(r) for r in range(reps)], [sim.counters(r) for r in range(reps)], e))
    for r in range(reps):
        assert np.array_equal(res[0][0][r], res[1][0][r]), f"replica {r}"
        assert res[0][1][r] == res[1][1][r]
    assert np.array_equal(res[0][2], res[1][2])
    assert not np.array_equal(res[0][0][0][..., :3], lats[0][..., :3])


@pytest.mark.parametrize("shape,want", [((48, 48, 48), "tiled"), ((80, 96, 112), "tiled"), ((100, 100, 100), "tiled"), ((45, 37, 36), "tiled"),
                                        ((100, 100, 98), "colour"), ((32, 32, 16), "colour"), ((64, 64, 28), "tiled"), ((30, 30, 32), "tiled"),
                                        ((64, 64, 64), "tiled"), ((20, 20, 28), "resident"), ((100, 100, 1), "resident")])
def test_kernel_selection(sn, shape, want):
    """SN_KERNEL_AUTO: the tiled kernel takes every cut-off-3 lattice with X, Y, Z >= 20 and Z a multiple of 4 that does not fit the resident kernel -- the last tile
    of an axis may be partial, an odd number of tiles along an axis gets a third tile colour -- small lattices live in shared
    memory, the rest runs colour passes."""
    ids = {"tiled": sn.SN_KERNEL_TILED, "colour": sn.SN_KERNEL_COLOUR, "resident": sn.SN_KERNEL_RESIDENT}
    with sn.Simulation(*shape) as sim:
        assert sim.kernel_in_use() == ids[want]


def test_kernel_selection_cutoff_2_runs_on_the_tiles_and_4_on_colour_passes(sn):
    with sn.Simulation(64, 64, 64, DipoleCutOff=2) as sim:
        assert sim.kernel_in_use() == sn.SN_KERNEL_TILED
    with sn.Simulation(64, 64, 64, DipoleCutOff=4) as sim:
        assert sim.kernel_in_use() == sn.SN_KERNEL_COLOUR


def test_odd_tile_counts_match_the_colour_kernel_statistics(sn):
    """48^3 (3 tiles per axis, 27 tile phases) on the tiled kernel against colour passes: equilibrium energy and
    acceptance agree within error bars (independent seeds)."""
    X, R = 48, 6
    lat0 = oa.random_lattice(X, X, X, seed=72)
    stats = {}
    for name, kern in (("tiled", sn.SN_KERNEL_TILED), ("colour", sn.SN_KERNEL_COLOUR)):
        with sn.Simulation(X, X, X, CageStrain=1.0, Efield=(0.3, 0, 0), beta=1.0, nreplicas=R, seed=950 + len(name), kernel=kern) as sim:
            for r in range(R):
                sim.set_lattice(lat0, r)
            sim.MC_sweeps(60)
            sim.reset_counters()
            es = np.zeros((R, 10))
            for k in range(10):
                sim.MC_sweeps(4)
                for r in range(R):
                    es[r, k] = sim.total_energy(sn.SN_PREC_F32, r).sum() / X ** 3
            acc = np.array([sim.counters(r)[0] / sum(sim.counters(r)[:2]) for r in range(R)])
        stats[name] = np.stack([es.mean(1), acc], 1)
    for col, what in enumerate(("energy per site", "acceptance ratio")):
        a, b = stats["tiled"][:, col], stats["colour"][:, col]
        se = np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R)
        assert abs(a.mean() - b.mean()) < 4.5 * se + 1e-4, f"{what}: tiled {a.mean():.5f} vs colour {b.mean():.5f} (se {se:.5f})"


def test_resident_kernel_refuses_what_it_cannot_hold(sn):
    with pytest.raises(sn.SnError, match="shared memory"):
        sn.Simulation(32, 32, 32, kernel=sn.SN_KERNEL_RESIDENT)


def test_tiled_kernel_statistics_match_colour_passes(sn):
    """The tiled dataflow kernel and the colour-pass kernel are different update orders of the same Markov
    kernel with the same stationary distribution: equilibrium energy, field-induced polarisation and
    acceptance agree within error bars (8 independent replicas each, 32^3, T = 300 K, E_x = 0.3)."""
    X, R = 32, 8
    lat0 = oa.random_lattice(X, X, X, seed=71)
    stats = {}
    for name, kern in (("tiled", sn.SN_KERNEL_TILED), ("colour", sn.SN_KERNEL_COLOUR)):
        with sn.Simulation(X, X, X, CageStrain=1.0, Efield=(0.3, 0, 0), beta=1.0, nreplicas=R, seed=900 + len(name), kernel=kern) as sim:
            for r in range(R):
                sim.set_lattice(lat0, r)
            sim.MC_sweeps(80)
            sim.reset_counters()
            es, ps = np.zeros((R, 25)), np.zeros((R, 25))
            for k in range(25):
                sim.MC_sweeps(4)
                for r in range(R):
                    es[r, k] = sim.total_energy(sn.SN_PREC_F32, r).sum() / X ** 3
                    ps[r, k] = sim.polarisation(r)[0]
            acc = np.array([sim.counters(r)[0] / sum(sim.counters(r)[:2]) for r in range(R)])
        stats[name] = np.stack([es.mean(1), ps.mean(1), acc], 1)
    for col, what in enumerate(("energy per site", "polarisation", "acceptance ratio")):
        a, b = stats["tiled"][:, col], stats["colour"][:, col]
        se = np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R)
        assert abs(a.mean() - b.mean()) < 4.5 * se + 1e-4, f"{what}: tiled {a.mean():.5f} vs colour {b.mean():.5f} (se {se:.5f})"


def test_correlation_functions_match_reference_chain(sn):
    """radial_order_parameter at equilibrium (T = 150 K): the nearest shells' FE and AFE correlations from the
    GPU chain vs the reference's serial chain (oracle f32), independent seeds, 4.5 combined standard errors."""
    X, T = 10, 150
    p = oa.make_params(X, X, X, 3, 1.0, 0.0, (0.0, 0.0, 0.0), 1.0 / (T / 300.0))
    lat0 = oa.random_lattice(X, X, X, seed=13)
    o = oa.Oracle("f32")
    n, eqm, nsamp, stride, shells = X ** 3, 100, 12, 5, [1, 2, 3, 4]
    ref = []
    for s in range(200, 205):
        lat = np.ascontiguousarray(lat0, np.float32)
        mt = o.mt(s)
        o.mc_moves(p, lat, mt, eqm * n)
        acc_fe, acc_afe = np.zeros(4), np.zeros(4)
        for _ in range(nsamp):
            o.mc_moves(p, lat, mt, stride * n)
            fe, afe, cnt = o.rdf(p, lat)
            acc_fe += np.asarray(fe, np.float64)[shells] / cnt[shells]
            acc_afe += np.asarray(afe, np.float64)[shells] / cnt[shells]
        ref.append(np.concatenate([acc_fe, acc_afe]) / nsamp)
    ref = np.array(ref)
    R = 12
    with sim_for(sn, p, nreplicas=R, seed=5151) as sim:
        for r in range(R):
            sim.set_lattice(lat0, r)
        sim.MC_sweeps(eqm)
        gpu = np.zeros((R, 8))
        for _ in range(nsamp):
            sim.MC_sweeps(stride)
            for r in range(R):
                fe, afe, cnt = sim.radial_order_parameter(r)
                gpu[r, :4] += fe[shells] / cnt[shells]
                gpu[r, 4:] += afe[shells] / cnt[shells]
        gpu /= nsamp
    for c in range(8):
        se = np.sqrt(ref[:, c].var(ddof=1) / len(ref) + gpu[:, c].var(ddof=1) / R)
        what = ("FE", "AFE")[c // 4] + f" correlation at r^2={shells[c % 4]}"
        assert abs(ref[:, c].mean() - gpu[:, c].mean()) < 4.5 * se + 2e-4, f"{what}: reference {ref[:, c].mean():.5f} vs GPU {gpu[:, c].mean():.5f} (se {se:.5f})"


def test_hysteresis_loop_matches_reference_chain(sn):
    """The field protocol main.c:229-238 sketches (commented out there; `Hysteresis` in our driver): a triangular
    ramp of Efield.x with a fixed number of sweeps per field point.  Polarisation at every point of the loop,
    GPU replicas vs the reference's serial chain driven through the same protocol, within error bars."""
    X, T, A, steps, spp = 8, 200, 0.6, 4, 6
    beta = 1.0 / (T / 300.0)
    fields = [A * (ph if ph < 1 else 2 - ph if ph < 3 else ph - 4) for ph in (s / steps for s in range(4 * steps))]
    lat0 = oa.random_lattice(X, X, X, seed=14)
    o = oa.Oracle("f32")
    n = X ** 3
    ref = []
    for s in range(300, 310):
        lat = np.ascontiguousarray(lat0, np.float32)
        mt = o.mt(s)
        o.mc_moves(oa.make_params(X, X, X, 3, 1.0, 0.0, (0.0, 0.0, 0.0), beta), lat, mt, 40 * n)
        row = []
        for e in fields:
            p = oa.make_params(X, X, X, 3, 1.0, 0.0, (float(np.float32(e)), 0.0, 0.0), beta)
            o.mc_moves(p, lat, mt, spp * n)
            row.append(o.polarisation(p, lat))
        ref.append(row)
    ref = np.array(ref)
    R = 24
    with sn.Simulation(X, X, X, CageStrain=1.0, beta=beta, nreplicas=R, seed=6161) as sim:
        for r in range(R):
            sim.set_lattice(lat0, r)
        sim.MC_sweeps(40)
        gpu = np.zeros((R, len(fields)))
        for k, e in enumerate(fields):
            for r in range(R):
                sim.set_efield((e, 0, 0), r)
            sim.MC_sweeps(spp)
            for r in range(R):
                gpu[r, k] = sim.polarisation(r)[0]
    for k, e in enumerate(fields):
        se = np.sqrt(ref[:, k].var(ddof=1) / len(ref) + gpu[:, k].var(ddof=1) / R)
        assert abs(ref[:, k].mean() - gpu[:, k].mean()) < 4.5 * se + 2e-3, f"E_x={e:+.2f} (point {k}): reference {ref[:, k].mean():.4f} vs GPU {gpu[:, k].mean():.4f} (se {se:.4f})"
    up, down = gpu[:, steps].mean(), gpu[:, 3 * steps].mean()
    # the loop polarises both ways; the reference's field term is +p.E (montecarlo-core.c:120-123), so P opposes E
    assert up < -0.05 and down > 0.05


def test_full_size_lattice_invariants(sn):
    """BASELINE's headline size (512^3, 1.3e8 sites, 2.1 GB) through the C ABI: size-independent properties --
    every attempt counted exactly once, unit dipoles, lengths untouched, a T = 0 quench never raises the
    energy, and the sweep count / counters survive a checkpoint-style round trip."""
    X = 512
    rng = np.random.default_rng(99)
    lat = np.empty((X, X, X, 4), np.float32)
    for x0 in range(0, X, 64):                       # chunks keep the float64 temporaries small
        v = rng.standard_normal((64, X, X, 3), dtype=np.float32)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        lat[x0:x0 + 64, ..., :3] = v
    lat[..., 3] = 1.0
    n = X ** 3
    with sn.Simulation(X, X, X, CageStrain=1.0, beta=float("inf"), seed=7) as sim:
        sim.set_lattice(lat)
        e0 = sim.total_energy(sn.SN_PREC_F32).sum()
        sim.MC_sweeps(2)
        acc, rej, vac = sim.counters()
        assert acc + rej == 2 * n and vac == 0
        e1 = sim.total_energy(sn.SN_PREC_F32).sum()
        assert e1 < e0 - 0.5 * n                       # a zero-temperature quench from a random start goes far downhill
        out = sim.get_lattice()
        assert np.array_equal(out[..., 3], lat[..., 3])
        norms = np.linalg.norm(out[::7, ::5, ::3, :3], axis=-1)
        assert np.max(np.abs(norms - 1.0)) < 2e-6
        assert sim.sweep_count() == 2
        P = sim.polarisation()
        assert np.all(np.abs(P) < 0.01)                # no net polarisation from a random start after 2 sweeps
    del lat, out


@pytest.mark.parametrize("shape,kernel", [((64, 32, 32), "auto"), ((20, 20, 28), "auto"), ((24, 16, 20), "colour")])
def test_async_double_buffer_equals_synchronous_calls(sn, shape, kernel):
    """sn_set_lattice_async / sn_get_lattice_async / sn_order_after: two handles used alternately (one lattice in flight
    over PCIe while the other is swept) give exactly what synchronous calls on one handle give."""
    import torch
    X, Y, Z = shape
    lats = [oa.random_lattice(X, Y, Z, seed=50 + i, lengths=(1.0, 0.5, 0.0), prevalence=(0.7, 0.2, 0.1)) for i in range(4)]
    lats[1][..., 3] = 1.0                              # one batch without species: the kernel specialisation must follow the data
    want = []
    with sn.Simulation(X, Y, Z, CageStrain=1.0, seed=3, kernel=kernel_id(sn, kernel)) as one:
        for lat in lats:
            one.set_lattice(lat)
            one.set_sweep_count(0)
            one.MC_sweeps(2)
            want.append(one.get_lattice())
    sims = [sn.Simulation(X, Y, Z, CageStrain=1.0, seed=3, kernel=kernel_id(sn, kernel)) for _ in range(2)]
    src = [torch.from_numpy(l).pin_memory() for l in lats]
    dst = [torch.empty((X, Y, Z, 4), dtype=torch.float32).pin_memory() for _ in lats]
    for i in range(4):
        s, o = sims[i % 2], sims[(i + 1) % 2]
        s.synchronize()
        s.set_sweep_count(0)
        s.set_lattice_async(src[i].data_ptr())
        s.order_after(o)
        s.MC_sweeps(2)
        s.get_lattice_async(dst[i].data_ptr())
    for s in sims:
        s.synchronize()
        s.close()
    for i in range(4):
        assert np.array_equal(dst[i].numpy(), want[i]), f"batch {i}"


@pytest.mark.parametrize("shape,kernel", [((24, 20, 1), "auto"), ((32, 32, 32), "tiled"), ((16, 12, 8), "colour")])
def test_batched_replica_transfer_equals_one_call_per_replica(sn, shape, kernel):
    """sn_set_lattices_async / sn_get_lattices_async move a whole replica batch with one copy and one kernel; the chain
    that follows must be the one the per-replica calls give, species flags included."""
    import torch
    X, Y, Z = shape
    reps = 5
    kid = {"auto": sn.SN_KERNEL_AUTO, "tiled": sn.SN_KERNEL_TILED, "colour": sn.SN_KERNEL_COLOUR}[kernel]
    lats = [oa.random_lattice(X, Y, Z, seed=60 + r, lengths=(1.0, 0.5, 0.0) if r % 2 else (1.0,), prevalence=(0.6, 0.3, 0.1) if r % 2 else (1.0,)) for r in range(reps)]
    block = torch.from_numpy(np.stack(lats)).pin_memory()
    out = torch.empty_like(block).pin_memory()
    res = []
    for batched in (False, True):
        with sn.Simulation(X, Y, Z, CageStrain=1.0, Efield=(0.02, 0, 0), nreplicas=reps, seed=9, kernel=kid) as sim:
            if batched:
                sim.set_lattices_async(block.data_ptr())
            else:
                for r in range(reps):
                    sim.set_lattice(lats[r], r)
            sim.MC_sweeps(3)
            if batched:
                sim.get_lattices_async(out.data_ptr()); sim.synchronize()
                got = [out[r].numpy().copy() for r in range(reps)]
                # a partial batch in the middle
                sim.get_lattices_async(out.data_ptr(), 1, 3); sim.synchronize()
                assert all(np.array_equal(out[k].numpy(), got[1 + k]) for k in range(3))
            else:
                got = [sim.get_lattice(r) for r in range(reps)]
            res.append((got, [sim.counters(r) for r in range(reps)]))
    for r in range(reps):
        assert np.array_equal(res[0][0][r], res[1][0][r]), f"replica {r}: batched transfer changed the chain"
        assert res[0][1][r] == res[1][1][r]
