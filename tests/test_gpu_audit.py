"""Attempt-by-attempt audit of the sweep kernels against the reference's site_energy / accept test.

sn_mc_sweep_audit runs one sweep of the kernel a handle would use for sn_mc_sweeps (the tiled kernel's AUDIT
instantiation on tiled handles) and returns, for every attempt, the trial dipole, the accept uniform, the dE the
kernel computed in registers, its decision and the ordinal of the group of mutually independent sites the attempt
belongs to.  The host replays the sweep in group order on a copy of the start lattice and, at every attempt's point
in time, asks the f64 oracle (the reference's site_energy with float -> double, montecarlo-core.c:76-141) for dE:

  * |dE_gpu - dE_ref| / sum|terms| < 1e-5 for every attempt (the FP32 bar of BASELINE.json's north_star),
  * the kernel's accept / reject equals the reference test `dE < 0 || exp(-dE beta) > u` (montecarlo-core.c:179)
    evaluated on dE_ref, except where dE_ref lies inside the FP32 band around the decision boundary,
  * the lattice the replay ends with is bit-identical to the one the kernel wrote; counters agree,
  * trial dipoles and uniforms are the documented functions of Philox4x32-10(site, replica, sweep) -- checked
    against an independent numpy Philox pinned by the Random123 known-answer vectors.
"""
import numpy as np
import pytest

from oracle import oracle_api as oa
from tests.helpers import PHILOX_KAT, philox4x32_10, term_scale_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


def test_philox_known_answers_host_and_device(sn):
    ck = [list(c) + list(k) for c, k, _ in PHILOX_KAT]
    host, dev = sn.philox_kat(ck, device=True)
    want = np.array([o for _, _, o in PHILOX_KAT], np.uint32)
    assert np.array_equal(host, want)
    assert np.array_equal(dev, want)
    rng = np.random.default_rng(1)
    rnd = rng.integers(0, 2 ** 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    host, dev = sn.philox_kat(rnd, device=True)
    ref = np.stack(philox4x32_10(*[rnd[:, i] for i in range(6)]), 1)
    assert np.array_equal(host, ref) and np.array_equal(dev, ref)


def _expected_draws(kind, X, Y, Z, z0, seed, tag, sweep, constrain, dim):
    """Trial dipoles and accept uniforms the kernels document, from an independent Philox.
    kind 'tiled': one call per pair of z-consecutive sites keyed by the even one, 20-bit u, v, 32-bit accept word
    (sn_sweep_tiled.cuh, draw()); kind 'colour': one call per site, 24-bit u, v, 32-bit accept word."""
    x, y, z = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    zk = z - (z & 1) if kind == "tiled" else z
    gsite = ((x.astype(np.uint64) * np.uint64(Y) + y.astype(np.uint64)) * np.uint64(Z) + (zk + z0).astype(np.uint64))
    c0 = gsite & np.uint64(0xFFFFFFFF)
    c1 = (gsite >> np.uint64(32)) ^ np.uint64(tag)
    r = philox4x32_10(c0, c1, sweep & 0xFFFFFFFF, sweep >> 32, seed & 0xFFFFFFFF, seed >> 32)
    if kind == "tiled":
        odd = (z & 1) == 1
        wa = np.where(odd, r[2], r[0]).astype(np.uint32)
        wb = np.where(odd, r[3], r[1]).astype(np.uint32)
        u = (wb >> 12).astype(np.float32) * np.float32(1.0 / 1048576.0)
        v = (((wb & 0xFFF) << 8) | (wa & 0xFF)).astype(np.float32) * np.float32(1.0 / 1048576.0)
        ua = wa.astype(np.float32) * np.float32(1.0 / 4294967296.0)
    else:
        u = (r[0] >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0)
        v = (r[1] >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0)
        ua = r[2].astype(np.float32) * np.float32(1.0 / 4294967296.0)
    if constrain:
        i = np.minimum((u * np.float32(6.0)).astype(np.int32), 5)
        s = np.where(i & 1, -1.0, 1.0).astype(np.float32)
        npd = np.zeros(u.shape + (3,), np.float32)
        for a in range(3):
            npd[..., a] = np.where((i >> 1) == a, s, 0.0)
    else:
        ang = (np.float32(6.283185307179586) * v).astype(np.float64)
        if dim < 3:
            npd = np.stack([np.cos(ang), np.sin(ang), np.zeros_like(ang)], -1)
        else:
            zc = (np.float32(1.0) - np.float32(2.0) * u).astype(np.float64)
            rr = np.sqrt(np.maximum(0.0, 1.0 - zc * zc))
            npd = np.stack([rr * np.cos(ang), rr * np.sin(ang), zc], -1)
    return npd, ua


def _replay(p, lat0, rec, final, what):
    """Replay one replica's audit records in group order against the f64 oracle; returns summary numbers."""
    o64 = oa.Oracle("f64")
    dxyz, d = oa.Oracle("f32").neighbours(p)
    X, Y, Z = p.X, p.Y, p.Z
    lat = np.ascontiguousarray(lat0, np.float32).copy()
    group = rec[..., 6].astype(np.int64)
    flag = rec[..., 5]
    vacant = lat0[..., 3] == 0
    assert np.array_equal(flag == 2, vacant), f"{what}: vacancy flags differ from the lattice (montecarlo-core.c:163)"
    assert np.all(np.isin(flag, (0.0, 1.0, 2.0)))
    norm = np.linalg.norm(rec[..., :3].astype(np.float64), axis=-1)
    assert np.all(np.abs(norm[~vacant] - 1.0) < 3e-6), f"{what}: a live site has no (unit) trial dipole recorded -- attempt missing"
    beta = float(p.beta)
    worst, n_flip, n_live, checked_groups = 0.0, 0, 0, 0
    rng = np.random.default_rng(0)
    order = np.unique(group)
    spot = set(rng.choice(order, size=min(6, len(order)), replace=False).tolist())
    for g in order:
        mask = group == g
        if g in spot:                                    # mutually independent: no two sites of a group within the cut-off
            for (dx, dy, dz) in dxyz:
                assert not (mask & np.roll(mask, (dx, dy, dz), (0, 1, 2))).any(), f"{what}: group {g} holds interacting sites"
            checked_groups += 1
        sites = np.argwhere(mask & ~vacant).astype(np.int32)
        if len(sites) == 0:
            continue
        sx, sy, sz = sites[:, 0], sites[:, 1], sites[:, 2]
        nd = rec[sx, sy, sz, 0:3]
        ua = rec[sx, sy, sz, 3].astype(np.float64)
        de_gpu = rec[sx, sy, sz, 4].astype(np.float64)
        dec = rec[sx, sy, sz, 5] == 1.0
        de_ref = o64.site_energy(p, lat, sites, nd)                      # the reference's site_energy, at this point of the chain
        scale = np.maximum(term_scale_batch(p, lat, sites, nd, dxyz, d), 1e-3)   # a ConstrainToX trial may equal the old dipole: every term 0
        err = np.abs(de_gpu - de_ref) / scale
        worst = max(worst, float(err.max()))
        assert err.max() < 1e-5, f"{what}: group {g}: |dE_gpu - dE_ref| / sum|terms| = {err.max():.3g}"
        de32 = de_ref.astype(np.float32).astype(np.float64)              # `float dE = site_energy(...)`, montecarlo-core.c:154
        with np.errstate(over="ignore", invalid="ignore"):
            ref_dec = (de32 < 0.0) | (np.exp(-de32 * beta) > ua)         # montecarlo-core.c:179
        flip = dec != ref_dec
        if flip.any():
            # the decision boundary sits at dE* = -ln(u) / beta (and at 0 when beta = inf); a flip needs dE_ref within the
            # FP32 error band of it: 1e-5 sum|terms| from the field gather plus the fast exponential's relative error
            with np.errstate(divide="ignore"):
                star = np.where(np.isinf(beta), 0.0, -np.log(np.maximum(ua, 1e-300)) / beta) if beta > 0 else np.zeros_like(ua)
            band = 1e-5 * scale + (0.0 if np.isinf(beta) or beta == 0 else 4e-6 * (1.0 + np.abs(de32 * beta)) / beta)
            off = np.abs(de32 - star)
            assert np.all(off[flip] <= band[flip]), f"{what}: group {g}: decision differs outside the FP32 band (off {off[flip].max():.3g})"
            n_flip += int(flip.sum())
        n_live += len(sites)
        a = sites[dec]
        lat[a[:, 0], a[:, 1], a[:, 2], 0:3] = nd[dec]                    # follow the kernel's own trajectory
    assert np.array_equal(lat, final), f"{what}: the replayed lattice differs from the one the kernel wrote"
    assert n_flip <= max(2, 5e-4 * n_live), f"{what}: {n_flip} decisions of {n_live} differ from the reference test"
    return worst, n_flip, n_live, len(order)


CASES = [
    # name, shape, lengths/prevalence, K, Efield, constrain, DIM, T
    ("unit lengths", (32, 32, 32), None, 0.0, (0.0, 0.0, 0.0), False, 3, 300),
    ("species + vacancies", (64, 32, 32), ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), 0.0, (0.02, 0.0, 0.0), False, 3, 300),
    ("K and field", (32, 32, 32), None, 0.6, (0.05, -0.03, 0.02), False, 3, 150),
    ("ConstrainToX", (32, 32, 32), ((1.0, 0.5), (0.7, 0.3)), 0.0, (0.01, 0.0, 0.0), True, 3, 300),
    ("DIM = 2", (32, 64, 32), None, 0.3, (0.0, 0.02, 0.0), False, 2, 75),
    ("T = 0", (32, 32, 32), None, 0.0, (0.0, 0.0, 0.0), False, 3, 0),
    ("odd tile counts", (48, 32, 48), ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), 0.0, (0.02, 0.0, 0.0), False, 3, 300),   # 3 x 2 x 3 tiles: 18 phases
    ("partial tiles", (40, 35, 44), ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), 0.3, (0.02, -0.01, 0.0), False, 3, 300),     # 3 x 3 x 3 tiles, the last of every axis cut
    ("two small tiles", (20, 27, 24), ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), 0.0, (0.02, 0.0, 0.01), False, 3, 300),     # 2 x 2 x 2 tiles, 16 + a few cells per axis
    ("cut-off 2", (32, 48, 32), ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), 0.4, (0.02, 0.0, -0.01), False, 3, 300, 2),       # DipoleCutOff = 2: 32 neighbours, ghost shell of 2
]


@pytest.mark.parametrize("kernel", ["tiled", "tiled_phased", "colour"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_every_attempt_matches_the_reference(sn, case, kernel):
    name, (X, Y, Z), species, K, E, constrain, dim, T = case[:8]
    cutoff = case[8] if len(case) > 8 else 3
    kern = {"tiled": sn.SN_KERNEL_TILED, "tiled_phased": sn.SN_KERNEL_TILED_PHASED, "colour": sn.SN_KERNEL_COLOUR}[kernel]
    if kernel != "tiled" and name not in ("species + vacancies", "K and field", "odd tile counts", "cut-off 2"):
        pytest.skip("the phased launch and the colour passes are audited on two cases")
    E = tuple(float(np.float32(v)) for v in E)
    lengths, prev = species if species else ((1.0,), (1.0,))
    reps, seed, cage = 2, 0xDEADBEEF + T, (1.0, 2.0)
    lats = [oa.random_lattice(X, Y, Z, seed=80 + r, lengths=lengths, prevalence=prev) for r in range(reps)]
    beta = sn.beta_of_T(T)
    with sn.Simulation(X, Y, Z, DipoleCutOff=cutoff, CageStrain=1.0, K=K, Efield=E, beta=beta, ConstrainToX=constrain, DIM=dim, nreplicas=reps,
                       seed=seed, kernel=kern) as sim:
        for r in range(reps):
            sim.set_lattice(lats[r], r)
            sim.set_cagestrain(cage[r], r)
        sim.MC_sweeps(2)                                   # audit a sweep from the middle of a chain, not the first one
        start = [sim.get_lattice(r) for r in range(reps)]
        c0 = [sim.counters(r) for r in range(reps)]
        rec = sim.MC_sweep_audit()
        final = [sim.get_lattice(r) for r in range(reps)]
        c1 = [sim.counters(r) for r in range(reps)]
        assert sim.sweep_count() == 3
    for r in range(reps):
        p = oa.make_params(X, Y, Z, cutoff, cage[r], K, E, beta, constrain, dim, T)
        worst, n_flip, n_live, ngroups = _replay(p, start[r], rec[r], final[r], f"{name} / {kernel} / replica {r}")
        flag = rec[r][..., 5]
        assert tuple(b - a for a, b in zip(c0[r], c1[r])) == (int((flag == 1).sum()), int((flag == 0).sum()), int((flag == 2).sum()))
        phases = int(np.prod([3 if ((n + 15) // 16) % 2 else 2 for n in (X, Y, Z)]))   # tile colours per axis: 2, or 3 for an odd tile count (partial tiles count)
        if name in ("partial tiles", "two small tiles"):                        # a cut tile has no sites in some (cx, cy) classes: those groups are empty
            assert phases * 48 <= ngroups <= phases * 64
        else:
            colour_groups = int(np.prod([(cutoff + 1) + n % (cutoff + 1) for n in (X, Y, Z)]))   # period cutoff + 1, leftover planes get colours of their own
            assert ngroups == (phases * 64 if kernel != "colour" else colour_groups)
        # the random numbers behind the records
        npd, ua = _expected_draws("colour" if kernel == "colour" else "tiled", X, Y, Z, 0, seed, r << 8, 2, constrain, dim)
        live = flag != 2
        assert np.array_equal(rec[r][..., 3][live], ua[live]), "accept uniforms are not the documented Philox words"
        assert np.max(np.abs(rec[r][..., :3][live] - npd[live])) < 4e-6, "trial dipoles are not the documented function of Philox"
        print(f"{name} / {kernel} / replica {r}: {n_live} attempts, max |dE err| / sum|terms| = {worst:.2e}, {n_flip} boundary decisions")


def test_audit_sweep_is_the_product_sweep(sn):
    """The AUDIT instantiation must walk the same chain as the product kernel: one audited sweep == one sn_mc_sweeps(1)."""
    X = 32
    lat = oa.random_lattice(X, X, X, seed=90, lengths=(1.0, 0.5, 0.0), prevalence=(0.6, 0.3, 0.1))
    outs = []
    for audit in (False, True):
        with sn.Simulation(X, X, X, CageStrain=1.0, Efield=(0.02, 0, 0), seed=5, kernel=sn.SN_KERNEL_TILED) as sim:
            sim.set_lattice(lat)
            sim.MC_sweeps(1)
            if audit:
                sim.MC_sweep_audit()
            else:
                sim.MC_sweeps(1)
            sim.MC_sweeps(1)
            outs.append((sim.get_lattice(), sim.counters()))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


@pytest.mark.parametrize("T", [75, 300, 600])
def test_tiled_kernel_matches_the_reference_serial_chain(sn, T):
    """Equilibrium <E>/N, <P_x>, acceptance of the TILED kernel (32^3; two species + vacancies at T = 300) against the
    reference's serial random-site MT19937 chain (oracle f32, bit-equal to the reference) on the same lattice.
    Both chains start from the same pre-equilibrated state (an anneal 2T -> T, run on the GPU: at 75 K a random start
    coarsens for thousands of sweeps and the two update orders relax at different rates, which is kinetics, not the
    stationary distribution under test); if the tiled kernel sampled a different distribution, the reference chain
    would walk away from that state.  Independent seeds on both sides, agreement within 4.5 combined standard errors."""
    X = 32
    species = T == 300
    lengths, prev = ((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)) if species else ((1.0,), (1.0,))
    Ex = float(np.float32(0.3))
    beta = sn.beta_of_T(T)
    p = oa.make_params(X, X, X, 3, 1.0, 0.0, (Ex, 0.0, 0.0), beta, 0, 3, T)
    lat0 = oa.random_lattice(X, X, X, seed=91, lengths=lengths, prevalence=prev)
    n = X ** 3
    with sn.Simulation(X, X, X, CageStrain=1.0, Efield=(Ex, 0, 0), beta=sn.beta_of_T(2 * T), seed=5, kernel=sn.SN_KERNEL_TILED) as sim:
        sim.set_lattice(lat0)
        sim.MC_sweeps(150)
        sim.set_T(T)
        sim.MC_sweeps(600)
        lat_eq = sim.get_lattice()
    o = oa.Oracle("f32")
    eqm, nsamp, stride = 12, 6, 3
    ref = []
    for s in range(400, 404):
        lat = np.ascontiguousarray(lat_eq, np.float32).copy()
        mt = o.mt(s)
        o.mc_moves(p, lat, mt, eqm * n)
        es, ps, acc, rej = [], [], 0, 0
        for _ in range(nsamp):
            a, rj = o.mc_moves(p, lat, mt, stride * n)
            acc += a; rej += rj
            es.append(o.total_energy(p, lat).sum() / n)
            ps.append(o.polarisation(p, lat))
        ref.append((np.mean(es), np.mean(ps), acc / (acc + rej)))
    ref = np.array(ref)
    R = 8
    with sn.Simulation(X, X, X, CageStrain=1.0, Efield=(Ex, 0, 0), beta=beta, nreplicas=R, seed=777 + T, kernel=sn.SN_KERNEL_TILED) as sim:
        for r in range(R):
            sim.set_lattice(lat_eq, r)
        sim.MC_sweeps(eqm)
        sim.reset_counters()
        es, ps = np.zeros((R, nsamp)), np.zeros((R, nsamp))
        for k in range(nsamp):
            sim.MC_sweeps(stride)
            for r in range(R):
                es[r, k] = sim.total_energy(sn.SN_PREC_F32, r).sum() / n
                ps[r, k] = sim.polarisation(r)[0]
        ratio = np.array([sim.counters(r)[0] / sum(sim.counters(r)[:2]) for r in range(R)])
    gpu = np.stack([es.mean(1), ps.mean(1), ratio], 1)
    for col, what in enumerate(("energy per site", "polarisation", "acceptance ratio")):
        m_ref, m_gpu = ref[:, col].mean(), gpu[:, col].mean()
        se = np.sqrt(ref[:, col].var(ddof=1) / len(ref) + gpu[:, col].var(ddof=1) / R)
        floor = 1.5e-3 if T < 100 else 3e-4          # at 75 K the state still coarsens slowly, at slightly different rates per update order
        assert abs(m_ref - m_gpu) < 4.5 * se + floor, f"T={T}: {what}: reference {m_ref:.5f} vs tiled kernel {m_gpu:.5f} (se {se:.5f})"
