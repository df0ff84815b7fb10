"""CPU tests (no GPU): the oracle restatement against the golden vectors that were
generated from the reference's own code, and -- where /root/reference (or the
prebuilt oracle/_ref) is present -- against that code directly."""
import numpy as np
import pytest

from oracle import oracle_api as oa
from tests.helpers import CASE_NAMES, GOLDEN, load_case

PRECS = ["f32", "f64"]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_energy_matches_golden(built, name, prec):
    g, p = load_case(name)
    o = oa.Oracle(prec)
    dxyz, d = o.neighbours(p)
    assert np.array_equal(dxyz, g["nb_dxyz_f32"])               # gen_neighbour order, montecarlo-core.c:47-62
    assert np.array_equal(d, g["nb_d_f32"])
    assert np.array_equal(o.site_energy(p, g["lattice"], g["sites"], g["newdip"]), g[f"dE_{prec}"])
    assert np.array_equal(o.site_interaction_map(p, g["lattice"]), g[f"interaction_{prec}"])
    # the per-site values above are bit-equal; the golden total was summed pairwise by numpy, the oracle
    # sums in site order, so the totals agree to summation rounding only
    assert np.allclose(o.total_energy(p, g["lattice"]), g[f"total_{prec}"], rtol=1e-13, atol=1e-11)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_chain_matches_golden(built, name, prec):
    """The MT19937-driven serial chain (MC_moves, montecarlo-core.c:143-191) bit for bit."""
    g, p = load_case(name)
    o = oa.Oracle(prec)
    lat = np.ascontiguousarray(g["lattice"], o.dtype)
    acc, rej = o.mc_moves(p, lat, o.mt(0xDEADBEEF + 300), 4000)
    assert [acc, rej] == list(g[f"chain_counters_{prec}"])
    assert np.array_equal(lat, g[f"chain_lattice_{prec}"])


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_observables_match_golden(built, name, prec):
    g, p = load_case(name)
    o = oa.Oracle(prec)
    lat = g["lattice"]
    assert o.polarisation(p, lat) == float(g[f"polarisation_{prec}"])
    assert o.landau_order(p, lat) == float(g[f"landau_{prec}"])
    if f"potential_{prec}" in g:
        assert np.array_equal(o.potential_map(p, lat), g[f"potential_{prec}"])
    if f"rdf_{prec}" in g:
        fe, afe, cnt = o.rdf(p, lat)
        rows = g[f"rdf_{prec}"]                      # r2, r, FE, AFE, count, T as printed (analysis.c:589)
        nz = np.nonzero(cnt)[0]
        assert np.array_equal(rows[:, 0].astype(int), nz)
        assert np.array_equal(rows[:, 4].astype(np.int64), cnt[nz])
        assert np.allclose(rows[:, 2], (fe[nz] / cnt[nz].astype(o.dtype)), atol=6e-7)   # file holds %f (6 decimals)
        assert np.allclose(rows[:, 3], (afe[nz] / cnt[nz].astype(o.dtype)), atol=6e-7)
        assert rows[0, 2] == 1.0 and rows[0, 3] == 1.0            # r^2 = 0 row is the self-correlation


def test_rdf_counts_known_answer(built):
    """orientational_count[r^2] = N x (#lattice vectors of that r^2), independent of the
    configuration; the historical data file in the reference shows the same numbers for 20^3."""
    p = oa.make_params(9, 9, 9)
    lat = oa.random_lattice(9, 9, 9, seed=4)
    _, _, cnt = oa.Oracle("f32").rdf(p, lat)
    per_site = {1: 6, 2: 12, 3: 8, 4: 6, 5: 24, 6: 24, 8: 12, 9: 30}
    for r2, m in per_site.items():
        assert cnt[r2] == m * 729
    assert cnt[7] == 0 and cnt[0] == 729


def test_initial_lattices_match_golden(built):
    g = dict(np.load(f"{GOLDEN}/initial_lattices.npz"))
    o = oa.Oracle("f32")
    p = oa.make_params(20, 20, 28)
    for kind in ("random", "ferroelectric", "buckled", "antiferro_wall", "ferro_wall", "antiferro_slip", "spectrum"):
        mt = o.mt(0xDEADBEEF + 300)
        lat = o.initialise_lattice(p, mt, kind)
        o.solid_solution(p, lat, mt, [1.0, 0.0, 0.0], [1.0, 0.0, 0.0])
        assert np.array_equal(lat, g[kind]), kind
    mt = o.mt(0xDEADBEEF + 300)
    lat = o.initialise_lattice(p, mt, "random")
    histo = o.solid_solution(p, lat, mt, [1.0, 0.5, 0.0], [0.6, 0.3, 0.1])
    assert np.array_equal(lat, g["random_mixed"])
    assert histo.sum() == 20 * 20 * 28
    # analytic states: ferroelectric => polarisation 1, FE correlation 1 at every r
    fe_lat = g["ferroelectric"]
    assert o.polarisation(p, fe_lat) == 1.0
    fe, afe, cnt = o.rdf(p, fe_lat)
    nz = cnt > 0
    assert np.all(fe[nz] == cnt[nz])


def test_mt19937_known_answers(built):
    g = dict(np.load(f"{GOLDEN}/mt19937.npz"))
    o = oa.Oracle("f32")
    mt = o.mt(5489)
    got = [o.lib.sno_mt_int32(oa.C.byref(mt)) for _ in range(16)]
    assert got == [int(v) for v in g["int32_seed5489"]]
    assert got[0] == 3499211612                       # published first output of MT19937 for seed 5489
    mt = o.mt(0xDEADBEEF + 300)
    assert [o.lib.sno_mt_real1(oa.C.byref(mt)) for _ in range(8)] == list(g["real1"])


@pytest.mark.skipif(not oa.ref_available("f32"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("prec", PRECS)
def test_oracle_vs_reference_build_fresh_inputs(built, prec):
    """Beyond the committed vectors: new seeded inputs through the reference's own code."""
    r = oa.RefLib(prec)
    o = oa.Oracle(prec)
    E = tuple(float(np.float32(v)) for v in (0.03, 0.01, -0.02))
    p = oa.make_params(10, 12, 9, 3, 0.8, 0.4, E, beta=0.9)
    lat = oa.random_lattice(10, 12, 9, seed=77, lengths=(1.0, 0.3, 0.0), prevalence=(0.5, 0.4, 0.1))
    r.configure(p)
    r.set_lattice(lat)
    rng = np.random.default_rng(1)
    sites = np.stack([rng.integers(0, 10, 300), rng.integers(0, 12, 300), rng.integers(0, 9, 300)], 1).astype(np.int32)
    nd = rng.normal(size=(300, 3)); nd /= np.linalg.norm(nd, axis=1, keepdims=True); nd = nd.astype(np.float32)
    assert np.array_equal(o.site_energy(p, lat, sites, nd), r.site_energy(sites, nd))
    r.seed(123)
    cr = r.mc_moves(20000)
    lo = np.ascontiguousarray(lat, o.dtype)
    co = o.mc_moves(p, lo, o.mt(123), 20000)
    assert cr == co
    assert np.array_equal(lo.astype(np.float64), r.get_lattice())
    assert np.array_equal(o.potential_map(p, lo), r.potential_map())
    assert o.polarisation(p, lo) == r.polarisation()
    assert o.landau_order(p, lo) == r.landau_order()


RECOMB_KEYS = ["ZBe", "ZBh", "ZFDe", "ZFDh", "R_Boltz", "R_FD", "FD-Total-electron", "FD-Total-hole"]
EXTRA_CASES = ["species3d", "flat2d", "odd_cut2", "cut4_constrain", "dim2"]


def parse_recombination_log(txt):
    """`T: %d ZBe: %e ... R_Boltz: %e R_FD: %e FD-Total-electron: %e FD-Total-hole: %e` (analysis.c:129-131,169-171)."""
    import re
    return np.array([float(re.search(re.escape(k) + r":\s*(\S+)", txt).group(1)) for k in RECOMB_KEYS])


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", EXTRA_CASES)
def test_efield_maps_and_recombination_match_golden(built, name, prec):
    """dipole_electricfield / dipole_electricfieldoffset (analysis.c:310-465) bit for bit, and the numbers
    recombination_calculator logs (analysis.c:96-171) to the 7 digits it prints."""
    g, p = load_case(name)
    x = dict(np.load(f"{GOLDEN}/analysis_extra.npz"))
    o = oa.Oracle(prec)
    lat = g["lattice"]
    assert np.array_equal(o.efield_map(p, lat, 4, False), x[f"{name}_efield_{prec}"])
    assert np.array_equal(o.efield_map(p, lat, 2, True), x[f"{name}_efieldoffset_{prec}"])
    ref = parse_recombination_log(bytes(x[f"{name}_recombination_{prec}"]).decode())
    got = o.recombination(p, lat)
    assert np.allclose(got[:8], ref, rtol=1.5e-6)
    assert got[6] == pytest.approx(1.0, abs=1e-12) and got[7] == pytest.approx(1.0, abs=1e-12)   # densities are normalised
