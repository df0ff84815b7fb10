"""The C driver (driver/starrynight_b200_main.c): same cfg keys, same initial state and the
same output files as the reference's main() (starrynight-main.c)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "driver", "starrynight-b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_files():
    return {k: bytes(v) for k, v in np.load(os.path.join(GOLDEN, "reference_run_files.npz")).items()}


def run_init_only(tmp_path, cfg_text, *args):
    (tmp_path / "starrynight.cfg").write_text(cfg_text)
    out = tmp_path / "init.bin"
    subprocess.run([DRIVER, "--init-only", str(out), *args], cwd=tmp_path, check=True, capture_output=True)
    return np.fromfile(out, np.float32)


def test_initial_state_matches_reference(built, tmp_path):
    """Stock starrynight.cfg: antiferro_wall, one species -- bit-equal to what the reference builds
    from init_genrand(0xDEADBEEF + T) (main.c:172-205)."""
    cfg = ref_files()["starrynight.cfg"].decode()
    g = dict(np.load(os.path.join(GOLDEN, "initial_lattices.npz")))
    lat = run_init_only(tmp_path, cfg).reshape(20, 20, 28, 4)
    assert np.array_equal(lat, g["antiferro_wall"])
    mixed = cfg.replace('InitialLattice="antiferro_wall"', 'InitialLattice="random"')
    mixed = mixed.replace("Dipoles    = [ 1.0, 0.0, 0.0]", "Dipoles    = [ 1.0, 0.5, 0.0]").replace("Prevalence = [ 1.0, 0.0, 0.0]", "Prevalence = [ 0.6, 0.3, 0.1]")
    lat = run_init_only(tmp_path, mixed).reshape(20, 20, 28, 4)
    assert np.array_equal(lat, g["random_mixed"])
    for kind in ("ferroelectric", "buckled", "ferro_wall", "antiferro_slip", "spectrum"):
        lat = run_init_only(tmp_path, cfg.replace('"antiferro_wall"', f'"{kind}"')).reshape(20, 20, 28, 4)
        assert np.array_equal(lat, g[kind]), kind


def test_cfg_type_rules_and_overrides(built, tmp_path):
    """libconfig's strict typing (config.c:123-179): a float where an int is expected is ignored,
    groups/arrays/comments parse, argv[1] overrides T (main.c:142-146) and thereby the seed."""
    cfg = ref_files()["starrynight.cfg"].decode()
    assert run_init_only(tmp_path, cfg.replace("Z=28", "Z=28.0")).size == 20 * 20 * 20 * 4      # default Z kept
    assert run_init_only(tmp_path, cfg.replace("Z=28", "Z : 12 ; // comment")).size == 20 * 20 * 12 * 4
    rnd = cfg.replace('"antiferro_wall"', '"random"')
    a = run_init_only(tmp_path, rnd)
    b = run_init_only(tmp_path, rnd, "310")                  # T = 310 -> seed 0xDEADBEEF + 310
    c = run_init_only(tmp_path, rnd.replace("T: 300", "T: 310"))
    assert not np.array_equal(a, b) and np.array_equal(b, c)
    bad = subprocess.run([DRIVER, "--init-only", "x.bin"], cwd=tmp_path / "..", capture_output=True)
    (tmp_path / "starrynight.cfg").write_text("X = [1, 2")
    bad = subprocess.run([DRIVER, "--init-only", "x.bin"], cwd=tmp_path, capture_output=True)
    assert bad.returncode != 0 and b"starrynight.cfg" in bad.stderr


def _floats(text, col):
    return np.array([float(l.split()[col]) for l in text.splitlines() if l.strip() and not l.startswith("#")])


@pytest.mark.gpu
def test_driver_run_writes_the_reference_files(built, tmp_path):
    ref = ref_files()
    (tmp_path / "starrynight.cfg").write_bytes(ref["starrynight.cfg"])
    run = subprocess.run([DRIVER], cwd=tmp_path, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "Neighbour list generated: 122 neighbours found with DipoleCutOff=3." in run.stderr
    assert "MC Moves (per second):" in run.stderr and "ACCEPT:" in run.stderr
    produced = set(os.listdir(tmp_path))
    for fn in ref:
        assert fn in produced, f"{fn} missing"
    assert "Recombination_T_0300.log" in produced
    # the initial analysis is deterministic: same lattice, so the same numbers in the same format
    got = (tmp_path / "initial_lattice_potential.xyz").read_text()
    want = ref["initial_lattice_potential.xyz"].decode()
    assert len(got.splitlines()) == len(want.splitlines()) == 20 * 20 * 28
    assert [l.split()[:3] for l in got.splitlines()] == [l.split()[:3] for l in want.splitlines()]
    assert np.max(np.abs(_floats(got, 3) - _floats(want, 3))) < 3e-6            # %f, reference sums float terms
    gc, wc = (tmp_path / "initial_lattice_potential.cube").read_text(), ref["initial_lattice_potential.cube"].decode()
    assert gc.splitlines()[:7] == wc.splitlines()[:7]
    gv = np.array(" ".join(gc.splitlines()[7:]).split(), float)
    wv = np.array(" ".join(wc.splitlines()[7:]).split(), float)
    assert gv.shape == wv.shape and np.allclose(gv, wv, rtol=2e-5, atol=3e-6)   # %g keeps 6 significant digits
    gp, wp = (tmp_path / "initial_pot.png").read_text().split(), ref["initial_pot.png"].decode().split()
    assert gp[:4] == wp[:4] and np.max(np.abs(np.array(gp[4:], int) - np.array(wp[4:], int))) <= 1
    gr, wr = (tmp_path / "rdf.dat").read_text(), ref["rdf.dat"].decode()
    assert gr.splitlines()[0] == wr.splitlines()[0]
    # the reference accumulates ~1e6 AFE terms per bin in a float (analysis.c:546-577): its printed AFE is
    # off by up to 5e-4 from the exact sum on this lattice; the kernel sums in FP64 (pinned against the
    # float->double reference build in test_gpu_observables.py)
    for col, tol in ((0, 0), (1, 1e-6), (2, 2e-6), (3, 1e-3), (4, 0), (5, 0)):
        assert np.max(np.abs(_floats(gr, col) - _floats(wr, col))) <= tol
    # production step files: same names, same shapes
    assert len((tmp_path / "T_0300_1_000_potential.xyz").read_text().splitlines()) == 20 * 20 * 28
    assert len(_floats((tmp_path / "T_0300_1_000-RDF.dat").read_text(), 0)) == len(_floats(ref["T_0300_1_000-RDF.dat"].decode(), 0))


@pytest.mark.gpu
def test_driver_efield_maps_and_recombination_log(built, tmp_path):
    """CalculateEfield / CalculateRecombination (main.c:31-32,47,73,85-86,225): same files, same line formats,
    numbers equal to the reference's routines on the same initial lattice."""
    from oracle import oracle_api as oa
    from tests.test_oracle import parse_recombination_log
    cfg = ref_files()["starrynight.cfg"].decode()
    cfg = cfg.replace("X=20", "X=12").replace("Y=20", "Y=10").replace("Z=28", "Z=12")
    cfg = cfg.replace("CalculateEfield: false", "CalculateEfield: true").replace("CalculateRecombination: false", "CalculateRecombination: true")
    if "CalculateEfield: true" not in cfg:
        cfg += "\nCalculateEfield: true\n"
    if "CalculateRecombination: true" not in cfg:
        cfg += "\nCalculateRecombination: true\n"
    cfg = cfg.replace("MCMegaSteps: 1", "MCMegaSteps: 2")                  # the stored cfg is the shortened one (1 / 1 / 2.0)
    assert "MCMegaSteps: 2" in cfg and "CalculateEfield: true" in cfg and "X=12" in cfg
    (tmp_path / "starrynight.cfg").write_text(cfg)
    run = subprocess.run([DRIVER], cwd=tmp_path, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    lat0 = run_init_only(tmp_path, cfg).reshape(12, 10, 12, 4)
    p = oa.make_params(12, 10, 12)
    o = oa.Oracle("f64")
    for fn, cut, half in (("initial_lattice_efield.xyz", 4, False), ("initial_lattice_efieldoffset.xyz", 2, True)):
        txt = (tmp_path / fn).read_text()
        assert len(txt.splitlines()) == 12 * 10 * 12
        assert np.max(np.abs(_floats(txt, 3) - o.efield_map(p, lat0, cut, half))) < 1e-6        # %f
    for fn in ("equilib_lattice_efield.xyz", "T_0300_1_000_efield.xyz", "T_0300_1_001_efield.xyz"):
        assert len((tmp_path / fn).read_text().splitlines()) == 12 * 10 * 12
    # initial recombination numbers go to stderr (main.c:47), one line per production mega-step to the log (:73)
    first = [l for l in run.stderr.splitlines() if l.startswith("T: 300 ZBe:") and "R_FD:" in l][0]
    assert np.allclose(parse_recombination_log(first), o.recombination(p, lat0)[:8], rtol=1.5e-6)
    log = (tmp_path / "Recombination_T_0300.log").read_text().splitlines()
    rows = [l for l in log if l.startswith("T: 300 ZBe:")]
    assert len(rows) == 2 and all("R_FD:" in l and "FD-Total-hole:" in l for l in rows)


@pytest.mark.gpu
@pytest.mark.parametrize("kern", ["auto", "colour"])
def test_driver_two_gpu_slabs_write_identical_files(built, tmp_path, kern):
    """GPUs = 2 (Z-slabs, boundary pushes over NVLink) must reproduce the single-GPU run file for file: the decomposed
    chain is bit-identical and the analysis runs on the slabs themselves (no gather).  On a one-GPU box both slabs share
    the device.  Kernel = "colour" with many sweeps per mega-step is the case where one host thread must not queue a
    whole mega-step per slab (the launch queue would fill before the second slab has queued anything)."""
    cfg = ref_files()["starrynight.cfg"].decode()
    cfg = cfg.replace("X=20", "X=32").replace("Y=20", "Y=32").replace("Z=28", "Z=64").replace('"antiferro_wall"', '"random"')
    moves = "24.0" if kern == "colour" else "3.0"
    cfg = cfg.replace("MCMegaSteps: 1", "MCMegaSteps: 2").replace("MCMoves: 2.0", f"MCMoves: {moves}")
    assert "MCMegaSteps: 2" in cfg and f"MCMoves: {moves}" in cfg and "Z=64" in cfg
    cfg += '\nHysteresis : { amplitude = 0.1; steps = 2; cycles = 1; };\n'
    if kern == "colour":
        cfg += 'Kernel = "colour";\n'
    outs = []
    for n in (1, 2):
        d = tmp_path / f"gpus{n}"
        d.mkdir()
        (d / "starrynight.cfg").write_text(cfg + f"\nGPUs = {n};\n")
        run = subprocess.run([DRIVER], cwd=d, capture_output=True, text=True)
        assert run.returncode == 0, run.stderr[-2000:]
        if n == 2:
            assert "Z-slab decomposition: 2 GPUs x 32 planes" in run.stderr
        files = {fn: (d / fn).read_bytes() for fn in sorted(os.listdir(d)) if not fn.startswith("Recombination") and fn != "starrynight.cfg"}
        acc = [l for l in run.stderr.splitlines() if "ACCEPT:" in l][0]
        outs.append((files, run.stdout, acc))
    assert outs[0][0].keys() == outs[1][0].keys() and len(outs[0][0]) >= 8
    for fn in outs[0][0]:
        assert outs[0][0][fn] == outs[1][0][fn], fn
    assert outs[0][1] == outs[1][1] and "Polar:" in outs[0][1]          # hysteresis trace (main.c:82 format)
    assert outs[0][2] == outs[1][2]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(20, 20, 28), (32, 32, 32)])
def test_driver_checkpoint_restart_continues_the_chain(built, tmp_path, shape):
    """Checkpoint after every mega-step, Restart from it: the resumed run writes the same later files as an
    uninterrupted one (lattice + sweep counter = Philox counter + ACCEPT/REJECT is the whole state).  Both
    sweep kernels: 20x20x28 runs colour passes, 32^3 the tiled dataflow kernel (tile versions restored)."""
    X, Y, Z = shape
    cfg = ref_files()["starrynight.cfg"].decode()
    cfg = cfg.replace("X=20", f"X={X}").replace("Y=20", f"Y={Y}").replace("Z=28", f"Z={Z}").replace('"antiferro_wall"', '"random"')
    runs = {}
    for name, steps, extra in (("full", 4, ""), ("first", 2, 'Checkpoint = "ck.bin";'), ("resumed", 4, 'Restart = "../first/ck.bin";')):
        d = tmp_path / name
        d.mkdir()
        (d / "starrynight.cfg").write_text(cfg.replace("MCMegaSteps: 1", f"MCMegaSteps: {steps}") + "\n" + extra + "\n")
        if name == "resumed":                          # the normal use: resume where the first run left its files
            (d / "Recombination_T_0300.log").write_text((runs["first"][0] / "Recombination_T_0300.log").read_text())
        run = subprocess.run([DRIVER], cwd=d, capture_output=True, text=True)
        assert run.returncode == 0, run.stderr[-2000:]
        runs[name] = (d, [l for l in run.stderr.splitlines() if "ACCEPT:" in l][0])
        assert ("Restart from" in run.stderr) == (name == "resumed")
    for step in (2, 3):
        for suffix in ("_potential.xyz", "-RDF.dat", "_potential.png"):
            fn = f"T_0300_1_{step:03d}{suffix}"
            assert (runs["full"][0] / fn).read_bytes() == (runs["resumed"][0] / fn).read_bytes(), fn
    assert not (runs["resumed"][0] / "T_0300_1_001_potential.xyz").exists()      # resumed at mega-step 2
    assert runs["full"][1] == runs["resumed"][1]                                   # counters carried over
    # the resumed run appends to the log of the interrupted one instead of truncating it
    first_log = (runs["first"][0] / "Recombination_T_0300.log").read_text()
    resumed_log = (runs["resumed"][0] / "Recombination_T_0300.log").read_text()
    assert resumed_log.startswith(first_log) and "# restarted from" in resumed_log[len(first_log):]


@pytest.mark.gpu
def test_driver_temperature_batch_equals_separate_runs(built, tmp_path):
    """Temperatures = [...]: all T as replicas of one handle; every T-tagged file equals the one a separate
    run at that T (argv[1], main.c:142-146) writes."""
    cfg = ref_files()["starrynight.cfg"].decode()
    cfg = cfg.replace("X=20", "X=16").replace("Y=20", "Y=12").replace("Z=28", "Z=12").replace('"antiferro_wall"', '"random"')
    cfg = cfg.replace("MCMegaSteps: 1", "MCMegaSteps: 2").replace("CalculateRecombination: false", "CalculateRecombination: true")
    temps = [150, 300, 450]
    b = tmp_path / "batch"
    b.mkdir()
    (b / "starrynight.cfg").write_text(cfg + "\nTemperatures = [" + ", ".join(map(str, temps)) + "];\n")
    run = subprocess.run([DRIVER], cwd=b, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "Temperature batch: 3 replicas" in run.stderr
    for T in temps:
        d = tmp_path / f"T{T}"
        d.mkdir()
        (d / "starrynight.cfg").write_text(cfg)
        one = subprocess.run([DRIVER, str(T)], cwd=d, capture_output=True, text=True)
        assert one.returncode == 0, one.stderr[-2000:]
        tagged = [fn for fn in sorted(os.listdir(d)) if fn.startswith(f"T_{T:04d}_")]
        assert len(tagged) >= 8
        for fn in tagged:
            assert (d / fn).read_bytes() == (b / fn).read_bytes(), fn
        strip = lambda t: [l for l in t.splitlines() if not l.startswith("#")]          # header carries time(NULL)
        assert strip((d / f"Recombination_T_{T:04d}.log").read_text()) == strip((b / f"Recombination_T_{T:04d}.log").read_text())
        acc = [l for l in one.stderr.splitlines() if "ACCEPT:" in l][0]
        assert f"T: {T} {acc}" in run.stderr


def test_driver_extra_keys_host_logic(built, tmp_path):
    """Host-side handling of the B200-only cfg keys (no GPU needed: --init-only stops before sn_create)."""
    cfg = ref_files()["starrynight.cfg"].decode().replace('"antiferro_wall"', '"random"')
    # Temperatures overrides T: the first entry seeds replica 0 exactly like `T: 310` / argv[1] = 310 would
    a = run_init_only(tmp_path, cfg + "\nTemperatures = [310, 350];\n")
    b = run_init_only(tmp_path, cfg.replace("T: 300", "T: 310"))
    assert np.array_equal(a, b)
    # combinations the driver refuses, with a message and a non-zero exit code
    for extra in ("Temperatures = [300, 350];\nGPUs = 2;", 'Temperatures = [300, 350];\nCheckpoint = "c.bin";', "GPUs = 0;", "GPUs = 99;"):
        (tmp_path / "starrynight.cfg").write_text(cfg + "\n" + extra + "\n")
        bad = subprocess.run([DRIVER, "--init-only", "x.bin"], cwd=tmp_path, capture_output=True, text=True)
        assert bad.returncode != 0 and ("Temperatures" in bad.stderr or "GPUs" in bad.stderr), extra
    # a restart file that does not match the configuration is refused before any GPU work
    (tmp_path / "starrynight.cfg").write_text(cfg + '\nRestart = "nope.bin";\n')
    bad = subprocess.run([DRIVER], cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode != 0 and "Restart" in bad.stderr
