"""Generate tests/golden/*.npz from the reference's OWN code (oracle/_ref, built
from /root/reference/src by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The vectors pin both the CPU oracle (tests/test_oracle.py, no GPU) and the CUDA
path (tests/test_gpu_*.py) on boxes where /root/reference does not exist.
Every case stores its inputs, so nothing has to be regenerated to check it.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_api as oa  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (X, Y, Z, cutoff, CageStrain, K, Efield, beta, lengths, prevalence, ConstrainToX, DIM)
    "species3d": (10, 9, 12, 3, 1.3, 0.7, (0.02, -0.01, 0.03), 1.2, (1.0, 0.5, 0.0), (0.6, 0.3, 0.1), 0, 3),
    "flat2d": (12, 10, 1, 3, 1.0, 0.0, (0.02, 0.0, 0.0), 1.0, (1.0,), (1.0,), 0, 3),
    "odd_cut2": (9, 11, 10, 2, 0.5, 0.0, (0.0, 0.0, 0.0), 2.0, (1.0, 0.25), (0.5, 0.5), 0, 3),
    "cut4_constrain": (9, 10, 9, 4, 2.0, 0.3, (0.0, 0.05, 0.0), 0.8, (1.0,), (1.0,), 1, 3),
    "tiny": (4, 4, 4, 3, 1.0, 0.0, (0.0, 0.0, 0.0), 1.0, (1.0,), (1.0,), 0, 3),
    "dim2": (9, 9, 9, 3, 1.0, 0.0, (0.01, 0.0, 0.0), 1.0, (1.0,), (1.0,), 0, 2),
}


def parse_rdf(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.startswith("#")]
    return np.array([[float(v) for v in r] for r in rows])


def case_inputs(name):
    X, Y, Z, cut, cage, K, E, beta, lens, prev, constrain, dim = CASES[name]
    E = tuple(float(np.float32(v)) for v in E)
    p = oa.make_params(X, Y, Z, cut, cage, K, E, beta, constrain, dim, 300)
    lat = oa.random_lattice(X, Y, Z, seed=sum(map(ord, name)), lengths=lens, prevalence=prev)
    return p, lat


def make_extra(refs):
    """analysis_extra.npz: dipole_electricfield (cutoff 4), dipole_electricfieldoffset (cutoff 2) maps and the
    two log lines of recombination_calculator, from the reference build, on the lattices of the cases above
    (same inputs: regenerated from the case name, checked against the stored lattice by the tests)."""
    out = {}
    for name in ("species3d", "flat2d", "odd_cut2", "cut4_constrain", "dim2"):
        p, lat = case_inputs(name)
        for prec, r in refs.items():
            r.configure(p)
            r.set_lattice(lat)
            out[f"{name}_efield_{prec}"] = r.efield_map(4, False)
            out[f"{name}_efieldoffset_{prec}"] = r.efield_map(2, True)
            with tempfile.NamedTemporaryFile(suffix=".log", delete=False) as f:
                path = f.name
            txt = r.recombination_log(path)
            os.unlink(path)
            out[f"{name}_recombination_{prec}"] = np.frombuffer(txt.encode(), np.uint8)
        print(name, txt.strip()[:100])
    np.savez_compressed(os.path.join(HERE, "analysis_extra.npz"), **out)


def main():
    refs = {prec: oa.RefLib(prec) for prec in ("f32", "f64")}
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        make_extra(refs)
        return
    make_extra(refs)
    for name, (X, Y, Z, cut, cage, K, E, beta, lens, prev, constrain, dim) in CASES.items():
        E = tuple(float(np.float32(v)) for v in E)   # Efield is a float in the reference and in the C ABI
        p = oa.make_params(X, Y, Z, cut, cage, K, E, beta, constrain, dim, 300)
        lat = oa.random_lattice(X, Y, Z, seed=sum(map(ord, name)), lengths=lens, prevalence=prev)
        rng = np.random.default_rng(11)
        n = 96
        sites = np.stack([rng.integers(0, X, n), rng.integers(0, Y, n), rng.integers(0, Z, n)], 1).astype(np.int32)
        nd = rng.normal(size=(n, 3))
        nd /= np.linalg.norm(nd, axis=1, keepdims=True)
        nd = nd.astype(np.float32)
        out = dict(params=np.array([X, Y, Z, cut, constrain, dim]), couplings=np.array([cage, K, *E, beta]),
                   lattice=lat, sites=sites, newdip=nd)
        for prec, r in refs.items():
            r.configure(p)
            r.set_lattice(lat)
            dxyz, d = r.neighbours()
            out[f"nb_dxyz_{prec}"] = dxyz
            out[f"nb_d_{prec}"] = d
            out[f"dE_{prec}"] = r.site_energy(sites, nd)
            out[f"interaction_{prec}"] = r.site_interaction_map()
            out[f"total_{prec}"] = r.total_energy()
            out[f"polarisation_{prec}"] = np.array(r.polarisation())
            out[f"landau_{prec}"] = np.array(r.landau_order())
            # the reference indexes lattice[(X+x+dx)%X] with |dx| up to 6 (potential) and 9 (RDF): extents
            # below that read out of bounds in the reference itself, so those cases carry no observable vectors
            small = min(X, Y, Z if Z > 1 else 99)
            if small >= 6:
                out[f"potential_{prec}"] = r.potential_map()
            if small >= 9:
                with tempfile.NamedTemporaryFile(suffix=".dat", delete=False) as f:
                    path = f.name
                os.unlink(path)
                r.rdf_file(path)
                out[f"rdf_{prec}"] = parse_rdf(path)
                os.unlink(path)
            # the reference chain: seed as main.c:172 does, 4000 attempts
            r.seed(0xDEADBEEF + 300)
            acc, rej = r.mc_moves(4000)
            out[f"chain_counters_{prec}"] = np.array([acc, rej], np.int64)
            out[f"chain_lattice_{prec}"] = r.get_lattice().astype(np.float32 if prec == "f32" else np.float64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "nb", len(out["nb_d_f32"]), "accept", out["chain_counters_f32"])

    # initial lattices + solid solution for the stock `make test` geometry (starrynight.cfg:10-15)
    X, Y, Z = 20, 20, 28
    p = oa.make_params(X, Y, Z)
    r = refs["f32"]
    r.configure(p)
    init = {}
    for kind in ("random", "ferroelectric", "buckled", "antiferro_wall", "ferro_wall", "antiferro_slip", "spectrum"):
        r.seed(0xDEADBEEF + 300)
        r.initialise_lattice(kind)
        r.solid_solution([1.0, 0.0, 0.0], [1.0, 0.0, 0.0])          # starrynight.cfg:45-46
        init[kind] = r.get_lattice().astype(np.float32)
    r.seed(0xDEADBEEF + 300)
    r.initialise_lattice("random")
    r.solid_solution([1.0, 0.5, 0.0], [0.6, 0.3, 0.1])
    init["random_mixed"] = r.get_lattice().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "initial_lattices.npz"), **init)

    # MT19937 known answers through the reference's generator
    r.seed(5489)
    mt = np.array([r.lib.ref_genrand_int32() for _ in range(16)], np.uint64)
    r.seed(0xDEADBEEF + 300)
    re1 = np.array([r.lib.ref_genrand_real1() for _ in range(8)])
    re2 = np.array([r.lib.ref_genrand_real2() for _ in range(8)])
    np.savez_compressed(os.path.join(HERE, "mt19937.npz"), int32_seed5489=mt, real1=re1, real2=re2)
    # the reference PROGRAM on a shortened copy of its own starrynight.cfg: the files its main() writes
    import shutil
    import subprocess
    work = tempfile.mkdtemp()
    cfg = open(os.path.join(oa.REF_ROOT, "starrynight.cfg")).read()
    cfg = cfg.replace("MCMegaSteps: 20", "MCMegaSteps: 1").replace("MCEqmSteps: 5", "MCEqmSteps: 1").replace("MCMoves: 200.0", "MCMoves: 2.0")
    cfg = cfg.replace("DisplayDumbTerminal: true", "DisplayDumbTerminal: false")
    open(os.path.join(work, "starrynight.cfg"), "w").write(cfg)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "starrynight_ref")], cwd=work, check=True, capture_output=True)
    files = {}
    for fn in sorted(os.listdir(work)):
        if fn.startswith("Recombination"):
            continue                     # carries time(NULL)
        files[fn] = np.frombuffer(open(os.path.join(work, fn), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_run_files.npz"), **files)
    print("reference run files:", ", ".join(f"{k} ({len(v)} B)" for k, v in files.items()))
    shutil.rmtree(work)
    print("done")


if __name__ == "__main__":
    main()
