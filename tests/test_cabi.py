"""CPU tests of the drop-in boundary: the shared library loads, exports exactly the
entry points include/starrynight_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "starrynight_b200.h")).read()
    return sorted(set(re.findall(r"SN_API\s+[\w\s\*]+?\b(sn_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("sn_create", "sn_destroy", "sn_set_lattice", "sn_get_lattice", "sn_mc_sweeps", "sn_site_energy",
                 "sn_total_energy", "sn_polarisation", "sn_landau_order", "sn_rdf", "sn_potential_map", "sn_efield_map", "sn_recombination",
                 "sn_get_counters", "sn_ipc_export", "sn_ipc_attach"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    import starrynight_b200 as sn
    lib = sn.load_library()
    syms = declared_symbols()
    assert len(syms) >= 27
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(sn.EXPORTS) == syms, "python binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", sn.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("sn_"))
    assert exported == syms, "library exports differ from the header"


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "starrynight_b200.h"\nint main(void){ sn_params p; (void)p; return sizeof(sn_params) > 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layout_matches_binding(built, tmp_path):
    """sizeof/offsetof of sn_params as the C compiler sees it vs the ctypes mirror."""
    import starrynight_b200 as sn
    src = tmp_path / "s.c"
    fields = [f[0] for f in sn.sn_params._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(sn_params, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "starrynight_b200.h"\n'
                   f'int main(void){{ printf("%zu\\n", sizeof(sn_params)); {body} return 0; }}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(sn.sn_params)
    assert vals[1:] == [getattr(sn.sn_params, f).offset for f in fields]


def test_no_cpu_fallback(built):
    """Without a CUDA device sn_create must fail with a message, never compute on the CPU."""
    import torch
    import starrynight_b200 as sn
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sn.SnError, match="no CUDA device"):
        sn.Simulation(8, 8, 8)


def test_product_never_touches_the_oracle():
    """The product path may not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "starrynight_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sn_oracle" not in text and "oracle_api" not in text and "libref_" not in text, os.path.join(dirpath, f)
    drv = os.path.join(ROOT, "driver")
    for f in os.listdir(drv):
        text = open(os.path.join(drv, f), errors="ignore").read()
        assert "sn_oracle" not in text and "libref_" not in text, f


def test_philox_known_answers_host_side(built):
    """the counter-based generator against the Random123 known-answer vectors (host compilation
    of the same __host__ __device__ function the kernels call; the device side is checked in test_gpu_audit.py)
    and against an independent numpy implementation."""
    import numpy as np
    import starrynight_b200 as sn
    from tests.helpers import PHILOX_KAT, philox4x32_10
    host, _ = sn.philox_kat([list(c) + list(k) for c, k, _ in PHILOX_KAT])
    assert np.array_equal(host, np.array([o for _, _, o in PHILOX_KAT], np.uint32))
    rnd = np.random.default_rng(0).integers(0, 2 ** 32, size=(1000, 6), dtype=np.uint64).astype(np.uint32)
    host, _ = sn.philox_kat(rnd)
    assert np.array_equal(host, np.stack(philox4x32_10(*[rnd[:, i] for i in range(6)]), 1))
