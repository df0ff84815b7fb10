"""TEST INFRASTRUCTURE -- ctypes access to the CPU checkers.

* ``Oracle``  wraps ``oracle/libsn_oracle.so`` (our C restatement, sn_oracle.c).
* ``RefLib``  wraps ``oracle/_ref/libref_f32.so`` / ``libref_f64.so`` (the
  reference's unmodified sources built by oracle/Makefile).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``starrynight_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


def build(quiet: bool = True) -> None:
    """Compile the restatement, and the reference libraries when the tree exists."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class Params(C.Structure):
    _fields_ = [("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int), ("cutoff", C.c_int),
                ("CageStrain", C.c_double), ("K", C.c_double), ("Efield", C.c_double * 3),
                ("beta", C.c_double), ("ConstrainToX", C.c_int), ("DIM", C.c_int), ("T", C.c_int)]


class MT(C.Structure):
    _fields_ = [("mt", C.c_ulong * 624), ("left", C.c_int), ("next", C.c_int)]


def make_params(X, Y, Z, cutoff=3, CageStrain=1.0, K=0.0, Efield=(0.0, 0.0, 0.0), beta=1.0,
                ConstrainToX=0, DIM=3, T=300) -> Params:
    p = Params()
    p.X, p.Y, p.Z, p.cutoff = X, Y, Z, cutoff
    p.CageStrain, p.K = CageStrain, K
    for i in range(3):
        p.Efield[i] = float(Efield[i])
    p.beta, p.ConstrainToX, p.DIM, p.T = beta, int(ConstrainToX), DIM, T
    return p


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Our restatement.  ``prec`` is 'f32' (native reference build) or 'f64'."""

    def __init__(self, prec: str = "f32"):
        path = os.path.join(HERE, "libsn_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.prec = prec
        self.dtype = np.float32 if prec == "f32" else np.float64
        self.ct = C.c_float if prec == "f32" else C.c_double
        self.lib.sno_mt_int32.restype = C.c_ulong
        self.lib.sno_mt_real1.restype = C.c_double
        self.lib.sno_mt_real2.restype = C.c_double
        for name in ("polarisation", "landau_order", "dipole_potential", "dipole_electricfield"):
            getattr(self.lib, f"sno_{name}_{prec}").restype = C.c_double

    def fn(self, name):
        return getattr(self.lib, f"sno_{name}_{self.prec}")

    # --- rng
    def mt(self, seed: int) -> MT:
        s = MT()
        self.lib.sno_mt_seed(C.byref(s), C.c_ulong(seed & 0xFFFFFFFF))
        return s

    def neighbours(self, p: Params):
        dxyz = np.zeros((10000, 3), np.int32)
        d = np.zeros(10000, np.float64)
        n = self.lib.sno_gen_neighbours(C.byref(p), _ptr(dxyz, C.c_int), _ptr(d, C.c_double))
        return dxyz[:n].copy(), d[:n].copy()

    def _lat(self, lat):
        a = np.ascontiguousarray(lat, dtype=self.dtype)
        return a

    def site_energy(self, p, lat, sites, newdip):
        lat = self._lat(lat)
        sites = np.ascontiguousarray(sites, np.int32)
        nd = np.ascontiguousarray(newdip, self.dtype)
        out = np.zeros(len(sites), np.float64)
        self.fn("site_energy_batch")(C.byref(p), _ptr(lat, self.ct), len(sites), _ptr(sites, C.c_int),
                                     _ptr(nd, self.ct), _ptr(out, C.c_double))
        return out

    def site_interaction_map(self, p, lat):
        lat = self._lat(lat)
        out = np.zeros(p.X * p.Y * p.Z, np.float64)
        self.fn("site_interaction_map")(C.byref(p), _ptr(lat, self.ct), _ptr(out, C.c_double))
        return out

    def total_energy(self, p, lat):
        lat = self._lat(lat)
        out = np.zeros(4, np.float64)
        self.fn("total_energy")(C.byref(p), _ptr(lat, self.ct), _ptr(out, C.c_double))
        return out

    def mc_moves(self, p, lat, mt, moves):
        """In-place on ``lat`` (must already be a contiguous array of self.dtype)."""
        assert lat.dtype == self.dtype and lat.flags.c_contiguous
        acc, rej = C.c_ulonglong(0), C.c_ulonglong(0)
        self.fn("mc_moves")(C.byref(p), _ptr(lat, self.ct), C.byref(mt), C.c_longlong(moves),
                            C.byref(acc), C.byref(rej))
        return acc.value, rej.value

    def initialise_lattice(self, p, mt, name):
        lat = np.zeros((p.X, p.Y, p.Z, 4), self.dtype)
        ok = self.fn("initialise_lattice")(C.byref(p), _ptr(lat, self.ct), C.byref(mt), name.encode())
        if not ok:
            raise ValueError(name)
        return lat

    def solid_solution(self, p, lat, mt, lengths, prevalence):
        ln = np.ascontiguousarray(lengths, np.float64)
        pv = np.ascontiguousarray(prevalence, np.float64)
        histo = np.zeros(10, np.int32)
        self.fn("solid_solution")(C.byref(p), _ptr(lat, self.ct), C.byref(mt), len(ln),
                                  _ptr(ln, C.c_double), _ptr(pv, C.c_double), _ptr(histo, C.c_int))
        return histo[:len(ln)]

    def polarisation(self, p, lat):
        lat = self._lat(lat)
        return self.fn("polarisation")(C.byref(p), _ptr(lat, self.ct))

    def landau_order(self, p, lat):
        lat = self._lat(lat)
        return self.fn("landau_order")(C.byref(p), _ptr(lat, self.ct))

    def potential_map(self, p, lat):
        lat = self._lat(lat)
        out = np.zeros(p.X * p.Y * p.Z, np.float64)
        self.fn("potential_map")(C.byref(p), _ptr(lat, self.ct), _ptr(out, C.c_double))
        return out

    def efield_map(self, p, lat, cutoff=4, half_offset=False):
        lat = self._lat(lat)
        out = np.zeros(p.X * p.Y * p.Z, np.float64)
        self.fn("efield_map")(C.byref(p), _ptr(lat, self.ct), int(cutoff), int(half_offset), _ptr(out, C.c_double))
        return out

    def recombination(self, p, lat):
        """ZBe ZBh ZFDe ZFDh R_Boltz R_FD e_total h_total eMAX hMAX RMAX (analysis.c:96-170)."""
        lat = self._lat(lat)
        out = np.zeros(11, np.float64)
        self.fn("recombination")(C.byref(p), _ptr(lat, self.ct), _ptr(out, C.c_double))
        return out

    def rdf(self, p, lat):
        """Accumulated (fe_sum, afe_sum, count) for r^2 = 0..80, before division."""
        lat = self._lat(lat)
        fe = np.zeros(81, self.dtype)
        afe = np.zeros(81, self.dtype)
        cnt = np.zeros(81, np.int32)
        self.fn("rdf")(C.byref(p), _ptr(lat, self.ct), _ptr(fe, self.ct), _ptr(afe, self.ct), _ptr(cnt, C.c_int))
        return fe, afe, cnt


def ref_available(prec: str = "f32") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libref_{prec}.so"))


class RefLib:
    """The reference's own code.  State is global inside the library (as in the
    reference), so use one instance per precision per process."""

    def __init__(self, prec: str = "f32"):
        path = os.path.join(HERE, "_ref", f"libref_{prec}.so")
        if not os.path.exists(path):
            if os.path.isdir(REF_ROOT):
                build()
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.prec = prec
        L = self.lib
        L.ref_configure.argtypes = [C.c_int] * 4 + [C.c_double] * 6 + [C.c_int] * 3
        L.ref_set_beta.argtypes = [C.c_double]
        L.ref_set_efield.argtypes = [C.c_double] * 3
        L.ref_set_cagestrain.argtypes = [C.c_double]
        L.ref_set_K.argtypes = [C.c_double]
        L.ref_seed.argtypes = [C.c_ulong]
        L.ref_mc_moves.argtypes = [C.c_int]
        L.ref_genrand_int32.restype = C.c_ulong
        for n in ("ref_genrand_real1", "ref_genrand_real2", "ref_polarisation", "ref_landau_order", "ref_dipole_potential",
                  "ref_dipole_electricfield"):
            getattr(L, n).restype = C.c_double
        L.ref_dipole_electricfield.argtypes = [C.c_int] * 4
        L.ref_dipole_potential.argtypes = [C.c_int] * 3
        self.p = None

    def configure(self, p: Params):
        self.p = p
        self.lib.ref_configure(p.X, p.Y, p.Z, p.cutoff, p.CageStrain, p.K, p.Efield[0], p.Efield[1], p.Efield[2],
                               p.beta, p.ConstrainToX, p.DIM, p.T)

    @property
    def n(self):
        return self.p.X * self.p.Y * self.p.Z

    def set_lattice(self, lat):
        a = np.ascontiguousarray(lat, np.float64)
        assert a.size == self.n * 4
        self.lib.ref_set_lattice(_ptr(a, C.c_double))

    def get_lattice(self):
        a = np.zeros((self.p.X, self.p.Y, self.p.Z, 4), np.float64)
        self.lib.ref_get_lattice(_ptr(a, C.c_double))
        return a

    def neighbours(self):
        n = self.lib.ref_neighbour_count()
        dxyz = np.zeros((n, 3), np.int32)
        d = np.zeros(n, np.float64)
        self.lib.ref_get_neighbours(_ptr(dxyz, C.c_int), _ptr(d, C.c_double))
        return dxyz, d

    def site_energy(self, sites, newdip):
        sites = np.ascontiguousarray(sites, np.int32)
        nd = np.ascontiguousarray(newdip, np.float64)
        out = np.zeros(len(sites), np.float64)
        self.lib.ref_site_energy_batch(len(sites), _ptr(sites, C.c_int), _ptr(nd, C.c_double), _ptr(out, C.c_double))
        return out

    def site_interaction_map(self):
        out = np.zeros(self.n, np.float64)
        self.lib.ref_site_interaction_map(_ptr(out, C.c_double))
        return out

    def total_energy(self):
        """Same recipe as sno_total_energy, through the reference's site_energy."""
        p = self.p
        lat = self.get_lattice()
        out = np.zeros(4)

        def run(cutoff, cage, K, E):
            q = make_params(p.X, p.Y, p.Z, cutoff, cage, K, E, p.beta, p.ConstrainToX, p.DIM, p.T)
            self.configure(q)
            self.set_lattice(lat)
            return self.site_interaction_map().sum()

        E = tuple(p.Efield)
        out[0] = 0.5 * run(p.cutoff, 0.0, 0.0, (0, 0, 0))
        out[1] = 0.5 * (run(1, p.CageStrain, 0.0, (0, 0, 0)) - run(1, 0.0, 0.0, (0, 0, 0)))
        out[2] = run(0, 0.0, 0.0, E)
        out[3] = run(0, 0.0, p.K, (0, 0, 0))
        self.configure(p)
        self.set_lattice(lat)
        return out

    def seed(self, s):
        self.lib.ref_seed(C.c_ulong(s & 0xFFFFFFFF))

    def mc_moves(self, n):
        self.lib.ref_reset_counters()
        self.lib.ref_mc_moves(int(n))
        a, r = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.ref_get_counters(C.byref(a), C.byref(r))
        return a.value, r.value

    def initialise_lattice(self, name):
        if not self.lib.ref_initialise_lattice(name.encode()):
            raise ValueError(name)

    def solid_solution(self, lengths, prevalence):
        ln = np.ascontiguousarray(lengths, np.float64)
        pv = np.ascontiguousarray(prevalence, np.float64)
        self.lib.ref_solid_solution(len(ln), _ptr(ln, C.c_double), _ptr(pv, C.c_double))

    def polarisation(self):
        return self.lib.ref_polarisation()

    def landau_order(self):
        return self.lib.ref_landau_order()

    def potential_map(self):
        out = np.zeros(self.n, np.float64)
        self.lib.ref_potential_map(_ptr(out, C.c_double))
        return out

    def rdf_file(self, path):
        self.lib.ref_radial_order_parameter(path.encode())

    def efield_map(self, cutoff=4, half_offset=False):
        out = np.zeros(self.n, np.float64)
        self.lib.ref_efield_map(int(cutoff), int(half_offset), _ptr(out, C.c_double))
        return out

    def recombination_log(self, path):
        """The two lines recombination_calculator writes to its log (analysis.c:129-131,169-171)."""
        self.lib.ref_recombination(path.encode())
        return open(path).read()


def random_lattice(X, Y, Z, seed=0, lengths=(1.0,), prevalence=(1.0,), dtype=np.float32):
    """Seeded unit dipoles (float32-representable) with species lengths; the
    synthetic input used across tests and bench."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(X, Y, Z, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    lat = np.zeros((X, Y, Z, 4), np.float32)
    lat[..., :3] = v.astype(np.float32)
    u = rng.random((X, Y, Z))
    edges = np.cumsum(prevalence)
    idx = np.minimum(np.searchsorted(edges, u, side="left"), len(lengths) - 1)
    lat[..., 3] = np.asarray(lengths, np.float32)[idx]
    return lat.astype(dtype)
