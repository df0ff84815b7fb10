"""starrynight_b200 -- host-side mirror of libstarrynight_b200.so (ctypes).

The product is the C-ABI library (include/starrynight_b200.h) and the C driver
(driver/); this module is the thin Python binding the tests and bench.py use.
It mirrors the reference's operations for the hot path with the reference's
names (MC_moves, site_energy, polarisation, landau_order,
radial_order_parameter, dipole_potential) so the parity tests read like calls
into /root/reference/src/starrynight-*.c.

There is no CPU path here: loading fails loudly when the shared library has not
been built, and ``Simulation(...)`` fails loudly when no CUDA device is usable.
Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SN_B200_LIB", os.path.join(_HERE, "libstarrynight_b200.so"))   # override: kernel experiments only

SN_PREC_F32, SN_PREC_F64, SN_PREC_REPLICA = 0, 1, 2
SN_KERNEL_AUTO, SN_KERNEL_COLOUR, SN_KERNEL_TILED, SN_KERNEL_TILED_PHASED, SN_KERNEL_RESIDENT = 0, 1, 2, 3, 4
SN_RDF_BINS = 81

EXPORTS = [
    "sn_last_error", "sn_version", "sn_device_count", "sn_default_params", "sn_create", "sn_destroy", "sn_neighbour_table",
    "sn_set_lattice", "sn_get_lattice", "sn_set_lattice_async", "sn_get_lattice_async", "sn_set_lattices_async", "sn_get_lattices_async", "sn_order_after", "sn_pull_ghosts", "sn_set_beta", "sn_set_efield", "sn_set_cagestrain", "sn_set_replica_cagestrain", "sn_mc_sweeps", "sn_mc_sweep_audit",
    "sn_mc_sweeps_timed", "sn_synchronize", "sn_get_counters", "sn_reset_counters", "sn_set_counters",
    "sn_get_sweep_count", "sn_set_sweep_count", "sn_set_replica_seed", "sn_site_energy",
    "sn_total_energy", "sn_polarisation", "sn_landau_order", "sn_rdf", "sn_potential_map", "sn_efield_map", "sn_recombination", "sn_recombination_partial", "sn_recombination_finish", "sn_get_boundary",
    "sn_set_ghost", "sn_ipc_export", "sn_ipc_attach", "sn_attach_peer", "sn_bench_fp32_peak", "sn_bench_fp64_peak", "sn_philox_kat", "sn_state_hash", "sn_kernel_in_use", "sn_tile_schedule",
]


class SnError(RuntimeError):
    pass


class sn_params(C.Structure):
    _fields_ = [("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int), ("cutoff", C.c_int),
                ("CageStrain", C.c_double), ("K", C.c_double), ("Efield", C.c_float * 3),
                ("beta", C.c_double), ("ConstrainToX", C.c_int), ("DIM", C.c_int),
                ("nreplicas", C.c_int), ("seed", C.c_ulonglong), ("device", C.c_int),
                ("z0", C.c_int), ("nz", C.c_int), ("kernel", C.c_int)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SnError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(make -C starrynight_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.sn_last_error.restype = C.c_char_p
    lib.sn_version.restype = C.c_char_p
    H = C.c_void_p
    lib.sn_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.sn_create.argtypes = [C.POINTER(sn_params), C.POINTER(H)]
    lib.sn_destroy.argtypes = [H]
    lib.sn_neighbour_table.argtypes = [H, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    lib.sn_set_lattice.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_get_lattice.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_set_lattice_async.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_get_lattice_async.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_set_lattices_async.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.sn_get_lattices_async.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.sn_order_after.argtypes = [H, H]
    lib.sn_pull_ghosts.argtypes = [H]
    lib.sn_set_beta.argtypes = [H, C.c_int, C.c_double]
    lib.sn_set_efield.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    lib.sn_set_cagestrain.argtypes = [H, C.c_double]
    lib.sn_set_replica_cagestrain.argtypes = [H, C.c_int, C.c_double]
    lib.sn_mc_sweeps.argtypes = [H, C.c_longlong]
    lib.sn_mc_sweep_audit.argtypes = [H, C.c_void_p]
    lib.sn_mc_sweeps_timed.argtypes = [H, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.sn_synchronize.argtypes = [H]
    lib.sn_get_counters.argtypes = [H, C.c_int] + [C.POINTER(C.c_ulonglong)] * 3
    lib.sn_reset_counters.argtypes = [H]
    lib.sn_site_energy.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sn_total_energy.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.sn_polarisation.argtypes = [H, C.c_int, C.POINTER(C.c_double)]
    lib.sn_landau_order.argtypes = [H, C.c_int, C.POINTER(C.c_double)]
    lib.sn_rdf.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sn_set_counters.argtypes = [H, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]
    lib.sn_set_replica_seed.argtypes = [H, C.c_int, C.c_ulonglong]
    lib.sn_get_sweep_count.argtypes = [H, C.POINTER(C.c_ulonglong)]
    lib.sn_set_sweep_count.argtypes = [H, C.c_ulonglong]
    lib.sn_potential_map.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_efield_map.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.sn_recombination.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_recombination_partial.argtypes = [H, C.c_int, C.c_void_p]
    lib.sn_recombination_finish.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    lib.sn_get_boundary.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.sn_set_ghost.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.sn_ipc_export.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.sn_ipc_attach.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    lib.sn_attach_peer.argtypes = [H, C.c_int, H]
    lib.sn_bench_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.sn_bench_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.sn_philox_kat.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sn_state_hash.argtypes = [H, C.c_int, C.POINTER(C.c_ulonglong)]
    lib.sn_kernel_in_use.argtypes = [H, C.POINTER(C.c_int)]
    _lib = lib
    return lib


def default_params() -> sn_params:
    p = sn_params()
    _check(load_library().sn_default_params(C.byref(p)))
    return p


def _check(rc: int) -> None:
    if rc != 0:
        raise SnError(f"libstarrynight_b200 error {rc}: {load_library().sn_last_error().decode()}")


def philox_kat(counter_key, device=False):
    """sn_philox4x32_10 for rows of (counter[4], key[2]): returns (host words, device words or None)."""
    ck = np.ascontiguousarray(counter_key, np.uint32).reshape(-1, 6)
    host = np.zeros((len(ck), 4), np.uint32)
    dev = np.zeros((len(ck), 4), np.uint32) if device else None
    _check(load_library().sn_philox_kat(len(ck), ck.ctypes.data, host.ctypes.data, dev.ctypes.data if device else None))
    return host, dev


def tile_schedule(X, Y, Z, nreplicas=1, sweep=0):
    """Work order of the tiled kernel for one sweep (host logic only): int array [n][5] = replica, tx, ty, tz, phase."""
    lib = load_library()
    n = C.c_int(0)
    lib.sn_tile_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_int), C.c_void_p, C.c_int]
    _check(lib.sn_tile_schedule(X, Y, Z, nreplicas, sweep, C.byref(n), None, 0))
    items = np.zeros((n.value, 5), np.int32)
    _check(lib.sn_tile_schedule(X, Y, Z, nreplicas, sweep, C.byref(n), items.ctypes.data, n.value))
    return items


def recombination_finish(parts):
    """Merge sn_recombination_partial results of the slabs of one lattice (sn_recombination_finish)."""
    p = np.ascontiguousarray(parts, np.float64).reshape(-1, 9)
    out = np.zeros(11, np.float64)
    _check(load_library().sn_recombination_finish(len(p), p.ctypes.data, out.ctypes.data))
    return out


def beta_of_T(T: float) -> float:
    """beta = 1/((float)T/300.0), main.c:215 (T = 0 gives +inf, as in the reference)."""
    t = float(np.float32(T)) / 300.0
    return float("inf") if t == 0.0 else 1.0 / t


class Simulation:
    """One device-resident lattice (or a batch of replicas, or one Z-slab of a
    larger lattice).  Method names follow the reference functions they replace."""

    def __init__(self, X, Y, Z, DipoleCutOff=3, CageStrain=1.0, K=0.0, Efield=(0.0, 0.0, 0.0), beta=1.0,
                 ConstrainToX=False, DIM=3, nreplicas=1, seed=0xDEADBEEF + 300, device=0, z0=0, nz=0,
                 kernel=SN_KERNEL_AUTO):
        self.lib = load_library()
        p = default_params()
        p.X, p.Y, p.Z, p.cutoff = X, Y, Z, DipoleCutOff
        p.CageStrain, p.K, p.beta = CageStrain, K, beta
        for i in range(3):
            p.Efield[i] = float(Efield[i])
        p.ConstrainToX, p.DIM, p.nreplicas = int(ConstrainToX), DIM, nreplicas
        p.seed, p.device, p.z0, p.nz, p.kernel = seed, device, z0, nz, kernel
        self.params = p
        self.X, self.Y, self.Z = X, Y, Z
        self.nz = nz if nz > 0 else Z
        self.nreplicas = nreplicas
        self.h = C.c_void_p()
        _check(self.lib.sn_create(C.byref(p), C.byref(self.h)))

    # -- lifetime
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.sn_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def nsites(self):
        return self.X * self.Y * self.nz

    # -- gen_neighbour (montecarlo-core.c:38)
    def neighbours(self):
        n = C.c_int(0)
        _check(self.lib.sn_neighbour_table(self.h, C.byref(n), None, None))
        dxyz = np.zeros((n.value, 3), np.int32)
        d = np.zeros(n.value, np.float32)
        _check(self.lib.sn_neighbour_table(self.h, C.byref(n), dxyz.ctypes.data, d.ctypes.data))
        return dxyz, d

    # -- lattice[x][y][z] (config.c:32-36)
    def set_lattice(self, lat, replica=0):
        a = np.ascontiguousarray(lat, np.float32)
        if a.size != self.nsites * 4:
            raise SnError(f"set_lattice: expected {self.X}x{self.Y}x{self.nz}x4 floats, got {a.shape}")
        _check(self.lib.sn_set_lattice(self.h, replica, a.ctypes.data))

    def get_lattice(self, replica=0, out=None):
        a = out if out is not None else np.empty((self.X, self.Y, self.nz, 4), np.float32)
        _check(self.lib.sn_get_lattice(self.h, replica, a.ctypes.data))
        return a

    def set_lattice_ptr(self, ptr, replica=0):
        """Host pointer variant (e.g. a pinned torch tensor's data_ptr())."""
        _check(self.lib.sn_set_lattice(self.h, replica, C.c_void_p(ptr)))

    def get_lattice_ptr(self, ptr, replica=0):
        _check(self.lib.sn_get_lattice(self.h, replica, C.c_void_p(ptr)))

    def set_lattice_async(self, ptr, replica=0):
        """Queue the upload from pinned host memory at `ptr`; the buffer must stay valid until synchronize()."""
        _check(self.lib.sn_set_lattice_async(self.h, replica, C.c_void_p(ptr)))

    def get_lattice_async(self, ptr, replica=0):
        _check(self.lib.sn_get_lattice_async(self.h, replica, C.c_void_p(ptr)))

    def set_lattices_async(self, ptr, first=0, count=None):
        """Queue the upload of `count` consecutive replicas from one dense pinned block float[count][X][Y][nz][4]."""
        _check(self.lib.sn_set_lattices_async(self.h, first, self.nreplicas - first if count is None else count, C.c_void_p(ptr)))

    def get_lattices_async(self, ptr, first=0, count=None):
        _check(self.lib.sn_get_lattices_async(self.h, first, self.nreplicas - first if count is None else count, C.c_void_p(ptr)))

    def order_after(self, other: "Simulation"):
        _check(self.lib.sn_order_after(self.h, other.h))

    def pull_ghosts(self):
        _check(self.lib.sn_pull_ghosts(self.h))

    def set_beta(self, beta, replica=0):
        _check(self.lib.sn_set_beta(self.h, replica, float(beta)))

    def set_T(self, T, replica=0):
        self.set_beta(beta_of_T(T), replica)

    def set_efield(self, E, replica=0):
        e = (C.c_float * 3)(*[float(v) for v in E])
        _check(self.lib.sn_set_efield(self.h, replica, e))

    def set_cagestrain(self, c, replica=None):
        """CageStrain of every replica, or of one (the reference's T x CageStrain grid, Makefile:52-54)."""
        if replica is None:
            _check(self.lib.sn_set_cagestrain(self.h, float(c)))
        else:
            _check(self.lib.sn_set_replica_cagestrain(self.h, int(replica), float(c)))

    # -- MC_moves (montecarlo-core.c:143): one sweep = X*Y*Z attempts
    def MC_sweeps(self, nsweeps=1):
        _check(self.lib.sn_mc_sweeps(self.h, int(nsweeps)))

    def MC_sweeps_timed(self, nsweeps=1):
        ms, n = C.c_double(0), C.c_longlong(0)
        _check(self.lib.sn_mc_sweeps_timed(self.h, int(nsweeps), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def synchronize(self):
        _check(self.lib.sn_synchronize(self.h))

    def MC_sweep_audit(self):
        """One sweep with every attempt recorded: array [replica][x][y][z][8] =
        (trial x, y, z, accept uniform, dE, decision 1/0/2=vacant, group ordinal, 0) -- sn_mc_sweep_audit."""
        rec = np.zeros((self.nreplicas, self.X, self.Y, self.nz, 8), np.float32)
        _check(self.lib.sn_mc_sweep_audit(self.h, rec.ctypes.data))
        return rec

    def counters(self, replica=0):
        a, r, v = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
        _check(self.lib.sn_get_counters(self.h, replica, C.byref(a), C.byref(r), C.byref(v)))
        return a.value, r.value, v.value

    def reset_counters(self):
        _check(self.lib.sn_reset_counters(self.h))

    # -- site_energy (montecarlo-core.c:76)
    def site_energy(self, sites, newdip, precision=SN_PREC_F64, replica=0):
        s = np.ascontiguousarray(sites, np.int32).reshape(-1, 3)
        nd = np.ascontiguousarray(newdip, np.float32).reshape(-1, 3)
        if len(s) != len(nd):
            raise SnError("site_energy: sites and newdip differ in length")
        out = np.zeros(len(s), np.float64)
        _check(self.lib.sn_site_energy(self.h, replica, precision, len(s), s.ctypes.data, nd.ctypes.data, out.ctypes.data))
        return out

    def total_energy(self, precision=SN_PREC_F64, replica=0):
        out = (C.c_double * 4)()
        _check(self.lib.sn_total_energy(self.h, replica, precision, out))
        return np.array(list(out))

    # -- analysis.c
    def polarisation(self, replica=0):
        out = (C.c_double * 3)()
        _check(self.lib.sn_polarisation(self.h, replica, out))
        return np.array(list(out))

    def landau_order(self, replica=0):
        out = C.c_double(0)
        _check(self.lib.sn_landau_order(self.h, replica, C.byref(out)))
        return out.value

    def radial_order_parameter(self, replica=0):
        fe = np.zeros(SN_RDF_BINS, np.float64)
        afe = np.zeros(SN_RDF_BINS, np.float64)
        cnt = np.zeros(SN_RDF_BINS, np.int64)
        _check(self.lib.sn_rdf(self.h, replica, fe.ctypes.data, afe.ctypes.data, cnt.ctypes.data))
        return fe, afe, cnt

    def dipole_potential(self, replica=0, out=None):
        """Potential map of the handle's sites (analysis.c:65-94); `out`: a C-contiguous float64 array of nsites (e.g. pinned)."""
        v = np.zeros(self.nsites, np.float64) if out is None else out
        assert v.dtype == np.float64 and v.size == self.nsites and v.flags.c_contiguous
        _check(self.lib.sn_potential_map(self.h, replica, v.ctypes.data))
        return v.reshape(self.X, self.Y, self.nz)

    def dipole_electricfield(self, cutoff=4, half_offset=False, replica=0, out=None):
        """|E| per site: dipole_electricfield / dipole_electricfieldoffset (analysis.c:310-465)."""
        v = np.zeros(self.nsites, np.float64) if out is None else out
        assert v.dtype == np.float64 and v.size == self.nsites and v.flags.c_contiguous
        _check(self.lib.sn_efield_map(self.h, replica, int(cutoff), int(half_offset), v.ctypes.data))
        return v.reshape(self.X, self.Y, self.nz)

    def recombination(self, replica=0):
        """recombination_calculator (analysis.c:96-170): ZBe ZBh ZFDe ZFDh R_Boltz R_FD e_total h_total eMAX hMAX RMAX."""
        v = np.zeros(11, np.float64)
        _check(self.lib.sn_recombination(self.h, replica, v.ctypes.data))
        return v

    def state_hash(self, replica=0):
        """Position-keyed 64-bit hash of this handle's sites; slabs' hashes add (mod 2^64) to the full lattice's."""
        v = C.c_ulonglong(0)
        _check(self.lib.sn_state_hash(self.h, replica, C.byref(v)))
        return v.value

    def kernel_in_use(self):
        v = C.c_int(0)
        _check(self.lib.sn_kernel_in_use(self.h, C.byref(v)))
        return v.value

    def recombination_partial(self, replica=0):
        v = np.zeros(9, np.float64)
        _check(self.lib.sn_recombination_partial(self.h, replica, v.ctypes.data))
        return v

    def set_replica_seed(self, seed, replica):
        _check(self.lib.sn_set_replica_seed(self.h, replica, int(seed)))

    # -- checkpoint / restart
    def sweep_count(self):
        n = C.c_ulonglong(0)
        _check(self.lib.sn_get_sweep_count(self.h, C.byref(n)))
        return n.value

    def set_sweep_count(self, n):
        _check(self.lib.sn_set_sweep_count(self.h, int(n)))

    def set_counters(self, accept, reject, vacant, replica=0):
        _check(self.lib.sn_set_counters(self.h, replica, int(accept), int(reject), int(vacant)))

    # -- Z-slab plumbing
    def get_boundary(self, side, replica=0):
        g = self.params.cutoff
        a = np.empty((self.X, self.Y, g, 4), np.float32)
        _check(self.lib.sn_get_boundary(self.h, replica, side, a.ctypes.data))
        return a

    def set_ghost(self, side, planes, replica=0):
        a = np.ascontiguousarray(planes, np.float32)
        _check(self.lib.sn_set_ghost(self.h, replica, side, a.ctypes.data))

    def ipc_export(self):
        a, b = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
        _check(self.lib.sn_ipc_export(self.h, a, b))
        return bytes(a), bytes(b)

    def ipc_attach(self, side, lattice_handle: bytes, flags_handle: bytes):
        a = (C.c_ubyte * 64).from_buffer_copy(lattice_handle)
        b = (C.c_ubyte * 64).from_buffer_copy(flags_handle)
        _check(self.lib.sn_ipc_attach(self.h, side, a, b))

    def fp64_peak_tflops(self):
        out = C.c_double(0)
        _check(self.lib.sn_bench_fp64_peak(self.params.device, C.byref(out)))
        return out.value

    def fp32_peak_tflops(self):
        """FFMA microbenchmark on this handle's device (roofline denominator for bench.py)."""
        out = C.c_double(0)
        _check(self.lib.sn_bench_fp32_peak(self.params.device, C.byref(out)))
        return out.value

    def attach_peer(self, side, peer: "Simulation"):
        _check(self.lib.sn_attach_peer(self.h, side, peer.h))
