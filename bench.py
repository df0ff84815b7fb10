#!/usr/bin/env python
"""bench.py -- attempted Metropolis site-updates/s of the StarryNight sweep on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON
line; for N > 1 it is launched under torchrun, one rank per GPU.

Default workload (BASELINE.json configs[4], the configuration the metric is quoted on):
512^3 lattice, DipoleCutOff = 3, T = 300 K, CageStrain = 1, one species, seeded
random start.  One "step" = SWEEPS_PER_STEP full-lattice sweeps (one attempt at
every site per sweep).  At N > 1 the same lattice is Z-slab decomposed (strong
scaling), halo planes pushed GPU-to-GPU over NVLink by the sweep kernel itself.
``--workload c2|c3|c4|c5b`` measures the other BASELINE configurations through the
same contract (see WORKLOADS); ``--workload obs`` measures one analysis pass (the observables of
analysis_midpoint) in sites/s; the default line is unchanged.

  value  whole-job attempts/s with the lattice resident in HBM (CUDA events on the
         library's own stream, max over ranks)
  e2e    the same metric through the C ABI with HOST buffers: every step uploads the
         step's lattice from pinned host memory (sn_set_lattice_async), fills the slab
         ghost planes device to device (sn_pull_ghosts), sweeps, and reads the lattice
         and counters back (sn_get_lattice_async, sn_get_counters).  Successive steps
         are independent batches, so two handles are used in
         rotation: step i+1's upload and step i-1's download travel over PCIe while step i is swept
         (copies at PCIe rate cannot be hidden inside one 20-sweep step: 2 x 2.1 GB at
         ~55 GB/s is 78 ms beside 115-145 ms of sweeps).  ``e2e.serial`` is the same
         loop with one handle and synchronous calls.
  state_hash  64-bit position-keyed hash of the final lattice bits of the
         device-resident run: identical at every N iff the N-GPU chain is the 1-GPU chain
  roofline  FP32 CUDA-core roofline of the sweep kernel: algorithmic 2500 flop per
         attempt x attempts per launch / measured launch duration,
         against an FMA-peak microbenchmark run here (MEASURED_PEAKS.json has no FP32
         entry); HBM figures beside it
  cpu_baseline  the reference's own CPU code (oracle/_ref, else the oracle port) timed
         on this box's cores on a bounded sample of the same workload

``--impl reference`` times only that CPU implementation (all host cores, the
reference's one parallel mode: independent processes, Makefile:49-63).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "metropolis_site_updates_per_s"
UNIT = "attempts/s"
FLOP_PER_ATTEMPT = 2500.0       # 20*N_nb + 6*N_nn + 24, N_nb = 122, N_nn = 6
BYTES_PER_ATTEMPT = 32.0        # 16 B read + <= 16 B write per site per sweep
SWEEPS_PER_STEP = 20
T_KELVIN = 300


WORKLOADS = {
    # name: lattice, replicas per job, couplings; `slabs`: Z-slab decomposition at N > 1 (strong scaling), else the
    # replicas are dealt out to the ranks (weak scaling: every GPU gets `replicas` of its own)
    "c5": dict(what="512^3 lattice, DipoleCutOff=3, T=300 K, CageStrain=1, Efield=0, one species, random start (BASELINE.json configs[4])",
               shape=(512, 512, 512), replicas=1, slabs=True, sweeps=20),
    "c5b": dict(what="1024^3 lattice, DipoleCutOff=3, T=300 K, CageStrain=1, one species, random start (BASELINE.json configs[4], larger size)",
                shape=(1024, 1024, 1024), replicas=1, slabs=True, sweeps=5),
    "c2": dict(what="2-D 100x100x1 lattice, T=300 K, Efield=(0.02,0,0), DipoleCutOff=3 (28 neighbours), independent replicas/seeds "
                    "(BASELINE.json configs[1]; starrynight.cfg:20)",
               shape=(100, 100, 1), replicas=1184, slabs=False, sweeps=200, efield=(0.02, 0.0, 0.0)),
    "c3": dict(what="64^3 lattice, DipoleCutOff=3, T = 0..500 K step 25 x CageStrain {0,1,2} = 63 replicas, one launch "
                    "(BASELINE.json configs[2]; the reference's `superparallel` grid, Makefile:52-54)",
               shape=(64, 64, 64), replicas=63, slabs=False, sweeps=40, temps=list(range(0, 501, 25)), cages=(0.0, 1.0, 2.0)),
    "obs": dict(what="one analysis pass (analysis_midpoint, main.c:51-96: recombination_calculator + radial_order_parameter + lattice_Efield + "
                     "potential map, plus polarisation / Landau order) over a resident 256^3 lattice, DipoleCutOff=3, after 2 sweeps at T=300 K",
                shape=(256, 256, 256), replicas=1, slabs=True, sweeps=2, analysis=True),
    "c4": dict(what="128^3 MA/FA solid solution: Dipoles=[1.0,0.5,0.0] Prevalence=[0.6,0.3,0.1] (two species + vacancies), triangular Efield.x "
                    "ramp +-0.1 in 64 points, 1 sweep per point, 8 independent loops (seeds) (BASELINE.json configs[3])",
               shape=(128, 128, 128), replicas=8, slabs=True, sweeps=64, species=((1.0, 0.5, 0.0), (0.6, 0.3, 0.1)), ramp=(0.1, 64)),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=None, help="lattice edge override for the cubic workloads")
    ap.add_argument("--shape", default=None, help="X,Y,Z override (experiments: e.g. the per-GPU slab shape of a larger run)")
    ap.add_argument("--sweeps-per-step", type=int, default=None)
    ap.add_argument("--replicas", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU sample for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: device-resident part only")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    if a.size:
        w["shape"] = (a.size, a.size, a.size if w["shape"][2] > 1 else 1)
    if a.shape:
        w["shape"] = tuple(int(v) for v in a.shape.split(","))
    if a.sweeps_per_step:
        w["sweeps"] = a.sweeps_per_step
    if a.replicas:
        w["replicas"] = a.replicas
    a.w = w
    return a


def flop_per_attempt(shape):
    """20 N_nb + 6 N_nn + 24: 2500 for 3-D cut-off 3 (122 neighbours), 608 for Z == 1 (28)."""
    nb, nn = (28, 4) if shape[2] == 1 else (122, 6)
    return 20.0 * nb + 6.0 * nn + 24.0


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process of the reference's replica-parallel mode: own lattice, own MT stream."""
    shape, moves, seed, use_ref, efield, species = args
    from oracle import oracle_api as oa
    X, Y, Z = shape
    beta = 1.0 / (float(np.float32(T_KELVIN)) / 300.0)
    p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, efield, beta, 0, 3, T_KELVIN)
    lengths, prev = species if species else ([1.0, 0.0, 0.0], [1.0, 0.0, 0.0])
    if use_ref:
        r = oa.RefLib("f32")
        r.configure(p)
        r.seed(seed)
        r.initialise_lattice("random")               # lattice.c:25-37 through the reference's own code
        r.solid_solution(list(lengths), list(prev))
        r.mc_moves(min(moves, 20000))                # touch the code path once
        t0 = time.perf_counter()
        r.mc_moves(moves)                            # MC_moves(int), montecarlo-core.c:143
        dt = time.perf_counter() - t0
    else:
        o = oa.Oracle("f32")
        mt = o.mt(seed)
        lat = o.initialise_lattice(p, mt, "random")
        o.solid_solution(p, lat, mt, list(lengths), list(prev))
        o.mc_moves(p, lat, mt, min(moves, 20000))
        t0 = time.perf_counter()
        o.mc_moves(p, lat, mt, moves)
        dt = time.perf_counter() - t0
    return moves, dt


def cpu_reference_rate(w, seconds, procs=None):
    """Attempts/s of the reference CPU implementation on a bounded sample of the workload:
    `procs` independent processes (the reference's only parallel mode, Makefile:49-63), each running
    MC_moves on a lattice of the workload's cross-section, cut-off 3, T = 300.  The Z extent is the full
    one when memory allows, else the largest slab that fits (the chain's cost per attempt is set by the
    cache-missing 122-neighbour gather, which a slab of >= 64 planes of a 512^2 cross-section -- 268 MB,
    far beyond any cache -- reproduces)."""
    import multiprocessing as mp
    from oracle import oracle_api as oa
    use_ref = oa.ref_available("f32")
    ncpu = os.cpu_count() or 1
    procs = procs or ncpu
    X, Y, Z = w["shape"]
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    per_plane = X * Y * 16.0
    zs = Z
    while zs > 64 and procs * zs * per_plane > 0.4 * avail:
        zs //= 2
    while procs > 1 and procs * zs * per_plane > 0.4 * avail:
        procs //= 2
    rate_guess = 1.5e5 if X * Y * zs > 4e6 else 3.5e5 if Z > 1 else 1.5e6      # attempts/s/core
    moves = int(min(2 ** 31 - 1, max(2e5, rate_guess * seconds)))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [((X, Y, zs), moves, 0xDEADBEEF + T_KELVIN + i, use_ref, tuple(w.get("efield", (0.0, 0.0, 0.0))),
                                      w.get("species")) for i in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(m for m, _ in res)
    slowest = max(dt for _, dt in res)
    rate = total / slowest
    return dict(value=rate, unit=UNIT, cores=procs, kind="reference" if use_ref else "port",
                sample=f"{procs} independent processes x MC_moves({moves}) on a {X}x{Y}x{zs} random lattice, "
                       f"cutoff 3, T=300 ({slowest:.1f} s of CPU work each; {wall:.1f} s wall incl. lattice init)",
                per_core=rate / procs, attempts=total, seconds=slowest)


def run_reference(args):
    """The reference's own CPU code on this box's cores.  A step here is the bounded sample itself: `attempts` moves
    spread over all cores, timed by the slowest process -- ms_per_step is that measured time, not an extrapolation to
    the GPU arm's step size (config.attempts_per_step); value = attempts / time is the rate the ratio is taken on."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals, secs, atts = [], [], []
    info = None
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        info = cpu_reference_rate(args.w, per_step)
        if i >= args.warmup:
            vals.append(info["value"]); secs.append(info["seconds"]); atts.append(info["attempts"])
    v = float(np.sum(atts) / np.sum(secs))
    cfg = workload_config(args, 1)
    cfg["reference_step"] = {"attempts": float(np.mean(atts)), "note": "one step of this arm = one bounded sample (all cores), see cpu_baseline.sample"}
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "strong" if args.w["slabs"] else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def workload_config(args, n):
    w = args.w
    X, Y, Z = w["shape"]
    reps = w["replicas"] * (1 if w["slabs"] else n)
    nb = 28 if Z == 1 else 122
    return {"workload": w["what"], "name": args.workload, "lattice": [X, Y, Z], "replicas": reps, "cutoff": 3, "neighbours": nb,
            "sweeps_per_step": w["sweeps"], "attempts_per_step": w["sweeps"] * X * Y * Z * reps,
            "decomposition": (f"z-slabs x{n}" if n > 1 else "single GPU") if w["slabs"] else f"{w['replicas']} replicas per GPU x{n}",
            "l2": ("lattice (%.1f GB) exceeds the 126 MB L2; no flush needed" % (X * Y * Z * reps * 16 / 1e9)) if X * Y * Z * reps * 16 > 2e8
                  else "working set fits in L2/shared memory by design (small lattices stay on chip); a 256 MB buffer is written between steps to flush L2"}


def synthetic_slab(shape, z0, nz, seed=1234, species=None):
    """Seeded unit dipoles for planes [z0, z0+nz) of the lattice, reproducible plane by plane
    (any rank can generate any plane), generated on the GPU and returned in pinned host memory."""
    import torch
    X, Y, Z = shape
    out = torch.empty((X, Y, nz, 4), dtype=torch.float32, pin_memory=True)
    dev = torch.device("cuda", torch.cuda.current_device())

    def s64(c):                                       # 64-bit constant as a signed int64
        return c - (1 << 64) if c >= (1 << 63) else c

    ax = torch.arange(X, device=dev, dtype=torch.int64)
    ay = torch.arange(Y, device=dev, dtype=torch.int64)
    chunk = 64                                        # planes per batch: a handful of launches for the whole slab
    for c0 in range(0, nz, chunk):
        zz = (torch.arange(c0, min(nz, c0 + chunk), device=dev, dtype=torch.int64) + z0) % Z
        idx = (ax[:, None, None] * Y + ay[None, :, None]) * Z + zz[None, None, :]
        # splitmix64 of (global site index, seed): the value of a site does not depend on who generates it
        h = idx * s64(0x9E3779B97F4A7C15) + seed
        h = (h ^ ((h >> 30) & ((1 << 34) - 1))) * s64(0xBF58476D1CE4E5B9)
        h = (h ^ ((h >> 27) & ((1 << 37) - 1))) * s64(0x94D049BB133111EB)
        h = h ^ ((h >> 31) & ((1 << 33) - 1))
        u1 = ((h >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)
        u2 = ((h >> 8) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)
        cz = 1.0 - 2.0 * u1
        phi = 6.283185307179586 * u2
        r = torch.sqrt(torch.clamp(1.0 - cz * cz, min=0.0))
        if species:                                   # lengths by prevalence (solid_solution, lattice.c:139-165), from the low hash bits
            lengths, prev = species
            u3 = (h & 0xFF).to(torch.float32) * (1.0 / 256.0)
            edges = torch.tensor(np.cumsum(prev) / np.sum(prev), device=dev, dtype=torch.float32)
            ln = torch.tensor(lengths, device=dev, dtype=torch.float32)[torch.bucketize(u3, edges[:-1], right=True)]
        else:
            ln = torch.ones_like(cz)
        block = torch.stack([r * torch.cos(phi), r * torch.sin(phi), cz, ln], -1)
        out[:, :, c0:c0 + block.shape[2], :].copy_(block)
        del idx, h, u1, u2, cz, phi, r, block
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few ms from a
    thread (the timed region of an 8-GPU run is tens of ms -- too short for an nvidia-smi loop), with
    `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period=0.004):
        self.index = index
        self.period = period
        self.rows = []          # (sm_mhz, reasons bitmask or set)
        self.max_mhz = None
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._visible_index()}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, {k for k, b in bits.items() if mask & b}))
            except Exception:
                pass
            time.sleep(self.period)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                mhz, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            self.max_mhz = max(self.max_mhz or 0.0, mx)
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((mhz, {n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")}))

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
        sm = [r[0] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons |= r[1]
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import starrynight_b200 as sn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    n = world
    torch.cuda.set_device(local)
    if n > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = args.w
    X, Y, Z = w["shape"]
    slabs = w["slabs"] and n > 1
    if slabs and Z % (32 * n):
        raise SystemExit(f"Z = {Z} must be a multiple of {32 * n} for {n} slabs")
    nz = Z // n if slabs else Z
    z0 = rank * nz if slabs else 0
    reps = w["replicas"]
    beta = sn.beta_of_T(T_KELVIN)
    species = w.get("species")
    seed0 = 0xDEADBEEF + T_KELVIN + (0 if w["slabs"] else 7919 * rank)      # replica workloads: every rank its own streams

    def make_sim():
        sim = sn.Simulation(X, Y, Z, DipoleCutOff=3, CageStrain=1.0, K=0.0, Efield=tuple(w.get("efield", (0.0, 0.0, 0.0))), beta=beta,
                            nreplicas=reps, seed=seed0, device=local, z0=z0 if slabs else 0, nz=nz if slabs else 0)
        if "temps" in w:                              # T x CageStrain grid, one replica per point
            grid = [(t, c) for c in w["cages"] for t in w["temps"]]
            for r, (t, c) in enumerate(grid[:reps]):
                sim.set_T(t, r)
                sim.set_cagestrain(c, r)
        return sim

    def wire(sim):
        if slabs:                                     # the NVLink path: CUDA IPC handles around the ring
            handles = [None] * n
            dist.all_gather_object(handles, sim.ipc_export())
            sim.ipc_attach(0, *handles[(rank - 1) % n])
            sim.ipc_attach(1, *handles[(rank + 1) % n])

    hosts = [synthetic_slab(w["shape"], z0, nz, seed=1234 + 17 * r + (0 if w["slabs"] else 100003 * rank), species=species) for r in range(min(reps, 8))]
    host_of = lambda r: hosts[r % len(hosts)]                         # replica workloads re-use 8 distinct start lattices
    sim = make_sim()
    wire(sim)
    for r in range(reps):
        sim.set_lattice_ptr(host_of(r).data_ptr(), r)
    sim.pull_ghosts()
    barrier()

    spp = w["sweeps"]
    ramp = w.get("ramp")

    def step(s, timed):
        """One step's sweeps on handle s: (device ms, launches) when timed."""
        if not ramp:
            return s.MC_sweeps_timed(spp) if timed else (s.MC_sweeps(spp), 0)
        amp, npts = ramp                              # hysteresis: triangular Efield.x ramp, spp / npts sweeps per point
        ms = nl = 0.0
        for k in range(npts):
            ph = 4.0 * k / npts
            e = amp * (ph if ph < 1 else 2 - ph if ph < 3 else ph - 4)
            for r in range(reps):
                s.set_efield((e, 0.0, 0.0), r)
            if timed:
                a, b = s.MC_sweeps_timed(max(1, spp // npts)); ms += a; nl += b
            else:
                s.MC_sweeps(max(1, spp // npts))
        return ms, int(nl)

    sites_job = X * Y * Z * reps * (1 if w["slabs"] else n)            # whole job
    attempts_step = spp * sites_job
    small = X * Y * Z * reps * 16 < 2e8
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda") if small else None
    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(args.warmup):
        step(sim, True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_list, launches = [], 0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1.0)
        barrier()
        ms, nl = step(sim, True)                      # CUDA events on the library's stream
        ms_list.append(ms)
        launches += nl
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms_local = float(np.sum(ms_list))
    if n > 1:
        t = torch.tensor([ms_local], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    else:
        ms_total = ms_local
    value = attempts_step * args.steps / (ms_total * 1e-3)

    # lattice-wide observables and content hash of the final state of the device-resident run
    from starrynight_b200 import slab as sn_slab
    merged = sn_slab.merge_observables(sim, dist if n > 1 else None, n, precision=sn.SN_PREC_F32)
    hsh = sum(sim.state_hash(r) for r in range(reps)) % (1 << 64)
    if n > 1:
        parts = [None] * n
        dist.all_gather_object(parts, hsh)
        hsh = sum(parts) % (1 << 64)

    # ---- end to end through the C ABI with host buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        # handles used in rotation.  Three were measured too (profiles/r02_bench_c5_n8_three_handles.json): no gain -- where the
        # copies outlast the sweeps (8 GPUs) the host side is saturated and the loop runs at the serial sum either way
        NH = 2
        # replica batches travel as one dense block per direction (sn_set_lattices_async / sn_get_lattices_async)
        batch = reps > 1
        if batch:
            host_all = torch.stack([host_of(r) for r in range(reps)]).pin_memory()
            outs = [torch.empty_like(host_all).pin_memory() for _ in range(NH)]
        else:
            outs = [[torch.empty_like(host_of(r)).pin_memory() for r in range(reps)] for _ in range(NH)]
        h2d = sum(host_of(r).numel() * 4 for r in range(reps))
        d2h = h2d + 24 * reps
        sims = [sim] + [make_sim() for _ in range(NH - 1)]
        for s_ in sims[1:]:
            wire(s_)

        def e2e_loop(nsteps, pipelined):
            barrier()
            t0 = time.perf_counter()
            for i in range(nsteps):
                s, o = (sims[i % NH], sims[(i - 1) % NH]) if pipelined else (sims[0], None)
                s.synchronize()                       # its previous download has landed: host buffers are free again
                if i >= (NH if pipelined else 1):
                    s.counters()                      # the step's result: ACCEPT / REJECT (and the lattice in outs)
                if batch:
                    s.set_lattices_async(host_all.data_ptr())         # H2D from pinned memory, every replica in one block
                else:
                    s.set_lattice_async(host_of(0).data_ptr(), 0)     # H2D from pinned memory
                if o is not None:
                    s.order_after(o)                  # the handles' sweep kernels keep one order on every GPU
                s.pull_ghosts()                       # slab ghost planes, device to device, handshake included
                step(s, False)
                if batch:
                    s.get_lattices_async(outs[i % NH].data_ptr())     # D2H
                else:
                    s.get_lattice_async(outs[i % NH][0].data_ptr(), 0)
            for s in (sims if pipelined else sims[:1]):
                s.synchronize()
                s.counters()
            barrier()
            dt = time.perf_counter() - t0
            if n > 1:
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        e2e_loop(NH, True)                            # warm-up (staging buffers, pinned mappings)
        dt_pipe = e2e_loop(args.steps, True)
        e2e_loop(1, False)
        dt_serial = e2e_loop(args.steps, False)
        e2e = {"value": attempts_step * args.steps / dt_pipe, "unit": UNIT, "h2d_bytes_per_step": h2d * n, "d2h_bytes_per_step": d2h * n,
               "mode": "%d handles used in rotation: step i+1's upload and step i-1's download overlap step i's sweeps; every step's "
                       "lattice is uploaded from and downloaded to pinned host memory inside the timed region" % NH,
               "serial": {"value": attempts_step * args.steps / dt_serial, "mode": "one handle, each step upload -> sweeps -> download back to back"},
               "seconds": dt_pipe}
        for s_ in sims[1:]:
            s_.close()

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    fp32_peak = sim.fp32_peak_tflops() if rank == 0 else None
    kern_id = sim.kernel_in_use()
    kern_name = {sn.SN_KERNEL_TILED: "sn_tiled_kernel", sn.SN_KERNEL_RESIDENT: "sn_resident_kernel", sn.SN_KERNEL_COLOUR: "sn_colour_pass_kernel"}.get(kern_id, "?")
    sweep_launches = max(1, args.steps * (ramp[1] if ramp else 1))     # one sweep-kernel launch per sn_mc_sweeps call (tiled / resident)
    avg_launch_ms = ms_local / sweep_launches
    attempts_per_launch = (X * Y * nz * reps) * spp * args.steps / sweep_launches
    W = flop_per_attempt(w["shape"])
    achieved_tf = W * attempts_per_launch / (avg_launch_ms * 1e-3) / 1e12
    achieved_gbs = BYTES_PER_ATTEMPT * attempts_per_launch / (avg_launch_ms * 1e-3) / 1e9

    traffic = None
    try:                                              # DRAM bytes per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kern_name == "sn_tiled_kernel":
            traffic = tj["bytes_per_launch"] / tj["attempts_per_launch"] * attempts_per_launch
    except Exception:
        pass
    if rank == 0:
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        theo = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if w["slabs"] else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "kernel": kern_name, "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak if fp32_peak else None,
                         "basis": "ALGORITHMIC flops: the local-field form of site_energy costs 20 flop per neighbour (SURVEY.md 8d), "
                                  "%d per attempt; the kernel executes fewer (pair symmetry and vanishing tensor entries: see "
                                  "roofline.executed)" % int(W),
                         "peak_source": "FFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 entry)",
                         "peak_theoretical": theo, "flop_per_attempt": W,
                         "attempts_per_launch": attempts_per_launch, "avg_launch_ms": avg_launch_ms,
                         "executed": executed_roofline(kern_name, attempts_per_launch, avg_launch_ms, theo),
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                                 "peak_source": hbm_src, "bytes_per_attempt": BYTES_PER_ATTEMPT},
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per launch, scaled by attempts per launch)"},
            "accept_ratio": merged["accept"] / max(1, merged["accept"] + merged["reject"]),
            "energy_per_site": float(merged["energy"].sum() / merged["nsites"]),
            "state_hash": "%016x" % hsh,
            "wall_s_timed_region": t_wall,
        }
        if n == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_rate(w, args.cpu_seconds)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:                     # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    sim.close()
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- analysis workload
RDF_OFFSETS, POT_OFFSETS, EF_OFFSETS = 2969, 924, 256      # lattice vectors with r^2 <= 80 / 0 < d <= 6 / 0 < d <= 4 (analysis.c:540, 68, 397)
RDF_FLOP_PER_PAIR = 22.0     # FE dot 5 + two n.p dots 10 + AFE combine 3 + two histogram adds 2 + the products with n 2 (analysis.c:566-578)


def _analysis_cpu(shape, seed):
    """The reference's own analysis routines (oracle/_ref) on a small lattice: seconds per site of each."""
    from oracle import oracle_api as oa
    import tempfile
    X, Y, Z = shape
    p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, (0.0, 0.0, 0.0), 1.0, 0, 3, T_KELVIN)
    use_ref = oa.ref_available("f32")
    lat = oa.random_lattice(X, Y, Z, seed=seed)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        if use_ref:
            r = oa.RefLib("f32"); r.configure(p); r.set_lattice(lat)
            calls = {"rdf": lambda: r.rdf_file(os.path.join(d, "rdf.dat")), "potential": r.potential_map,
                     "efield": lambda: r.efield_map(4, False), "recombination": lambda: r.recombination_log(os.path.join(d, "rec.log"))}
        else:
            o = oa.Oracle("f32")
            calls = {"rdf": lambda: o.rdf(p, lat), "potential": lambda: o.potential_map(p, lat),
                     "efield": lambda: o.efield_map(p, lat, 4, False), "recombination": lambda: o.recombination(p, lat)}
        for k, f in calls.items():
            t0 = time.perf_counter(); f(); out[k] = time.perf_counter() - t0
    return out, ("reference" if use_ref else "port")


def run_analysis(args):
    """--workload obs: sites per second through one analysis pass.  value: lattice resident in HBM, maps left to the
    caller in pinned host memory (the pass the driver runs every mega-step); e2e: the same pass including the upload of
    the lattice from pinned host memory.  Z-slabs at N > 1: every rank analyses its own slab (neighbour planes are read
    over NVLink by the kernels), the histogram / partition sums are merged with one small all_reduce."""
    import torch
    import torch.distributed as dist
    import starrynight_b200 as sn
    from starrynight_b200 import slab as sn_slab

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n = world
    w = args.w
    X, Y, Z = w["shape"]
    if args.impl == "reference":
        if rank != 0:
            return
        t_all = time.perf_counter()
        edge = 40
        secs, kind = _analysis_cpu((edge, edge, edge), 5)
        for _ in range(max(0, args.warmup + args.steps - 1)):
            secs, kind = _analysis_cpu((edge, edge, edge), 5)
        tot = sum(secs.values())
        v = edge ** 3 / tot
        cfg = workload_config(args, 1); cfg["reference_step"] = {"sites": edge ** 3, "seconds": secs}
        print(json.dumps({"metric": "analysis_sites_per_s", "value": v, "unit": "sites/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * tot, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "sites/s", "cores": 1, "kind": kind,
                                           "sample": f"the reference's serial analysis routines on a {edge}^3 random lattice (cost per site does not depend on the size)"},
                          "e2e": {"value": v, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
                          "wall_s": time.perf_counter() - t_all}), flush=True)
        return
    torch.cuda.set_device(local)
    if n > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nz = Z // n; z0 = rank * nz
    host = synthetic_slab(w["shape"], z0, nz, seed=1234)
    sim = sn.Simulation(X, Y, Z, DipoleCutOff=3, CageStrain=1.0, beta=sn.beta_of_T(T_KELVIN), seed=0xDEADBEEF + T_KELVIN, device=local,
                        z0=z0 if n > 1 else 0, nz=nz if n > 1 else 0)
    if n > 1:
        sn_slab.wire_ipc(sim, dist, n, rank)
    sim.set_lattice_ptr(host.data_ptr(), 0)
    sim.pull_ghosts()
    sim.MC_sweeps(w["sweeps"])
    sim.synchronize()
    nsl = X * Y * nz
    vmap = torch.empty(nsl, dtype=torch.float64).pin_memory().numpy()
    emap = torch.empty(nsl, dtype=torch.float64).pin_memory().numpy()
    parts = {}

    def timed(name, f, *a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); parts[name] = parts.get(name, 0.0) + time.perf_counter() - t0
        return r

    def one_pass():
        res = {}
        res["P"] = timed("polarisation", sim.polarisation)
        res["landau"] = timed("landau_order", sim.landau_order)
        res["rec"] = timed("recombination", sim.recombination_partial)
        res["rdf"] = timed("rdf", sim.radial_order_parameter)
        timed("efield_map", sim.dipole_electricfield, 4, False, 0, emap)
        timed("potential_map", sim.dipole_potential, 0, vmap)
        if n > 1:                                     # the one collective of the pass: histogram and partition sums
            t = torch.from_numpy(np.concatenate([res["rdf"][0], res["rdf"][1], np.asarray(res["rec"], np.float64)[:5]])).cuda()
            dist.all_reduce(t)
            res["merged"] = t.cpu().numpy()
        return res

    for _ in range(args.warmup):
        one_pass()
    parts.clear()
    barrier()
    sampler = ClockSampler(local, period=0.25)         # many short synchronous calls here: frequent NVML queries contend with the launches and slow rank 0 down
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = one_pass()
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the lattice comes from pinned host memory every step
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        sim.set_lattice_ptr(host.data_ptr(), 0)
        sim.pull_ghosts()
        one_pass_res = one_pass()
    barrier()
    dt_e2e = time.perf_counter() - t1
    if n > 1:
        t = torch.tensor([dt, dt_e2e], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt, dt_e2e = [float(v) for v in t.tolist()]
    if rank == 0:
        sites = X * Y * Z
        rdf_ms = 1e3 * parts["rdf"] / (2 * args.steps)              # both loops ran it
        fp64_peak = sim.fp64_peak_tflops()
        ach = RDF_FLOP_PER_PAIR * RDF_OFFSETS * nsl / (rdf_ms * 1e-3) / 1e12
        fe, afe, cnt = res["rdf"]
        line = {"metric": "analysis_sites_per_s", "value": sites * args.steps / dt, "unit": "sites/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, n),
                "e2e": {"value": sites * args.steps / dt_e2e, "unit": "sites/s", "h2d_bytes_per_step": sites * 16, "d2h_bytes_per_step": 2 * sites * 8 + 8 * (2 * 81 + 20),
                        "mode": "lattice uploaded from pinned host memory every step; potential and |E| maps downloaded to pinned host memory"},
                "gpu_launches": 9 * args.steps, "clocks": clocks,
                "parts_ms": {k: 1e3 * v / (2 * args.steps) for k, v in parts.items()},
                "roofline": {"bound": "fp64", "kernel": "sn_rdf_tiled_kernel", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                             "basis": "ALGORITHMIC flops of radial_order_parameter: %d pair terms per site x %d flop (analysis.c:566-578); the kernel walks one of every +-d pair" % (RDF_OFFSETS, int(RDF_FLOP_PER_PAIR)),
                             "peak_source": "DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)", "avg_launch_ms": rdf_ms, "traffic": None},
                "rdf_nearest_shell": {"fe": float(fe[1] / cnt[1]), "afe": float(afe[1] / cnt[1])}}
        if n == 1 and not args.no_cpu_baseline:
            try:
                secs, kind = _analysis_cpu((40, 40, 40), 5)
                line["cpu_baseline"] = {"value": 40 ** 3 / sum(secs.values()), "unit": "sites/s", "cores": 1, "kind": kind,
                                        "sample": "the reference's serial analysis routines (rdf %.2f s, potential %.2f s, efield %.2f s, recombination %.2f s) on a 40^3 random lattice"
                                                  % (secs["rdf"], secs["potential"], secs["efield"], secs["recombination"])}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "sites/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    sim.close()
    if n > 1:
        dist.barrier(); dist.destroy_process_group()


def executed_roofline(kernel, attempts_per_launch, avg_launch_ms, peak_tf):
    """Pipe-level view beside the algorithmic roofline: FP32 instructions the kernel really executes per attempt
    (ncu inst_executed_pipe_fma of the committed capture, profiles/traffic.json), each counted as one FMA = 2 flop."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kernel != "sn_tiled_kernel" or "fp32_instr_per_attempt" not in tj:
            return None
        f = tj["fp32_instr_per_attempt"]
        tf = 2.0 * f * attempts_per_launch / (avg_launch_ms * 1e-3) / 1e12
        return {"fp32_instr_per_attempt": f, "achieved": tf, "unit": "TFLOP/s", "frac_of_theoretical": tf / peak_tf,
                "source": tj.get("source", "profiles/traffic.json")}
    except Exception:
        return None


def main():
    args = parse_args()
    if args.w.get("analysis"):
        run_analysis(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
