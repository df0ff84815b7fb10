#!/usr/bin/env python
"""bench.py -- attempted Metropolis site-updates/s of the StarryNight sweep on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON
line; for N > 1 it is launched under torchrun, one rank per GPU.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on):
512^3 lattice, DipoleCutOff = 3, T = 300 K, CageStrain = 1, one species, seeded
random start.  One "step" = SWEEPS_PER_STEP full-lattice sweeps (one attempt at
every site per sweep).  At N > 1 the same lattice is Z-slab decomposed (strong
scaling), halo planes pushed GPU-to-GPU over NVLink by the sweep kernel itself.

  value  whole-job attempts/s with the lattice resident in HBM (CUDA events on the
         library's own stream, max over ranks)
  e2e    the same metric through the C ABI with HOST buffers: every step uploads the
         lattice from pinned host memory (sn_set_lattice), sweeps, and reads the
         lattice and counters back (sn_get_lattice, sn_get_counters)
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="lattice edge (default: the 512^3 headline config)")
    ap.add_argument("--sweeps-per-step", type=int, default=SWEEPS_PER_STEP)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU sample for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process of the reference's replica-parallel mode: own lattice, own MT stream."""
    size, sample_sites_z, moves, seed, use_ref = args
    from oracle import oracle_api as oa
    X = Y = size
    Z = sample_sites_z
    beta = 1.0 / (float(np.float32(T_KELVIN)) / 300.0)
    p = oa.make_params(X, Y, Z, 3, 1.0, 0.0, (0.0, 0.0, 0.0), beta, 0, 3, T_KELVIN)
    if use_ref:
        r = oa.RefLib("f32")
        r.configure(p)
        r.seed(seed)
        r.initialise_lattice("random")               # lattice.c:25-37 through the reference's own code
        r.solid_solution([1.0, 0.0, 0.0], [1.0, 0.0, 0.0])
        r.mc_moves(min(moves, 20000))                # touch the code path once
        t0 = time.perf_counter()
        r.mc_moves(moves)                            # MC_moves(int), montecarlo-core.c:143
        dt = time.perf_counter() - t0
    else:
        o = oa.Oracle("f32")
        mt = o.mt(seed)
        lat = o.initialise_lattice(p, mt, "random")
        o.solid_solution(p, lat, mt, [1.0, 0.0, 0.0], [1.0, 0.0, 0.0])
        o.mc_moves(p, lat, mt, min(moves, 20000))
        t0 = time.perf_counter()
        o.mc_moves(p, lat, mt, moves)
        dt = time.perf_counter() - t0
    return moves, dt


def cpu_reference_rate(size, seconds, procs=None):
    """Attempts/s of the reference CPU implementation on a bounded sample of the workload:
    `procs` independent processes (the reference's only parallel mode), each running
    MC_moves on a size x size x Zs random lattice, cut-off 3, T = 300.  Zs is the full
    edge when memory allows, else the largest slab that fits (the chain's cost per attempt
    is set by the cache-missing 122-neighbour gather, which a slab of >= 64 planes of a
    512^2 cross-section -- 268 MB, far beyond any cache -- reproduces)."""
    import multiprocessing as mp
    from oracle import oracle_api as oa
    use_ref = oa.ref_available("f32")
    ncpu = os.cpu_count() or 1
    procs = procs or ncpu
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    # reference allocates (16 B site) + pointer tables; python-side copies are not made in the ref path
    per_plane = size * size * 16 * (1.0 if use_ref else 1.0)
    zs = size
    while zs > 64 and procs * zs * per_plane > 0.4 * avail:
        zs //= 2
    while procs > 1 and procs * zs * per_plane > 0.4 * avail:
        procs //= 2
    rate_guess = 1.5e5                                # attempts/s/core at this size
    moves = int(min(2 ** 31 - 1, max(2e5, rate_guess * seconds)))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(size, zs, moves, 0xDEADBEEF + T_KELVIN + i, use_ref) for i in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(m for m, _ in res)
    slowest = max(dt for _, dt in res)
    rate = total / slowest
    return dict(value=rate, unit=UNIT, cores=procs, kind="reference" if use_ref else "port",
                sample=f"{procs} independent processes x MC_moves({moves}) on a {size}x{size}x{zs} random lattice, "
                       f"cutoff 3, T=300 ({slowest:.1f} s of CPU work each; {wall:.1f} s wall incl. lattice init)",
                per_core=rate / procs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    info = None
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        info = cpu_reference_rate(args.size, per_step)
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.mean(vals))
    attempts_per_step = args.sweeps_per_step * args.size ** 3
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * attempts_per_step / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def workload_config(args, n):
    s = args.size
    return {"workload": f"{s}^3 lattice, DipoleCutOff=3, T={T_KELVIN} K, CageStrain=1, Efield=0, one species, random start "
                        f"(BASELINE.json configs[4])",
            "lattice": [s, s, s], "cutoff": 3, "neighbours": 122, "sweeps_per_step": args.sweeps_per_step,
            "attempts_per_step": args.sweeps_per_step * s ** 3, "decomposition": f"z-slabs x{n}" if n > 1 else "single GPU",
            "l2": "lattice (2.1 GB at 512^3) exceeds the 126 MB L2; no flush needed"}


def synthetic_slab(size, z0, nz, seed=1234):
    """Seeded unit dipoles for planes [z0, z0+nz) of the size^3 lattice, reproducible plane by plane
    (any rank can generate any plane), generated on the GPU and returned in pinned host memory."""
    import torch
    out = torch.empty((size, size, nz, 4), dtype=torch.float32, pin_memory=True)
    dev = torch.device("cuda", torch.cuda.current_device())

    def s64(c):                                       # 64-bit constant as a signed int64
        return c - (1 << 64) if c >= (1 << 63) else c

    ax = torch.arange(size, device=dev, dtype=torch.int64)
    chunk = 64                                        # planes per batch: a handful of launches for the whole slab
    for c0 in range(0, nz, chunk):
        zz = (torch.arange(c0, min(nz, c0 + chunk), device=dev, dtype=torch.int64) + z0) % size
        idx = (ax[:, None, None] * size + ax[None, :, None]) * size + zz[None, None, :]
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
        block = torch.stack([r * torch.cos(phi), r * torch.sin(phi), cz, torch.ones_like(cz)], -1)
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

    def __init__(self, index):
        self.index = index
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
            time.sleep(0.004)

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

    size = args.size
    if size % (32 * n):
        raise SystemExit(f"lattice edge {size} must be a multiple of {32 * n} for {n} slabs")
    nz = size // n
    z0 = rank * nz
    beta = sn.beta_of_T(T_KELVIN)
    sim = sn.Simulation(size, size, size, DipoleCutOff=3, CageStrain=1.0, K=0.0, Efield=(0.0, 0.0, 0.0), beta=beta,
                        seed=0xDEADBEEF + T_KELVIN, device=local, z0=z0 if n > 1 else 0, nz=nz if n > 1 else 0)
    host = synthetic_slab(size, z0, nz)
    host_out = torch.empty_like(host).pin_memory()

    ghosts = None
    if n > 1:                                        # ghost planes: the neighbours' boundary planes, regenerated locally
        ghosts = (synthetic_slab(size, (z0 - 3) % size, 3), synthetic_slab(size, (z0 + nz) % size, 3))

    def upload():
        sim.set_lattice_ptr(host.data_ptr())
        if n > 1:
            sim.set_ghost(0, ghosts[0].numpy())
            sim.set_ghost(1, ghosts[1].numpy())

    upload()
    if n > 1:                                        # wire the NVLink path: CUDA IPC handles around the ring
        handles = [None] * n
        dist.all_gather_object(handles, sim.ipc_export())
        lo_r, hi_r = (rank - 1) % n, (rank + 1) % n
        sim.ipc_attach(0, *handles[lo_r])
        sim.ipc_attach(1, *handles[hi_r])
        barrier()

    spp = args.sweeps_per_step
    attempts_step = spp * size ** 3                   # whole job
    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(args.warmup):
        sim.MC_sweeps_timed(spp)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_list, launches = [], 0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        barrier()
        ms, nl = sim.MC_sweeps_timed(spp)             # CUDA events on the library's stream
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

    # ---- end to end through the C ABI with host buffers -------------------------------------
    h2d = host.numel() * 4 + (sum(g.numel() * 4 for g in ghosts) if ghosts else 0)
    d2h = host_out.numel() * 4 + 24
    for _ in range(min(1, args.warmup)):
        upload(); barrier(); sim.MC_sweeps(spp); sim.get_lattice_ptr(host_out.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload()                                      # H2D from pinned memory (slab + its ghost planes)
        if n > 1:
            barrier()                                 # neighbours' uploads done before anyone pushes ghosts
        sim.MC_sweeps(spp)
        sim.get_lattice_ptr(host_out.data_ptr())      # D2H (synchronises)
        acc, rej, vac = sim.counters()
    barrier()
    e2e_s = time.perf_counter() - t0
    if n > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = attempts_step * args.steps / e2e_s
    # lattice-wide observables of the final state, merged over the slabs by one FP64 all_reduce (sanity of the run)
    from starrynight_b200 import slab as sn_slab
    merged = sn_slab.merge_observables(sim, dist if n > 1 else None, n, precision=sn.SN_PREC_F32)

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    fp32_peak = sim.fp32_peak_tflops() if rank == 0 else None
    sweep_launches = args.steps                      # one sn_tiled_kernel launch per sn_mc_sweeps call (all sweeps of a step);
    #                                                  slab runs add signal + wait kernels around it (in `launches`)
    avg_launch_ms = ms_local / max(1, sweep_launches)
    attempts_per_launch = (size * size * nz) * spp * args.steps / max(1, sweep_launches)
    achieved_tf = FLOP_PER_ATTEMPT * attempts_per_launch / (avg_launch_ms * 1e-3) / 1e12
    achieved_gbs = BYTES_PER_ATTEMPT * attempts_per_launch / (avg_launch_ms * 1e-3) / 1e9

    traffic = None
    try:                                              # DRAM bytes per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["bytes_per_launch"] / tj["attempts_per_launch"] * attempts_per_launch
    except Exception:
        pass
    if rank == 0:
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        theo = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * n, "d2h_bytes_per_step": d2h * n},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "kernel": "sn_tiled_kernel", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak if fp32_peak else None,
                         "peak_source": "FFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 entry)",
                         "peak_theoretical": theo, "flop_per_attempt": FLOP_PER_ATTEMPT,
                         "attempts_per_launch": attempts_per_launch, "avg_launch_ms": avg_launch_ms,
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                                 "peak_source": hbm_src, "bytes_per_attempt": BYTES_PER_ATTEMPT},
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per launch, scaled by attempts per launch)"},
            "accept_ratio": merged["accept"] / max(1, merged["accept"] + merged["reject"]),
            "energy_per_site": float(merged["energy"].sum() / merged["nsites"]),
            "wall_s_timed_region": t_wall,
        }
        if n == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_rate(size, args.cpu_seconds)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:                     # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    sim.close()
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
