"""Summarise an .ncu-rep (first kernel): key metrics, stall-reason shares, and sample attribution
to coarse SASS regions.  usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [region_size]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; region = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
m = dict(zip(hdr, r))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
print("kernel:", m.get("Kernel Name", "?")[:60])
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__icc_request_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    print(f"  {k} = {m.get(k)} [{units[hdr.index(k)] if k in hdr else ''}]")
tot = g("smsp__pcsamp_sample_count")
print("stall shares (% of samples):")
for k in sorted(hdr):
    if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and g(k) > 0.01 * tot:
        print(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):22s} {100 * g(k) / tot:5.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(x[ia]) for x in data)
print(f"SASS regions of {region} instr (samples %, executed warp-instr):")
for k in range(0, len(data), region):
    seg = data[k:k + region]; s = sum(int(x[ia]) for x in seg)
    if s: print(f"  {k:5d} {100 * s / tot:5.1f}%  exec {sum(int(x[ie]) for x in seg):9d}  {seg[0][1].strip()[:44]}")
top = sorted(range(len(data)), key=lambda i: -int(data[i][ia]))[:12]
print("top instructions:")
for i in sorted(top):
    print(f"  {i:5d} {100 * int(data[i][ia]) / tot:5.1f}%  {data[i][1].strip()[:70]}")
