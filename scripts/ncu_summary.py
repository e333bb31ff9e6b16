"""Summarise an ncu report: per launch time, grid, registers, DRAM bytes, tensor-pipe activity, issue utilisation, L2 hit rate,
shared-memory bank conflicts and the top warp-stall reasons.  Usage: python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    if k not in idx: return float("nan")
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return float("nan")
def gu(r, k):
    return f"{g(r, k):.6g}{units[idx[k]]}" if k in idx else "n/a"
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
def find(sub):
    c = [h for h in hdr if sub in h]
    return c[0] if c else None
TEN = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"                       # B200_PROFILING.md: sm__pipe_tensor_cycles_active
DRAMP = find("dram__throughput.avg.pct_of_peak_sustained_elapsed") or "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("slotvps::", "").replace("void ", "")
    st = sorted(((g(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)[:3]
    print(f"{name:28s} time={gu(r, 'gpu__time_duration.sum')}  grid={r[idx['Grid Size']]}  regs={gu(r, 'launch__registers_per_thread')}  "
          f"dram_rd={gu(r, 'dram__bytes_read.sum')}  dram_wr={gu(r, 'dram__bytes_write.sum')}  dram%={g(r, DRAMP):.1f}  "
          f"tensor_pipe%_active={g(r, TEN):.1f}  issue_active%={g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}  "
          f"l2_hit%={g(r, 'lts__t_sector_hit_rate.pct'):.1f}  smem_bank_conflicts={g(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.0f}  "
          f"stalls={[(round(a, 1), b) for a, b in st]}")
