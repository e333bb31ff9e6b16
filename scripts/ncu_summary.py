"""Summarise an ncu report (ncu -i rep --page raw --csv piped to a file): per kernel time, DRAM bytes, tensor-pipe activity,
issue utilisation, top stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return float("nan")
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    grid = r[idx["Grid Size"]] if "Grid Size" in idx else "?"
    t = g(r, "gpu__time_duration.sum")
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    ten = g(r, "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active" in idx else float("nan")
    tens = [h for h in hdr if "pipe_tensor" in h and "pct" in h]
    tv = {h: r[idx[h]] for h in tens[:6]}
    iss = g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    st = sorted(((g(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)[:4]
    print(f"{name:28s} grid {grid:>8s} time {t:9.1f} us  dram rd {rd/1e6:8.1f} MB wr {wr/1e6:8.1f} MB -> {(rd+wr)/t/1e3 if t==t else 0:7.1f} GB/s | issue active {iss:5.1f} % | stalls {[(round(a,1), b) for a, b in st]}")
    print("     tensor:", tv)
