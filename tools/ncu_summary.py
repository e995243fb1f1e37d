"""Summarise an .ncu-rep: per profiled launch -- duration, DRAM bytes, instruction count, issue utilisation, stalls.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source N]   (N = launch index for the per-line histogram)"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
print("units:", {w: rows[1][i] for w, i in idx if rows[1][i]})
for n, r in enumerate(rows[2:]):
    print(f"-- launch {n}")
    for w, i in idx:
        print(f"   {w[:66]:68s}{r[i]}")
if "--source" in sys.argv:
    k = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hidx[2 * k] if 2 * k < len(hidx) else hidx[k]
    nxt = [h for h in hidx if h > start]
    end = nxt[0] - 1 if nxt else len(rows)
    h = rows[start]
    body = [r for r in rows[start + 1:end] if len(r) == len(h)]
    ci = {n: i for i, n in enumerate(h)}
    c = Counter()
    for r in body:
        c[int(r[ci["Instructions Executed"]] or 0)] += 1
    print("total", sum(a * b for a, b in c.items()))
    for n, cnt in sorted(c.items(), key=lambda x: -x[0] * x[1])[:10]:
        print(f"   executed {n:>8} times: {cnt:>4} SASS lines -> {n*cnt}")
    top = sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:25]
    for r in top:
        print(r[ci["# Samples"]].rjust(6), r[ci["Instructions Executed"]].rjust(8), r[ci["Source"]][:110])
