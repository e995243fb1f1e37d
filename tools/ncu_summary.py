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

if "--table" in sys.argv:   # one line per launch: the columns the round's notes quote
    cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
            ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
            ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_longsb"),
            ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
            ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
            ("sm__icc_request_hit_rate.pct", "icc_hit%")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ix = [(n, hdr.index(m)) for m, n in cols if m in hdr]
    print("units: " + ", ".join(f"{n}={rows[1][i]}" for n, i in ix if rows[1][i]))
    print(" ".join(f"{n:>10s}" for n, _ in ix) + "  kernel")
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(" ".join(f"{float(r[i] or 0):10.2f}" for _, i in ix) + "  " + r[kn].split("(")[0].replace("void ", ""))
