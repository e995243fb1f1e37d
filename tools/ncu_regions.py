"""Where a kernel's samples fall: per profiled launch, the contiguous SASS regions that executed at all, with their
execution counts, stall samples and instruction-fetch stalls, mapped back to source lines (needs -lineinfo).
usage: python tools/ncu_regions.py x.ncu-rep kernel.cubin mangled_kernel_name [launch indices ...]"""
import csv
import re
import subprocess
import sys

rep, cubin, kname = sys.argv[1:4]
which = [int(a) for a in sys.argv[4:]]
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text." + kname)][0]
cur, instr = None, []
for l in sass[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        instr.append((cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hidx[0]]
ci = {n: i for i, n in enumerate(h)}
tables = []
for k in range(len(hidx)):
    s, e = hidx[k], (hidx[k + 1] - 1 if k + 1 < len(hidx) else len(rows))
    body = [r for r in rows[s + 1:e] if len(r) == len(h)]
    if len(body) == len(instr):
        tables.append(body)
tables = tables[::2] if len(tables) > 1 and len(tables) % 2 == 0 and tables[0] == tables[1] else tables
print(f"{len(instr)} SASS instructions, {len(tables)} launch tables")
for k, body in enumerate(tables):
    if which and k not in which:
        continue
    ex = [int(r[ci["Instructions Executed"]] or 0) for r in body]
    ns = [int(r[ci["# Samples"]] or 0) for r in body]
    noi = [int(r[ci["stall_no_inst"]] or 0) for r in body]
    print(f"-- launch {k}: executed at least once {sum(1 for e in ex if e)} instructions ({sum(1 for e in ex if e) * 16 / 1024:.1f} KB); "
          f"samples {sum(ns)}, stall_no_inst {sum(noi)} ({100 * sum(noi) / max(1, sum(ns)):.0f} %)")
    i = 0
    while i < len(body):
        if ex[i] > 0:
            j = i
            while j < len(body) and (ex[j] > 0 or (j + 1 < len(body) and ex[j + 1] > 0)):
                j += 1
            cnt = sorted(e for e in ex[i:j] if e)
            files = {}
            for c, _ in instr[i:j]:
                if c:
                    files.setdefault(c[0], []).append(c[1])
            desc = " ".join(f"{f}:{min(v)}-{max(v)}" for f, v in files.items() if f.endswith((".cu", ".cuh")))
            print(f"   [{i:5d},{j:5d}) n={j - i:4d} exec/instr median {cnt[len(cnt) // 2]:7d} samples {sum(ns[i:j]):5d} no_inst {sum(noi[i:j]):5d}  {desc}")
            i = j
        else:
            i += 1
