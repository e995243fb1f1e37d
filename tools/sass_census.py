"""Per-kernel census of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one
(/opt/skills/guides/B200_PROFILING.md): tcgen05 (UTC*MMA, LDTM/STTM), TMA (UBLKCP, UTMALDG/UTMASTG), legacy tensor paths
(HMMA, IMMA), integer dots (IDP), cluster / PDL / mbarrier machinery.  Runs on the CPU-only box (cuobjdump).
usage: python tools/sass_census.py [libggufb200.so] > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "llama-gguf-inference_b200", "libggufb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "IMMA", "IDP", "SYNCS", "UCGABAR", "ACQBULK", "LDGSTS", "REDUX", "DADD"]
per = collections.OrderedDict()
name = None
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and name:
        op = m.group(1)
        per[name]["_n"] += 1
        for c in COLS:
            if op.startswith(c):
                per[name][c] += 1
demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.split("\n")
print(f"# {os.path.basename(lib)}: SASS census per kernel (cuobjdump -sass); instr = SASS instructions of the kernel")
print("# " + " ".join(f"{c:>7s}" for c in ["instr"] + COLS) + "  kernel")
for (k, c), d in zip(per.items(), demangle):
    d = re.sub(r"\(.*", "", d)
    print("  " + " ".join(f"{c[x]:7d}" for x in ["_n"] + COLS) + "  " + d)
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# total: " + ", ".join(f"{x} {tot[x]}" for x in COLS if tot[x]))
