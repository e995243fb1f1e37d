"""One eager batched decode step (no graph) for a launch list: run under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/batch_launches.py --batch 16
and summarise with --summarise gpurun_out/x.csv (per kernel name and grid: launches, total and mean microseconds of the LAST step)."""
import argparse
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path, last_n):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"].split("(")[0][:60], r.get("Grid Size", ""), us))
    rows = rows[-last_n:] if last_n else rows
    agg = collections.OrderedDict()
    for name, grid, us in rows:
        k = (name, grid)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:60s} {grid:>14s} n={n:4d} total={us:9.1f} us mean={us / n:7.2f} us  {100 * us / tot:5.1f} %")
    print(f"{'sum':60s} {'':>14s} n={len(rows):4d} total={tot:9.1f} us")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--summarise")
    ap.add_argument("--last", type=int, default=0)
    args = ap.parse_args()
    if args.summarise:
        return summarise(args.summarise, args.last)
    import bench
    from ggufb200.model import Engine
    B = args.batch
    eng = Engine(bench.model_path(args.model, args.ftype, 0xB200), n_ctx=256, n_slots=B, use_graph=False)
    for s in range(B):
        eng.slots[s].reset()
        eng.slots[s].prefill([1] + [300 + s + j for j in range(31)])
    last = [eng.slots[s].read_last_token() for s in range(B)]
    for _ in range(2):
        last = eng.batch.step([(s, last[s], eng.slots[s].n_past) for s in range(B)])
    print("launches per step (upper bound)", eng.batch.launches_per_step(B))


if __name__ == "__main__":
    main()
