"""Per-phase timeline of the persistent decode kernel (csrc/mega.cu): %globaltimer stamps of every CTA at
phase entry / after the grid barrier / after the prologue / after the main loop / after the epilogue / after the
pre-attention barrier.  usage: GGB_MEGA=1 python tools/mega_timeline.py [--model llama3-8b] [--phases 12]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ggufb200.model import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--phases", type=int, default=12)
    ap.add_argument("--pos", type=int, default=256)
    args = ap.parse_args()
    os.environ["GGB_MEGA"] = "1"
    path = bench.model_path(args.model, args.ftype, 0xB200)
    eng = Engine(path, n_ctx=1024, use_graph=False)
    sl = eng.slots[0]
    sl.reset()
    sl.prefill([1] + list(range(300, 300 + args.pos - 1)))
    n_ph, G = sl._mega.n_phases, 148
    tl = torch.zeros(n_ph * G * 6, dtype=torch.int64, device=eng.dev)
    sl._mega.timeline = tl.data_ptr()
    sl.decode(3)
    torch.cuda.synchronize()
    t = tl.cpu().numpy().reshape(n_ph, G, 6).astype(np.float64)
    t0 = t[0, :, 0].min()
    names = ["QKV", "O", "GU", "DOWN"]
    print("phase  kind   enter(min)  bar_wait(avg/max)  prologue(avg)  main(avg/min/max)  epilogue(avg)  attn_bar(avg)  phase_total(us)")
    for ph in list(range(min(args.phases, n_ph))) + [n_ph - 1]:
        a = t[ph]
        kind = "HEAD" if ph == n_ph - 1 else names[ph % 4]
        enter = a[:, 0].min() - t0
        bw = a[:, 1] - a[:, 0]
        pro = a[:, 2] - a[:, 1]
        mn = a[:, 3] - a[:, 2]
        ep = a[:, 4] - a[:, 3]
        ab = (a[:, 5] - a[:, 4]) if kind == "QKV" else np.zeros(G)
        nxt = t[ph + 1][:, 0].min() if ph + 1 < n_ph else a[:, 4].max()
        print(f"{ph:4d}  {kind:5s} {enter/1e3:9.2f}   {bw.mean()/1e3:6.2f}/{bw.max()/1e3:6.2f}     {pro.mean()/1e3:6.2f}       "
              f"{mn.mean()/1e3:6.2f}/{mn.min()/1e3:6.2f}/{mn.max()/1e3:6.2f}    {ep.mean()/1e3:6.2f}       {ab.mean()/1e3:6.2f}       {(nxt - a[:, 0].min())/1e3:7.2f}")
    total = t[n_ph - 1][:, 4].max() - t0
    print(f"kernel total {total/1e3:.1f} us; sum over phases: barrier {np.sum([(t[p][:,1]-t[p][:,0]).mean() for p in range(n_ph)])/1e3:.1f}, "
          f"prologue {np.sum([(t[p][:,2]-t[p][:,1]).mean() for p in range(n_ph)])/1e3:.1f}, main {np.sum([(t[p][:,3]-t[p][:,2]).mean() for p in range(n_ph)])/1e3:.1f}, "
          f"epilogue {np.sum([(t[p][:,4]-t[p][:,3]).mean() for p in range(n_ph)])/1e3:.1f}")
    eng.close()


if __name__ == "__main__":
    main()
