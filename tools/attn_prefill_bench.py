"""prefill attention alone: 2048 tokens, Llama-3-8B head geometry, both kernels"""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from ggufb200 import cabi
L = cabi.lib()
dev = torch.device("cuda", 0)
n_head, n_kv, hd = 32, 8, 128
for T, pos0 in ((2048, 0), (512, 0), (2048, 6144)):
    q = torch.randn(T, n_head * hd, device=dev)
    kc = (torch.randn(pos0 + T, n_kv * hd, device=dev) * 0.5).to(torch.float16)
    vc = torch.randn(pos0 + T, n_kv * hd, device=dev).to(torch.float16)
    out = torch.zeros(T, n_head * hd, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    for mode in ("1", "0"):
        os.environ["GGB_ATTN_PREFILL_TC"] = mode
        for _ in range(3):
            cabi.check(L.ggb_attn_prefill(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), T, pos0, n_head, n_kv, hd, out.data_ptr(), s))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cabi.check(L.ggb_attn_prefill(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), T, pos0, n_head, n_kv, hd, out.data_ptr(), s))
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 10, out.clone())
    flop = 4.0 * n_head * hd * (T * pos0 + T * (T + 1) / 2)
    d = (res["1"][1] - res["0"][1]).abs().max().item()
    print(f"T={T} pos0={pos0}: tcgen05 {res['1'][0]*1e3:.0f} us ({flop/res['1'][0]/1e9:.0f} TFLOP/s)  mma.sync {res['0'][0]*1e3:.0f} us ({flop/res['0'][0]/1e9:.0f} TFLOP/s)  max|diff| {d:.2e}")
