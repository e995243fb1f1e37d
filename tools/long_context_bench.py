"""bs=1 decode at a long sequence position: python tools/long_context_bench.py <n_ctx> <position>.  From model.LONG_SEQ positions on the\nattention runs as launches over the whole GPU (GGB_ATTN_SPLIT=0 keeps the cluster kernel)."""
import os, sys, json
sys.path.insert(0, "/root/repo")
import torch, bench
from ggufb200.model import Engine
ctx = int(sys.argv[1]); n = int(sys.argv[2])
eng = Engine(bench.model_path("llama3-8b", "Q4_K_M", 0xB200), n_ctx=ctx)
eng.warmup()
s0 = eng.slots[0]
s0.reset(); s0.prefill([1] + list(range(300, 300 + n - 1))); s0.decode(8)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.stream): e0.record(eng.stream)
s0.decode(32)
with torch.cuda.stream(eng.stream): e1.record(eng.stream)
torch.cuda.synchronize()
print(f"ctx {ctx} pos {n}: {e0.elapsed_time(e1)/32:.3f} ms/token  (GGB_ATTN_SPLIT={os.environ.get('GGB_ATTN_SPLIT', 'default')})")
eng.close()
