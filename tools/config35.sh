#!/bin/bash
# BASELINE.json configs 3 and 5 on one B200:
#   3: Llama-3-8B Q8_0 and Q6_K, 2048-token prefill (tcgen05 dequant-GEMM path) + bs=1 decode
#   5: Llama-3-8B Q4_K_M, 16 concurrent streaming requests (batched decode) through the llama-server process
mkdir -p gpurun_out
for ft in Q8_0 Q6_K; do
  echo "== config 3: $ft prefill"; timeout 600 python tools/prefill_bench.py --ftype $ft --tokens 2048 2>&1 | tail -8 | tee gpurun_out/config3_prefill_$ft.txt
  echo "== config 3: $ft decode"; timeout 600 python bench.py --ftype $ft --steps 256 --warmup 8 --no-cpu > gpurun_out/config3_decode_$ft.json 2> gpurun_out/config3_decode_$ft.err; tail -c 400 gpurun_out/config3_decode_$ft.json
done
echo "== config 5"; timeout 900 python tools/serve_bench.py --concurrent 1,4,16 --max-tokens 128 2>&1 | tail -5
