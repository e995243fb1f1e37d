#!/bin/bash
# GPU call 1 (round 2): tests, prefetch sweep, A/B of the interleaved main loop, in-situ timeline
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_tests.log
tail -3 gpurun_out/c1_tests.log
P=$PWD/llama-gguf-inference_b200
timeout 900 python tools/pf_sweep.py "" "GGB_PF_TAIL_KB=64" "GGB_PF_TAIL_KB=128" "GGB_PF_TAIL_KB=256" \
   "GGB_PF_ATTN_KB=128" "GGB_PF_ATTN_KB=256" "GGB_PF_TAIL_KB=128 GGB_PF_ATTN_KB=128" "GGB_PF_TAIL_KB=128 GGB_PF_ATTN_KB=256" \
   "GGB_PF_TAIL_KB=256 GGB_PF_ATTN_KB=256" "GGB_PF_TAIL_KB=128 GGB_PF_WHEN=1" "GGB_PF_TAIL_KB=128 GGB_PF_WHEN=2" \
   "GGB_PF_TAIL_KB=128 GGB_PF_ATTN_KB=256 GGB_PF_WHEN=6" > gpurun_out/c1_sweep.txt 2>&1
cat gpurun_out/c1_sweep.txt
GGB_LIB_PATH=$P/libggb_noil.so timeout 300 python tools/pf_sweep.py "" > gpurun_out/c1_sweep_noil.txt 2>&1
cat gpurun_out/c1_sweep_noil.txt
timeout 300 python tools/gemv_bench.py > gpurun_out/c1_gemv_bench.txt 2>&1
GGB_LIB_PATH=$P/libggb_noil.so timeout 300 python tools/gemv_bench.py > gpurun_out/c1_gemv_bench_noil.txt 2>&1
GGB_LIB_PATH=$P/libggb_tl.so timeout 300 python tools/step_timeline.py > gpurun_out/c1_timeline_default.txt 2>&1
GGB_PF_TAIL_KB=128 GGB_PF_ATTN_KB=256 GGB_LIB_PATH=$P/libggb_tl.so timeout 300 python tools/step_timeline.py > gpurun_out/c1_timeline_pf.txt 2>&1
tail -5 gpurun_out/c1_gemv_bench.txt gpurun_out/c1_gemv_bench_noil.txt
echo done
