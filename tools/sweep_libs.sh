#!/bin/bash
# A/B sweep over experiment builds: tools/sweep_libs.sh name1 name2 ... (libggb_v_<name>.so next to the product library)
L=/root/repo/llama-gguf-inference_b200
args=()
for n in "$@"; do args+=("GGB_LIB_PATH=$L/libggb_v_$n.so"); done
bash tools/sweep.sh "${args[@]}"
