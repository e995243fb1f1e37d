#!/bin/bash
# One GPU call that regenerates the round's evidence under gpurun_out/ (copied into profiles/ by hand after reading):
#   1. bench line as the driver runs it;  2. ncu launch list of the same command;  3. ncu --set full of one layer's GEMV
#   launches + the lm-head, and of the decode attention;  4. DRAM bytes of every GEMV launch of one token of the REAL
#   32-layer model (-> roofline.traffic).  Each ncu run follows a plain run of the same command that exited 0.
set -x
R=${1:-r02}
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $O/${R}_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_ll.log 2>&1
python tools/profile_step.py --layers 4 --steps 4 > $O/plain_ps.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ggb_dq_gemv|attn_decode' -s 64 -c 21 -o $O/${R}_step_full \
    python tools/profile_step.py --layers 4 --steps 4 > $O/ncu_full.log 2>&1
python tools/profile_step.py --layers 32 --steps 2 > $O/plain_ps32.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ggb_dq_gemv -s 514 -c 129 --csv \
    --log-file $O/${R}_gemv_dram_one_token.csv python tools/profile_step.py --layers 32 --steps 2 > $O/ncu_dram.log 2>&1
tail -2 $O/plain_ps32.log; cat $O/${R}_bench.json
