#!/bin/bash
# Llama-3-70B Q4_K_M tensor-parallel decode at N = the GPU counts given (default "1 2 4"), one box.
# usage: tools/tp70b.sh [steps] [N ...]   -> gpurun_out/tp70b_N.json
set -u
TAG=${TAG:-}   # prefix of the output files, e.g. TAG=nccl_ GGB_TP_EXCHANGE=nccl tools/tp70b.sh ...
mkdir -p gpurun_out
STEPS=${1:-128}; shift || true
NS=${*:-1 2 4}
df -h /dev/shm | tail -1; free -g | sed -n 2p; nproc
avail=$(df --output=avail -BG /dev/shm | tail -1 | tr -dc 0-9)
if [ "$avail" -lt 50 ]; then echo "not enough tmpfs for the 42.5 GB synthetic model ($avail GB)"; exit 1; fi
t0=$(date +%s)
python -c "import bench; print(bench.model_path('llama3-70b','Q4_K_M',0xB200))" || exit 1
echo "model written in $(( $(date +%s) - t0 )) s"
for n in $NS; do
  t0=$(date +%s)
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --model llama3-70b --steps $STEPS --warmup 8 --no-cpu > gpurun_out/tp70b_${TAG}$n.json 2> gpurun_out/tp70b_${TAG}$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --model llama3-70b --gpus $n --steps $STEPS --warmup 8 > gpurun_out/tp70b_${TAG}$n.json 2> gpurun_out/tp70b_${TAG}$n.err
  fi
  echo "N=$n rc=$? $(( $(date +%s) - t0 )) s"; tail -c 600 gpurun_out/tp70b_${TAG}$n.json; tail -3 gpurun_out/tp70b_${TAG}$n.err
done
