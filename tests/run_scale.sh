#!/bin/bash
# scaling evidence on one multi-GPU box: training step at N = 4, 8 and the default line (inference + train summary) at N = 8
#   gpurun --gpus 8 --timeout 900 -- 'bash tests/run_scale.sh [train-only]'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # N, log, extra args
  n=$1; log=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n "$@" > gpurun_out/$log 2>&1
  echo "N=$n $* exit=$?"; tail -n 1 gpurun_out/$log | cut -c 1-260
}
run 4 bench_train_4gpu.log --workload train --steps 20 --warmup 3
run 8 bench_train_8gpu.log --workload train --steps 20 --warmup 3
[ "$1" = "train-only" ] || run 8 bench_default_8gpu.log --steps 10 --warmup 3
