#!/bin/bash
# Hardware bring-up driver: each stage in its own process with its own timeout.
#   gpurun --timeout 900 -- 'bash tests/run_bringup.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/bringup.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/bringup_gpu.txt 2>&1
run() {
  echo "=== $* ===" | tee -a gpurun_out/bringup.log
  timeout 240 python tests/gpu_bringup.py "$@" >> gpurun_out/bringup.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/bringup.log
}
: > gpurun_out/bringup.log
run mel
run conv 0
run conv 1
for mode in ${MODES:-0 1}; do
  run gen tiny 2 16 $mode
  run gen v1 1 32 $mode
done
tail -c 6000 gpurun_out/bringup.log
