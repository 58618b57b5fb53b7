#!/bin/bash
# Performance bring-up: whole-generator timing + per-layer-shape timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/perf.log
rm -f gpurun_out/bringup.jsonl
run() {
  echo "=== $* ===" | tee -a gpurun_out/perf.log
  timeout 300 python tests/gpu_bringup.py "$@" >> gpurun_out/perf.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/perf.log
}
run time v1 64 1024 0
run time v1 1 256 0
run time v3 64 1024 0
run layers 64 1024
cp gpurun_out/bringup.jsonl gpurun_out/perf.jsonl
tail -c 1500 gpurun_out/perf.log
