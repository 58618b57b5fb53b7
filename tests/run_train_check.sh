#!/bin/bash
# training-path check after a scheduling / kernel change: GPU tests of the training step, bench, kernel timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/pytest_train.log 2>&1; echo "pytest train exit=$?"
tail -n 3 gpurun_out/pytest_train.log
timeout 400 python bench.py --workload train --steps 20 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1; echo "bench train exit=$?"
tail -n 1 gpurun_out/bench_train.log | cut -c 1-330
bash tests/run_trace.sh 16 > gpurun_out/trace.log 2>&1; echo "trace exit=$?"
python tests/trace_agg.py gpurun_out/train_trace.json 1 2>&1 | head -24
