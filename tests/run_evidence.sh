#!/bin/bash
# Round evidence: bench lines (default + train + reference arm), torch-on-the-same-GPU comparison, DRAM traffic per
# launch of the inference conv kernels (ncu), launch list of the default bench command.
#   gpurun --timeout 1800 -- 'bash tests/run_evidence.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/torch_gpu_compare.jsonl
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default exit=$?"
timeout 400 python bench.py --workload train --steps 20 > gpurun_out/bench_train.log 2>&1; echo "bench train exit=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench reference exit=$?"
timeout 600 python tests/torch_gpu_compare.py > gpurun_out/torch_gpu_compare.log 2>&1; echo "torch compare exit=$?"
HG_BENCH_PROFILE=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:'conv1d_tc|resblock_pair|conv_post_tanh|ncl_to_nlc' -c 128 --csv --log-file gpurun_out/r01_traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit=$?"
tail -n 1 gpurun_out/bench_default.log | cut -c 1-300
tail -n 1 gpurun_out/bench_train.log | cut -c 1-300
cat gpurun_out/torch_gpu_compare.jsonl
