#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the default bench line, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit=$?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench reference exit=$?"
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default exit=$?"
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], 'launches', d['gpu_launches'])
print('train', d['train']['value'], d['train']['ms_per_step'], d['train']['roofline']['frac'])
print('clocks', d['clocks']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
