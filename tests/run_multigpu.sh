#!/bin/bash
# Multi-GPU evidence for one box of N GPUs (gpurun --gpus N -- bash tests/run_multigpu.sh N):
#   1. data-parallel gradient equality + replica consistency on hardware (tests/dp_check.py under torchrun, via pytest)
#   2. the training metric at BASELINE configs[3]'s global batch 128 (per-GPU batch 128 / N) and at per-GPU batch 16
# Logs go to gpurun_out/r02_mgpu_<N>*.log
N=${1:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_DP" ]; then
  timeout 300 python -m pytest tests/test_gpu_named_configs.py -q -k cfg4 > gpurun_out/r02_mgpu_${N}_dpcheck.log 2>&1
  echo "dp_check rc=$?"; tail -3 gpurun_out/r02_mgpu_${N}_dpcheck.log; cat gpurun_out/dp_check_${N}gpu.log 2>/dev/null | grep -E "worst|bit-identical"
fi
run() {  # per-gpu batch, tag
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --workload train --steps 20 --warmup 5 --per-gpu-batch $1 --no-cpu-baseline 2> gpurun_out/r02_mgpu_${N}_train_b$1.err | grep '^{' | tail -1 > gpurun_out/r02_mgpu_${N}_train_b$1.json
  [ -s gpurun_out/r02_mgpu_${N}_train_b$1.json ] || { echo "N=$N per-gpu batch $1: no result"; tail -5 gpurun_out/r02_mgpu_${N}_train_b$1.err; return; }
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_mgpu_${N}_train_b$1.json").read())
print("N=$N per-gpu batch $1: %.3f ms/step, %.1f segments/s, global batch %d, launches/step %.0f, clocks %s" % (d["ms_per_step"], d["value"], d["config"]["global_batch"], d["gpu_launches"]/d["steps"], d["clocks"]))
PY
}
run $((128 / N))
[ "$N" != "8" ] && run 16
# the driver's scaling command: the default workload (inference + mel + training sub-record) on N ranks
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_mgpu_${N}_default.err | grep '^{' | tail -1 > gpurun_out/r02_mgpu_${N}_default.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_mgpu_${N}_default.json").read())
    t=d.get("train") or {}
    print("N=$N default line: %.3f ms/step, %.4g samples/s, e2e %.4g; train %.3f ms/step %.1f segments/s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], t.get("ms_per_step", float("nan")), t.get("value", float("nan"))))
except Exception as e:
    print("N=$N default line: no result", e)
PY
true
