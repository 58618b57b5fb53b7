#!/bin/bash
# ncu --set full of the training step's heaviest tensor-core launches (one eager step between cudaProfilerStart/Stop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count
  timeout 280 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 \
    -o gpurun_out/r01_full_$1 -f python tests/gpu_bringup_train.py profile 16 > gpurun_out/ncu_full_$1.log 2>&1
  echo "$1 exit=$?"
}
cap train_wgrad 'wgrad_tc_kernel' 0 2
cap train_conv1024 'conv1d_tc_kernel<64, 256>' 22 2
cap train_firstbwd 'disc_first_bwd_kernel' 4 2
ls -la gpurun_out/r01_full_train_*.ncu-rep
