#!/bin/bash
# Round-2 ncu evidence (one gpurun call; every ncu command runs only after the same command exited 0 without ncu):
#   1. --set full rows of the conv launches of one inference forward of BASELINE configs[1], stage 2 (ups1) to conv_post
#      (43 launches; stage 1's CTA-pair kernel is in r01_ncu_full.csv)                          -> gpurun_out/r02_ncu_full_raw.csv
#   2. launch list (gpu__time_duration.sum) of one eager training step                        -> gpurun_out/r02_train_launches.csv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
INF="python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline"
SKIP=${1:-83}      # launches to skip: 63 (the warm-up forward) + index of the first launch wanted (20 = ups1, 40 = stage 3)
COUNT=${2:-43}
TAG=${3:-r02_ncu_full_raw}
HG_BENCH_PROFILE=1 timeout 200 $INF > gpurun_out/r02_ncu_plain_infer.log 2>&1 &&
HG_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none -k regex:'conv1d_tc|resblock_pair|conv_post_tanh' -s $SKIP -c $COUNT \
  -o /tmp/r02_full_infer -f $INF > gpurun_out/r02_ncu_full_infer.log 2>&1
echo "ncu full inference exit=$?"
# the report is ~3 MB per launch: only its raw-page CSV travels back (gpurun_out/ is capped at 64 MiB)
ncu -i /tmp/r02_full_infer.ncu-rep --page raw --csv > gpurun_out/$TAG.csv 2> gpurun_out/r02_ncu_export.log
ls -la /tmp/r02_full_infer.ncu-rep gpurun_out/$TAG.csv
[ -n "$4" ] && exit 0      # a fourth argument: inference rows only
TRN="python tests/gpu_bringup_train.py profile 16"
timeout 200 $TRN > gpurun_out/r02_ncu_plain_train.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_train_launches.csv $TRN > gpurun_out/r02_ncu_train.log 2>&1
echo "ncu train launch list exit=$?"
ls -la gpurun_out/r02_train_launches.csv
