#!/bin/bash
# A/B helper: bash tests/env_sweep.sh VAR v1 v2 ...  -> training ms/step with VAR set to each value
var=$1; shift
for v in "$@"; do
  echo "$var=$v: $(env $var=$v HG_BENCH_NO_TRACE=1 timeout 100 python bench.py --workload train --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3))')"
done
