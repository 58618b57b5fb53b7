#!/bin/bash
# A/B of the scale discriminators' lane priorities (HG_DISC_BOOST = boosts of MSD 0,1,2; lower = more urgent)
for b in ${@:-"-2,-1,0" "-2,-1,-1" "0,0,0" "-1,-1,-1" "0,-1,-1" "0,0,-1" "-1,0,-2"}; do
  echo "boost $b: $(HG_DISC_BOOST=$b HG_BENCH_NO_TRACE=1 timeout 100 python bench.py --workload train --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"])')"
done
