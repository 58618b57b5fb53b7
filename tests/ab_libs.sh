#!/bin/bash
# A/B of two BUILDS of the library inside one gpurun call (same box, alternating runs): how round 2 measured the
# epilogue flavours and the plain pair instantiation without the box-to-box spread (+-3 % under the power cap).
#
# Prepare both builds in the build container: for each variant, build (python -c "import hifigan_b200._lib as l; l.lib()")
# and copy  hifi-gan_b200/libhifigan_b200.so, hifi-gan_b200/libhifigan_b200.srchash  and the csrc files that differ
# into a directory (ab_a/, ab_b/; the source files keep the staleness hash consistent, so nothing rebuilds on the box).
#
#   gpurun -- 'bash tests/ab_libs.sh ab_a ab_b 3 python tests/gpu_bringup.py time v1 64 1024 1'
#
# prints the command's JSON lines per variant, `rounds` times, alternating.
A=$1; B=$2; ROUNDS=$3; shift 3
cd "$(dirname "$0")/.."
use() {
  cp "$1"/libhifigan_b200.so "$1"/libhifigan_b200.srchash hifi-gan_b200/
  for f in "$1"/*.cu "$1"/*.cuh; do [ -e "$f" ] && cp "$f" hifi-gan_b200/csrc/; done
}
for r in $(seq 1 "$ROUNDS"); do
  for v in "$A" "$B"; do
    use "$v"
    echo "== $v (round $r)"
    timeout 300 "$@" 2>&1 | grep '^{' | cut -c1-400
  done
done
