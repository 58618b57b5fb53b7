#!/bin/bash
# One compute-sanitizer tool per gpurun call (B200_PROFILING.md: several tools in one call have wedged the GPU).
#   tests/run_sanitizer.sh memcheck|racecheck|synccheck|initcheck [families...]
# Writes gpurun_out/sanitizer_<tool>.log; exit code = the tool's.
tool=${1:-memcheck}; shift
fams=${@:-conv conv2 pair wgrad mel ends}
mkdir -p gpurun_out
python tests/sanitize_cases.py $fams > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 50 python tests/sanitize_cases.py $fams > gpurun_out/sanitizer_$tool.log 2>&1
rc=$?
echo "compute-sanitizer $tool rc=$rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|ok$|all requested" gpurun_out/sanitizer_$tool.log | tail -30
exit $rc
