#!/bin/bash
# kernel timeline of the graph-replayed training step -> gpurun_out/train_trace.json (kernels only)
cd "$(dirname "$0")/.."
timeout 300 python tests/gpu_bringup_train.py trace ${1:-16} 2>&1 | tail -2
python - <<PY
import json
d = json.load(open("gpurun_out/train_trace_full.json"))
ev = [e for e in d["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
json.dump([{"n": e["name"][:60], "ts": e["ts"], "dur": e["dur"], "s": e["args"].get("stream")} for e in ev],
          open("gpurun_out/train_trace.json", "w"))
PY
rm -f gpurun_out/train_trace_full.json
