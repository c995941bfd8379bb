#!/bin/bash
mkdir -p gpurun_out
bash tools/sanitize.sh r02
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","single_window_ms")}); print({k:round(d["e2e"][k],1) for k in ("value","serial_value","pageable_value")}); g=d["gpu_reference"]; print({k:g[k] for k in g if "abs_rel" in k or "iou" in k or "speedup" in k})
PY
