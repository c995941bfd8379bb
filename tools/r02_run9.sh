#!/bin/bash
mkdir -p gpurun_out
for k in 4 6 8; do
  timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-gpu-reference --e2e-inflight $k > gpurun_out/r02_run9_e2e$k.json 2> gpurun_out/r02_run9_e2e$k.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_run9_e2e$k.json"))
print("e2e-inflight $k:", round(d["value"],1), round(d["ms_per_step"],4), round(d["single_window_ms"],4), {k:round(d["e2e"][k],1) for k in ("value","serial_value","pageable_value","pageable_serial_value")})
PY
done
