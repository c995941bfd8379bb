#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r02_run21_bench.json 2> gpurun_out/r02_run21_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run21_bench.json"))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single", round(d["single_window_ms"],4), "e2e", {k: round(v,1) for k,v in d["e2e"].items() if k.endswith("value")})
PY
