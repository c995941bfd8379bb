#!/bin/bash
mkdir -p gpurun_out
timeout 250 python tools/pdl_sweep.py pdl_min_smem_kb=100
timeout 250 python tools/pdl_sweep.py pdl_min_smem_kb=150
for kb in 120; do
  timeout 200 python bench.py --steps 60 --no-cpu-baseline --no-gpu-reference --opt use_pdl=1 --opt pdl_min_smem_kb=$kb > gpurun_out/r02_run20_kb$kb.json 2> gpurun_out/r02_run20_kb$kb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_run20_kb$kb.json"))
print("pdl kb $kb value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single", round(d["single_window_ms"],4), "e2e", {k: round(v,1) for k,v in d["e2e"].items() if k.endswith("value")})
PY
done
