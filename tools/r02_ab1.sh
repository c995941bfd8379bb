#!/bin/bash
# round-2 A/B #1: shared-memory budget of the tcgen05 tile planner (two CTAs per SM) with 8 windows in flight
mkdir -p gpurun_out
for kb in 225 110 72; do
  timeout 200 python bench.py --steps 60 --no-cpu-baseline --opt tc_smem_kb=$kb > gpurun_out/r02_ab1_smem$kb.json 2> gpurun_out/r02_ab1_smem$kb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_ab1_smem$kb.json"))
print("smem_kb $kb value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single", round(d["single_window_ms"],4), "e2e", round(d["e2e"]["value"],1))
PY
done
TDM_DEBUG_PLAN=1 TOPK=80 timeout 200 python tools/quick_profile.py mixed16 -1 > gpurun_out/r02_ab1_plan225.txt 2>&1
TDM_DEBUG_PLAN=1 TOPK=80 timeout 200 python tools/quick_profile.py mixed16 -1 tc_smem_kb=110 > gpurun_out/r02_ab1_plan110.txt 2>&1
grep "resident forward" gpurun_out/r02_ab1_plan*.txt
