#!/bin/bash
# programmatic dependent launch with the trigger AFTER griddepcontrol.wait (one kernel of look-ahead): A/B
mkdir -p gpurun_out
for f in 0 1; do
  echo "== use_pdl=$f"
  TOPK=70 timeout 200 python tools/quick_profile.py mixed16 -1 use_pdl=$f 2>&1 | grep -E "Abs Rel|resident forward|sum of kernels" | head -12
  timeout 200 python bench.py --steps 60 --no-cpu-baseline --no-gpu-reference --opt use_pdl=$f > gpurun_out/r02_run17_f$f.json 2> gpurun_out/r02_run17_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_run17_f$f.json"))
print("use_pdl $f value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single", round(d["single_window_ms"],4), "e2e", round(d["e2e"]["value"],1))
PY
done
