#!/bin/bash
# fused soft-argmin tail: bit-identity test, per-kernel profile, A/B of the bench legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mvsnet_gpu.py -x -q -k "regress_in_prob or benchmark_config or known_answer" 2>&1 | tail -5
for f in 1 0; do
  echo "== fused_regress=$f"
  TOPK=70 timeout 200 python tools/quick_profile.py mixed16 -1 fused_regress=$f 2>&1 | grep -E "resident forward|prob|regress|sum of kernels" | head -12
  timeout 200 python bench.py --steps 60 --no-cpu-baseline --no-gpu-reference --opt fused_regress=$f > gpurun_out/r02_run16_f$f.json 2> gpurun_out/r02_run16_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_run16_f$f.json"))
print("fused_regress $f value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single", round(d["single_window_ms"],4), "e2e", round(d["e2e"]["value"],1))
PY
done
