#!/bin/bash
# Ray-cast instruction diet (block de-duplication in the sampler, division-free hash, select-form rounding, hoisted pixel terms):
# whole GPU suite (incl. the strict variant-vs-variant bit-identity test), the TSDF A/B table, config 3, one full ncu capture.
R=r02c
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 200 python tools/bench_fusion_tracker.py > gpurun_out/${R}_fusion_tracker.txt 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c_fusion_tracker.txt").readline())
    for k, v in d["ab"].items():
        print(f"{k:34s}", {a: round(b, 4) for a, b in v.items()})
except Exception as e:
    print("unreadable:", e)
PY
timeout 150 python tools/bench_config3.py --frames 400 > gpurun_out/${R}_config3.txt 2>&1; tail -2 gpurun_out/${R}_config3.txt | cut -c1-500
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_raycast_shared" --launch-skip 30 --launch-count 1 \
    -f -o gpurun_out/${R}_tsdf python tools/bench_fusion_tracker.py 24 > gpurun_out/${R}_ncu_tsdf.log 2>&1
ls -la gpurun_out | grep ${R}_ | head -40
