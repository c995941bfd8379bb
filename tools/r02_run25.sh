#!/bin/bash
# Pointer-caching block cache + hoisted half-voxel in the ray-cast; source-level ncu capture of the cost-volume kernels; slab bench N=1
R=r02d
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fusion_gpu.py tests/test_reference_gpu.py -m gpu -q -x > gpurun_out/${R}_pytest_fusion.log 2>&1; tail -3 gpurun_out/${R}_pytest_fusion.log
timeout 200 python tools/bench_fusion_tracker.py > gpurun_out/${R}_fusion_tracker.txt 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02d_fusion_tracker.txt").readline())
    for k, v in d["ab"].items():
        print(f"{k:34s}", {a: round(b, 4) for a, b in v.items()})
except Exception as e:
    print("unreadable:", e)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cost_volume -c 3 -f -o gpurun_out/${R}_cost_volume \
    python tools/quick_profile.py mixed16 > gpurun_out/${R}_ncu_cv.log 2>&1; tail -2 gpurun_out/${R}_ncu_cv.log
timeout 200 python tools/bench_slab_tsdf.py --frames 24 > gpurun_out/${R}_slab_n1.txt 2> gpurun_out/${R}_slab_n1.err; cut -c 1-700 gpurun_out/${R}_slab_n1.txt
ls -la gpurun_out | grep ${R}_ | head
