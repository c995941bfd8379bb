#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mvsnet_gpu.py -m gpu -x -q 2>&1 | tail -3
TDM_DEBUG_PLAN=1 TOPK=70 timeout 200 python tools/quick_profile.py mixed16 > gpurun_out/r02_run11_persist1.txt 2>&1
TDM_IS_PERSIST=0 TOPK=70 timeout 200 python tools/quick_profile.py mixed16 > gpurun_out/r02_run11_persist0.txt 2>&1
grep -E "resident forward|conv0\[tc\]|prob\[tc\]|conv2\[tc\]|conv4\[tc\]" gpurun_out/r02_run11_persist1.txt | head -16
echo ---- one tile per CTA
grep -E "resident forward|conv0\[tc\]|prob\[tc\]|conv2\[tc\]|conv4\[tc\]" gpurun_out/r02_run11_persist0.txt | head -16
timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','single_window_ms')}, round(d['e2e']['value'],1))"
