#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mvsnet_gpu.py -m gpu -x -q -k "tma_staged or benchmark_config" 2>&1 | tail -4
for v in 3 5; do TDM_DEBUG_PLAN=1 TOPK=12 timeout 200 python tools/quick_profile.py mixed16 -1 cv_variant=$v 2>&1 | grep -E "cost_volume|resident forward|cv_tma|conv0.0" ; done > gpurun_out/r02_run8_cv_tma.txt 2>&1; cat gpurun_out/r02_run8_cv_tma.txt
