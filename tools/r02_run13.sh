#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mvsnet_gpu.py -m gpu -x -q 2>&1 | tail -3
TDM_DEBUG_PLAN=1 TOPK=70 timeout 200 python tools/quick_profile.py mixed16 > gpurun_out/r02_run13_ring8.txt 2>&1
grep -E "IS cin|resident forward|conv0\[tc\]|prob\[tc\]" gpurun_out/r02_run13_ring8.txt | head -24
