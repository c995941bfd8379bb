#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/precision_ab.py mixed16:cv_variant=0 mixed16:cv_variant=1 mixed16:cv_variant=2 mixed16:cv_variant=4 mixed16:use_tc=0 mixed16:use_tc=0,cv_variant=1 mixed16:use_is=0 > gpurun_out/r02_ab2_precision.txt 2>&1
cat gpurun_out/r02_ab2_precision.txt
timeout 300 python -m pytest tests/test_mvsnet_gpu.py tests/test_fusion_gpu.py tests/test_tracker_gpu.py -m gpu -x -q 2>&1 | tail -5
