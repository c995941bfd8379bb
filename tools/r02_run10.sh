#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mvsnet_gpu.py tests/test_fusion_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/precision_ab.py mixed16:fused_select=0 2>&1 | tail -3
