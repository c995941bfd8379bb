#!/bin/bash
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_fusion_gpu.py -m gpu -x -q 2>&1 | tail -4
bash tools/r02_slab.sh 2
