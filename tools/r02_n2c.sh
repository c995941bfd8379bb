#!/bin/bash
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 bash tools/r02_slab.sh 1 &
wait
bash tools/r02_slab.sh 2
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/bench_fusion_tracker.py 2>&1 | head -1 | cut -c 1-420
