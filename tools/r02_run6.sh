#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_reference_gpu.py tests/test_loop_gpu.py tests/test_front_gpu.py -m gpu -x -q > gpurun_out/r02_run6_pytest.log 2>&1; tail -4 gpurun_out/r02_run6_pytest.log
timeout 300 python tools/bench_fusion_tracker.py > gpurun_out/r02_run6_fusion_tracker.txt 2>&1; head -1 gpurun_out/r02_run6_fusion_tracker.txt | cut -c 1-700
TDM_FAST_DIV=0 timeout 300 python tools/bench_fusion_tracker.py 2>&1 | head -1 | cut -c 1-420
