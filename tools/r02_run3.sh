#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/precision_ab.py > gpurun_out/r02_run3_precision.txt 2>&1; cat gpurun_out/r02_run3_precision.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run3_pytest.log 2>&1; tail -5 gpurun_out/r02_run3_pytest.log
timeout 400 python bench.py --steps 60 > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err; tail -c 1500 gpurun_out/r02_run3_bench.json; tail -3 gpurun_out/r02_run3_bench.err
timeout 300 python tools/bench_fusion_tracker.py > gpurun_out/r02_run3_fusion_tracker.txt 2>&1; tail -4 gpurun_out/r02_run3_fusion_tracker.txt | cut -c 1-600
