#!/bin/bash
# Generates the measurement artifacts of round 2 on the GPU box (run under gpurun from the repo root, ONE GPU).
# Outputs land in gpurun_out/; tools/summarise_profiles.py r02 turns them into profiles/*.md|csv here.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 400 gpurun_out/${R}_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
TOPK=80 timeout 200 python tools/quick_profile.py mixed16 > gpurun_out/${R}_kernels.txt 2>&1
timeout 300 python tools/bench_fusion_tracker.py > gpurun_out/${R}_fusion_tracker.txt 2>&1
timeout 300 python tools/bench_loop.py 120 > gpurun_out/${R}_loop.txt 2>&1
timeout 400 python tools/bench_config3.py --frames 1000 > gpurun_out/${R}_config3.txt 2>&1
# every launch of the bench command with its device time (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --inflight 1 > gpurun_out/${R}_bench_under_ncu.log 2>&1
# the dominant kernels, full sections
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cost_volume -c 3 -f -o gpurun_out/${R}_cost_volume \
    python tools/quick_profile.py mixed16 > gpurun_out/${R}_ncu_cv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_is -c 6 -f -o gpurun_out/${R}_conv_tc_is \
    python tools/quick_profile.py mixed16 > gpurun_out/${R}_ncu_is.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_raycast_shared|k_integrate_list|k_visible|k_allocate" --launch-skip 140 --launch-count 4 \
    -f -o gpurun_out/${R}_tsdf python tools/bench_fusion_tracker.py 36 > gpurun_out/${R}_ncu_tsdf.log 2>&1
ls -la gpurun_out | grep ${R}_ | head -40
