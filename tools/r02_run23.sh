#!/bin/bash
# Round-2 re-entry run: the whole GPU suite on HEAD (occupancy shortcut in the ray-cast on by default), smoke, bench, the TSDF A/B
# (occ_skip on / off), config 3, the per-kernel table, the launch list and one full ncu capture of the new ray-cast.
R=r02b
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 200 python tools/bench_fusion_tracker.py > gpurun_out/${R}_fusion_tracker.txt 2>&1; head -c 1500 gpurun_out/${R}_fusion_tracker.txt | tr ',' '\n' | grep -E "raycast_ms|default|no_occ" | head -12
timeout 300 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02b_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "single_window_ms")}, {k: round(v, 1) for k, v in d["e2e"].items() if k.endswith("value")})
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 150 python tools/bench_config3.py --frames 400 > gpurun_out/${R}_config3.txt 2>&1; tail -2 gpurun_out/${R}_config3.txt | cut -c1-600
TOPK=80 timeout 120 python tools/quick_profile.py mixed16 > gpurun_out/${R}_kernels.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --inflight 1 > gpurun_out/${R}_bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_raycast_shared" --launch-skip 30 --launch-count 2 \
    -f -o gpurun_out/${R}_tsdf python tools/bench_fusion_tracker.py 24 > gpurun_out/${R}_ncu_tsdf.log 2>&1
CS="compute-sanitizer --error-exitcode 9 --print-limit 10 --launch-timeout 60"
timeout 150 $CS --tool memcheck python -m pytest tests/test_fusion_gpu.py -m gpu -q -x -p no:cacheprovider -k "integrate_and_render or negative_coordinates" > gpurun_out/${R}_sanitize_mem.log 2>&1
echo "memcheck: exit $? | $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/${R}_sanitize_mem.log | tr '\n' ' ')"
ls -la gpurun_out | grep ${R}_ | head -40
