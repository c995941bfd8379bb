#!/bin/bash
# Final confirmation of the round-2 tree: whole GPU suite, smoke, the driver-contract bench line, a memcheck leg over the new sampler paths.
R=r02
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 300 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "single_window_ms")}, {k: round(v, 1) for k, v in d["e2e"].items() if k.endswith("value")})
except Exception as e:
    print("bench line unreadable:", e)
PY
CS="compute-sanitizer --error-exitcode 9 --print-limit 10 --launch-timeout 60"
timeout 150 $CS --tool memcheck python -m pytest tests/test_fusion_gpu.py -m gpu -q -x -p no:cacheprovider -k "integrate_and_render or negative_coordinates or pixel_partitioned" > gpurun_out/${R}_sanitize_mem_fusion.log 2>&1
echo "memcheck: exit $? | $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/${R}_sanitize_mem_fusion.log | tr '\n' ' ')"
