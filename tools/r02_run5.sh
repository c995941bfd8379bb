#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/precision_ab.py > gpurun_out/r02_run5_precision.txt 2>&1; cat gpurun_out/r02_run5_precision.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run5_pytest.log 2>&1; tail -5 gpurun_out/r02_run5_pytest.log
timeout 500 python bench.py --steps 60 > gpurun_out/r02_run5_bench.json 2> gpurun_out/r02_run5_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_run5_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","single_window_ms")}); print(d["e2e"]); print(d["gpu_reference"]); print(d["cpu_baseline"])
PY
tail -3 gpurun_out/r02_run5_bench.err
