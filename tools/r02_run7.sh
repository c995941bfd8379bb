#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run7_pytest.log 2>&1; tail -5 gpurun_out/r02_run7_pytest.log
timeout 300 python tools/precision_ab.py mixed16:direct_conv00=0 > gpurun_out/r02_run7_precision.txt 2>&1; cat gpurun_out/r02_run7_precision.txt
TDM_IS_NPAD8=0 timeout 300 python tools/precision_ab.py 2>&1 | tail -1
TOPK=70 timeout 200 python tools/quick_profile.py mixed16 > gpurun_out/r02_run7_kernels.txt 2>&1; head -24 gpurun_out/r02_run7_kernels.txt
timeout 500 python bench.py --steps 60 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_run7_bench.json 2> gpurun_out/r02_run7_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_run7_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","single_window_ms")}); print({k:d["e2e"][k] for k in ("value","serial_value","pageable_value","pageable_serial_value")})
PY
tail -3 gpurun_out/r02_run7_bench.err
