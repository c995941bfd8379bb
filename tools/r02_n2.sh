#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
bash tools/r02_slab.sh 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n2.json"))
print({k:d[k] for k in ("value","ms_per_step","single_window_ms")}); print(d["e2e"])
PY
tail -3 gpurun_out/r02_bench_n2.err
