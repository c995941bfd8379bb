#!/bin/bash
# the multi-GPU measurements of round 2 on ONE 8-GPU box: MVSNet windows (weak scaling, e2e included) and the Z-slab TSDF (strong scaling)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo8.txt 2>&1
tr() { n=$1; port=$2; shift 2; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
tr 8 29521 bench.py --gpus 8 --steps 60 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
CUDA_VISIBLE_DEVICES=0,1,2,3 tr 4 29522 bench.py --gpus 4 --steps 60 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_n{n}.json"))
        print(n, {k: round(d[k], 3) for k in ("value", "ms_per_step", "single_window_ms")}, {k: round(d["e2e"][k], 1) for k in ("value", "serial_value", "pageable_value")}, d["e2e"]["affinity"])
    except Exception as e:
        print(n, "bench failed", e)
PY
tail -2 gpurun_out/r02_bench_n8.err | cut -c 1-300
for mode in "" "--peer"; do
  tr 8 29523 tools/bench_slab_tsdf.py --frames 24 $mode > gpurun_out/r02_slab_n8$mode.txt 2> gpurun_out/r02_slab_n8$mode.err; cut -c 1-1000 gpurun_out/r02_slab_n8$mode.txt; tail -1 gpurun_out/r02_slab_n8$mode.err | cut -c 1-300
  CUDA_VISIBLE_DEVICES=0,1,2,3 tr 4 29524 tools/bench_slab_tsdf.py --frames 24 $mode > gpurun_out/r02_slab_n4$mode.txt 2> gpurun_out/r02_slab_n4$mode.err; cut -c 1-1000 gpurun_out/r02_slab_n4$mode.txt
done
