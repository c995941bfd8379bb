#!/bin/bash
# slab TSDF: N ranks (N = number of visible GPUs), new path vs the round-1 path
N=${1:-1}
mkdir -p gpurun_out
run() {
  if [ "$N" = "1" ]; then timeout 300 python tools/bench_slab_tsdf.py "$@"
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_slab_tsdf.py "$@"; fi
}
run --frames 24 > gpurun_out/r02_slab_n$N.txt 2> gpurun_out/r02_slab_n$N.err; cut -c 1-1100 gpurun_out/r02_slab_n$N.txt; tail -3 gpurun_out/r02_slab_n$N.err
if [ "$N" != "1" ]; then
  run --frames 24 --peer > gpurun_out/r02_slab_n${N}_peer.txt 2>> gpurun_out/r02_slab_n$N.err; cut -c 1-1100 gpurun_out/r02_slab_n${N}_peer.txt; tail -2 gpurun_out/r02_slab_n$N.err | cut -c 1-300
fi
