#!/bin/bash
mkdir -p gpurun_out
bash tools/r02_slab.sh 2
