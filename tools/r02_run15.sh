#!/bin/bash
# Timing experiment: no epilogue work / both. Needs a library built with
#   make -C tandem_b200/csrc EXTRA_NVFLAGS=-DTDM_TIMING_EXPERIMENTS   (results are WRONG by construction; timing only)
for v in 3 4; do echo "== dbg $v (3: no epilogue work; 4: no epilogue work + 1/3 of the MMAs)"; TDM_DEBUG_ALIGNED_TAPS=$v TOPK=70 timeout 200 python tools/quick_profile.py mixed16 2>&1 | grep -E "resident forward|conv0\[tc\]|prob\[tc\]|conv2\[tc\]" | head -10; done
