#!/bin/bash
mkdir -p gpurun_out
TDM_DEBUG_PLAN=1 TDM_DEBUG_ALIGNED_TAPS=1 TOPK=70 timeout 200 python tools/quick_profile.py mixed16 2>&1 | grep -E "resident forward|conv0\[tc\]|prob\[tc\]|conv2\[tc\]|conv4\[tc\]" | head -14
