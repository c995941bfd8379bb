#!/bin/bash
# Timing experiment: a third of the MMA instructions. Needs a library built with
#   make -C tandem_b200/csrc EXTRA_NVFLAGS=-DTDM_TIMING_EXPERIMENTS   (results are WRONG by construction; timing only)
TDM_DEBUG_ALIGNED_TAPS=2 TOPK=70 timeout 200 python tools/quick_profile.py mixed16 2>&1 | grep -E "resident forward|conv0\[tc\]|prob\[tc\]|conv2\[tc\]|conv4\[tc\]" | head -14
