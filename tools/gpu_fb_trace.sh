#!/bin/bash
# phase timeline (clock64) of one steady-state CTA of k_front and k_back, 1024 x 10 s.  Needs a -DGTCRNB_FB_TRACE build:
#   GTCRNB_LIB=gtcrn_micro_b200/lib_fbtrace.so tools/gpu_fb_trace.sh
cd "$(dirname "$0")/.."
GTCRNB_FB_TRACE_PRINT=1 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000, device='cuda').mul_(0.1)
m.enhance(x); torch.cuda.synchronize()
PY
