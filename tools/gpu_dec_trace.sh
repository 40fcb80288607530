#!/bin/bash
# per-phase cycle counts of one warpgroup step of k_gtconv_dec_tc (clock64 stamps of two threads per warpgroup of CTA (0,0)), 1024 x 10 s.
# Needs a build with -DGTCRNB_DEC_TRACE (GTCRNB_LIB=...).  Usage: GTCRNB_LIB=gtcrn_micro_b200/lib_dectrace.so tools/gpu_dec_trace.sh
cd "$(dirname "$0")/.."
GTCRNB_DEC_TRACE_PRINT=1 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000, device='cuda').mul_(0.1)
for _ in range(2): m.enhance(x)
torch.cuda.synchronize()
PY
