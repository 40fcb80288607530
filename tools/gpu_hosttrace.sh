#!/bin/bash
# per-chunk engine timeline of gtcrnb_enhance_host (GTCRNB_HOST_TRACE=1) for 1024 x 10 s.  Usage: tools/gpu_hosttrace.sh [ENV=VAL ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
env "$@" python - <<'PY' 2> gpurun_out/host_trace.txt
import os, sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000).mul_(0.1).pin_memory(); y = torch.empty(1024, 160000).pin_memory()
for _ in range(2): m.enhance_host(x, y)
t0 = time.perf_counter()
for _ in range(3): m.enhance_host(x, y)
print("wall ms per call", (time.perf_counter() - t0) / 3 * 1e3, file=sys.stderr)
os.environ["GTCRNB_HOST_TRACE"] = "1"
m.enhance_host(x, y)
PY
cat gpurun_out/host_trace.txt
