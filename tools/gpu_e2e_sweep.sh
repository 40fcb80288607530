#!/bin/bash
# e2e (enhance_host) wall time for chunk / ramp settings, 1024 x 10 s.  Usage: tools/gpu_e2e_sweep.sh "192 3" "256 3" ...
cd "$(dirname "$0")/.."
python - "$@" <<'PY'
import os, sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000).mul_(0.1).pin_memory(); y = torch.empty(1024, 160000).pin_memory()
for cfg in sys.argv[1:]:
    c, r = cfg.split()
    os.environ["GTCRNB_HOST_CHUNK"] = c; os.environ["GTCRNB_HOST_RAMP"] = r
    for _ in range(2): m.enhance_host(x, y)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(4): m.enhance_host(x, y)
        best = min(best, (time.perf_counter() - t0) / 4 * 1e3)
    print(f"chunk {c} ramp {r}: {best:.3f} ms  -> {10240 / best:.1f} k audio-s/s")
PY
