#!/bin/bash
# e2e (enhance_host) wall time for host-pipeline schedules, 1024 x 10 s.
# Usage: tools/gpu_e2e_sweep.sh "CHUNK=384 G1=48 GLAST=24 COPY=24 GROWTH=200 DECAY=160" "CHUNK=256" ...   (GTCRNB_HOST_ prefix implied)
cd "$(dirname "$0")/.."
python - "$@" <<'PY'
import os, sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
x = torch.randn(1024, 160000).mul_(0.1).pin_memory(); y = torch.empty(1024, 160000).pin_memory()
for cfg in sys.argv[1:]:
    for k in [k for k in os.environ if k.startswith("GTCRNB_HOST_")]: del os.environ[k]
    for kv in cfg.split():
        k, v = kv.split("="); os.environ["GTCRNB_HOST_" + k] = v
    m = G.GTCRN() if os.environ.get('VARIANT') == 'full' else G.GTCRNMicro()    # stream priorities are fixed when the pipeline is first used
    for _ in range(2): m.enhance_host(x, y)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(4): m.enhance_host(x, y)
        best = min(best, (time.perf_counter() - t0) / 4 * 1e3)
    print(f"{cfg or 'defaults'}: {best:.3f} ms  -> {10240 / best:.1f} k audio-s/s", flush=True)
    del m
PY
