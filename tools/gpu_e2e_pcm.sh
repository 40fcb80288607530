#!/bin/bash
# e2e float32 vs PCM16 on the wire, alternating, 1024 x 10 s.  Usage: tools/gpu_e2e_pcm.sh [ENV=VAL ...]
cd "$(dirname "$0")/.."
env "$@" python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000).mul_(0.1).pin_memory(); y = torch.empty(1024, 160000).pin_memory()
x16 = (x * 32767).to(torch.int16).pin_memory(); y16 = torch.empty(1024, 160000, dtype=torch.int16).pin_memory()
for _ in range(2): m.enhance_host(x, y); m.enhance_host_pcm16(x16, y16)
for rep in range(3):
    for name, fn in (("f32", lambda: m.enhance_host(x, y)), ("pcm16", lambda: m.enhance_host_pcm16(x16, y16))):
        t0 = time.perf_counter()
        for _ in range(4): fn()
        ms = (time.perf_counter() - t0) / 4 * 1e3
        print(f"{name}: {ms:.3f} ms -> {10240 / ms:.1f} k audio-s/s", flush=True)
PY
