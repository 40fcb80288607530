#!/bin/bash
# wall-clock per-hop latency of the streaming path (launch + kernels + sync, what a real-time loop sees) for small stream counts
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro(); sm = G.StreamGTCRNMicro(); G.convert_to_stream(sm, m)
for N in (1, 16, 256, 1024, 4096):
    st = sm.new_state(N, 'cuda:0')
    hop = torch.randn(N, 256, device='cuda:0') * 0.1; out = torch.empty_like(hop)
    spec = torch.randn(N, 257, 2, device='cuda:0'); so = torch.empty_like(spec)
    for name, fn in (("step_wav", lambda: sm.step_wav(hop, st, out)), ("step_spec", lambda: sm.step(spec, st, so))):
        for _ in range(20): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
        ts.sort()
        print(f"N={N:5d} {name:9s}: p50 {ts[100]:7.1f} us  p99 {ts[197]:7.1f} us")
PY
