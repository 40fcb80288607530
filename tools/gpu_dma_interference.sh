#!/bin/bash
# Does PCIe traffic slow the kernels down?  Device-resident gtcrnb_enhance (1024 x 10 s) alone and with pinned H2D / D2H copies
# running on other streams for the whole duration.
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000, device='cuda').mul_(0.1)
hin = torch.randn(256, 160000).pin_memory(); hout = torch.empty(256, 160000).pin_memory()
din = torch.empty(256, 160000, device='cuda'); dout = torch.randn(256, 160000, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, pcm=False):
    for _ in range(2): m.enhance(x)
    torch.cuda.synchronize()
    n = 5
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 12 * n
    if h2d:
        with torch.cuda.stream(s1):
            for _ in range(reps): din.copy_(hin, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            for _ in range(reps): hout.copy_(dout, non_blocking=True)
    e0.record()
    for _ in range(n): m.enhance(x)
    e1.record()
    e1.synchronize()
    busy1 = not s1.query(); busy2 = not s2.query()
    torch.cuda.synchronize()
    print(f"h2d={h2d} d2h={d2h}: {e0.elapsed_time(e1) / n:.3f} ms per 1024 x 10 s  (copies still running at the end: {busy1 if h2d else '-'} {busy2 if d2h else '-'})")
run(False, False); run(True, False); run(False, True); run(True, True); run(False, False)
PY
