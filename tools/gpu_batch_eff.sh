#!/bin/bash
# device-resident gtcrnb_enhance time per utterance against the batch size (10 s utterances): the efficiency a chunked
# pipeline gives up.  Usage: tools/gpu_batch_eff.sh [ENV=VAL ...]
cd "$(dirname "$0")/.."
env "$@" python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.GTCRNMicro()
x = torch.randn(1024, 160000, device='cuda').mul_(0.1)
for B in (6, 12, 24, 48, 96, 128, 192, 256, 384, 512, 1024):
    xb = x[:B]
    for _ in range(3): m.enhance(xb)
    torch.cuda.synchronize()
    n = max(3, 1024 // B // 2)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m.enhance(xb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B {B:5d}: {ms:8.3f} ms  {ms / B * 1e3:7.2f} us/utt  ({ms / B * 1024:6.2f} ms per 1024)")
PY
