#!/bin/bash
# e2e (enhance_host) with N GPUs busy at once, one process per GPU, for host-buffer variants.  Usage: tools/gpu_multi_e2e.sh N mode...
# modes: torch (torch pin_memory in/out), lib (gtcrnb_host_alloc in/out), wc (write-combined input, gtcrnb_host_alloc output)
cd "$(dirname "$0")/.."
N=$1; shift
mkdir -p gpurun_out
cat > gpurun_out/multi_e2e.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
mode = sys.argv[1]
m = G.GTCRNMicro()
B, L = 1024, 160000
src = torch.randn(B, L).mul_(0.1)
if mode == "torch":
    x = torch.empty(B, L, pin_memory=True); y = torch.empty(B, L, pin_memory=True)
elif mode == "lib":
    x = G.pinned_empty((B, L)); y = G.pinned_empty((B, L))
else:
    x = G.pinned_empty((B, L), write_combined=True); y = G.pinned_empty((B, L))
x.copy_(src)
for _ in range(2): m.enhance_host(x, y)
# crude cross-process alignment: start at the next multiple of 4 s of wall clock
t = time.time(); time.sleep(4 - (t % 4))
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(4): m.enhance_host(x, y)
    best = min(best, (time.perf_counter() - t0) / 4 * 1e3)
ref = m.enhance(src[:2].cuda()).cpu()
print(f"gpu {os.environ.get('CUDA_VISIBLE_DEVICES')} mode {mode}: {best:.2f} ms/step -> {10240 / best:.1f} k audio-s/s  exact={bool(torch.equal(ref, y[:2]))}", flush=True)
PY
for mode in "$@"; do
  for ((g = 0; g < N; g++)); do CUDA_VISIBLE_DEVICES=$g python gpurun_out/multi_e2e.py $mode & done
  wait
done
