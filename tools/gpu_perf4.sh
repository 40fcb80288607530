#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py -q -x -m gpu -k "host or layer" 2>&1 | tail -2
for c in 64 128 256; do
  GTCRNB_HOST_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -2 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("host chunk $c: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
python - <<'PY'
import torch, time
x = torch.empty(1024, 160000, pin_memory=True); d = torch.empty(1024, 160000, device="cuda")
for name, f in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, f"{x.numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms for 655 MB)")
PY
