#!/bin/bash
# host pipeline: per-chunk timeline + e2e sweep (triple staging)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py -q -x -m gpu -k host 2>&1 | tail -2
python - <<'PY' 2> gpurun_out/host_trace.txt
import os, sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
m = G.load_checkpoint("tests/golden/weights_dns3.npz") if False else G.GTCRNMicro()
x = torch.randn(1024, 160000).mul_(0.1).pin_memory(); y = torch.empty(1024, 160000).pin_memory()
for _ in range(2): m.enhance_host(x, y)
os.environ["GTCRNB_HOST_TRACE"] = "1"
m.enhance_host(x, y)
PY
cat gpurun_out/host_trace.txt
for cfg in "128 2" "192 2" "256 2" "256 3" "192 3"; do
  set -- $cfg
  GTCRNB_HOST_CHUNK=$1 GTCRNB_HOST_RAMP=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("chunk $1 ramp $2: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
