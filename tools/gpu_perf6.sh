#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py tests/test_gpu_stream.py -q -m gpu 2>&1 | tail -3
for v in 1; do
  GTCRNB_TCN_SMEMW=$v python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -2 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("smemw=$v: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_class_ms"])
PY
done
