#!/bin/bash
# new GPU tests (file pipeline, metrics, host chunk schedules) + e2e sweep over the host pipeline's chunk schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_infer_pipeline.py tests/test_gpu_offline.py -q -x -m gpu 2>&1 | tail -5
for cfg in "128 0" "128 1" "128 2" "128 3" "192 2" "256 2" "256 3" "96 2"; do
  set -- $cfg
  GTCRNB_HOST_CHUNK=$1 GTCRNB_HOST_RAMP=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("chunk $1 ramp $2: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
