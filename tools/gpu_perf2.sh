#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py tests/test_gpu_stream.py -q -x -m gpu 2>&1 | tail -3
for cfg in "2 4"; do
  set -- $cfg
  GTCRNB_GT_PS_ENC=$1 GTCRNB_GT_PS_DEC=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -2 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("enc PS=$1 dec PS=$2: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_class_ms"], "dec TF", d["roofline"]["achieved"])
PY
done
