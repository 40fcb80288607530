#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "2 4" "2 14" "2 13" "12 14" "2 12"; do
  set -- $cfg
  GTCRNB_GT_PS_ENC=$1 GTCRNB_GT_PS_DEC=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -2 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("enc=$1 dec=$2: value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["per_class_ms"])
PY
done
