#!/bin/bash
# A/B of build variants (GTCRNB_LIB) and env knobs.  Usage: tools/gpu_ab.sh "ENV1=a ENV2=b" "..." ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("$cfg: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_class_ms"])
PY
done
