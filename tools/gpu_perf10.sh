#!/bin/bash
# FFMA2 A/B: parity tests on the FFMA2 build, then bench both builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for lib in "" "gtcrn_micro_b200/libgtcrn_b200_noffma2.so"; do
  GTCRNB_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("lib '$lib': value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_class_ms"])
PY
done
