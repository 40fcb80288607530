#!/bin/bash
# A/B of build variants (GTCRNB_LIB) and env knobs.  Usage: tools/gpu_ab.sh "ENV1=a ENV2=b" "..." ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_cfg.json"))
    print("$cfg: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["per_class_ms"], "parity", d.get("parity", {}).get("max_abs"), d.get("parity", {}).get("si_snr_delta_db"))
except Exception as e:
    print("$cfg: FAILED", e)
PY
done
