#!/bin/bash
# FFMA2 follow-up: decoder 3x3 feed variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env $1 python bench.py --steps 5 --warmup 3 --no-cpu --no-stream > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; tail -1 gpurun_out/bench_cfg.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg.json"))
print("$1: value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["per_class_ms"])
PY
}
run "GTCRNB_LIB=gtcrn_micro_b200/libgtcrn_b200_decscalar.so"
run "GTCRNB_GT_PS_DEC=14"
run "GTCRNB_GT_PS_DEC=13"
run "GTCRNB_GT_PS_DEC=3"
run "GTCRNB_GT_PS_DEC=2"
run "GTCRNB_GT_PS_ENC=3"
run "GTCRNB_GT_PS_ENC=12"
run "GTCRNB_TCN_SMEMW=0"
