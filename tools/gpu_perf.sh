#!/bin/bash
# quick iteration loop: a few parity tests + short bench (no CPU arm).  Usage: tools/gpu_perf.sh tag [extra bench args]
cd "$(dirname "$0")/.."
TAG=${1:-run}; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py tests/test_gpu_stream.py -q -x -m gpu -k "layer_by_layer or bitwise or ragged or stream_equals or wav_equals" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
print("per_class_ms", d["roofline"]["per_class_ms"], "dec TF", d["roofline"]["achieved"])
print("streaming", d.get("streaming"))
PY
