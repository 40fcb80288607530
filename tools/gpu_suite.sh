#!/bin/bash
# GPU validation suite: parity tests (-m gpu), smoke, bench, ncu launch list.  Usage: tools/gpu_suite.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -q -s -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$TAG.log | tail -3
grep -E "^\[|^\.\[|stream wav|pair|fp64|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-stream --no-cpu --batch 256 > /dev/null 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_$TAG.csv
