#!/bin/bash
# end-of-round artefacts: GPU parity suite, smoke, bench (both arms), ncu launch list, ncu summaries (text only).  Usage: tools/gpu_final.sh tag
cd "$(dirname "$0")/.."
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -q -s -m gpu 2>&1 | tail -70 > gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$TAG.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python -c "
import json; d = json.load(open('gpurun_out/bench_$TAG.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_pcm16']['value'], d['roofline']['per_class_ms'], d['roofline']['frac'], d['parity']['max_abs'])
s = d['streaming']; print(s['16384'], s['max_realtime_streams'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py --variant full --no-cpu > gpurun_out/bench_full_$TAG.json 2> gpurun_out/bench_full_$TAG.err; python -c "
import json; d = json.load(open('gpurun_out/bench_full_$TAG.json')); print('full', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-stream --no-cpu --batch 256 > /dev/null 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_$TAG.csv
bash tools/gpu_ncu_text.sh $TAG | tail -3
