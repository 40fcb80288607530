#!/bin/bash
# ncu --set full on selected kernels at the real bench batch.  Usage: tools/gpu_ncu.sh tag regex skip count [env...]
cd "$(dirname "$0")/.."
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
env "$@" timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu --no-parity --batch 1024 $BENCH_ARGS > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
