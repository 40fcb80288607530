#!/bin/bash
# one ncu --set full pass over every kernel of one enhance() call.  Usage: tools/gpu_ncu.sh tag [skip] [count] [regex]
cd "$(dirname "$0")/.."
TAG=${1:-run}; SKIP=${2:-39}; CNT=${3:-13}; RE=${4:-"k_stft|k_front|k_gtconv|k_gtcn|k_back|k_istft"}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu --batch 256 > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep
