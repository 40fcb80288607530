#!/bin/bash
# Multi-GPU evidence on one box: bare aggregate DMA bandwidth, weak-scaling bench and the configs[4] sweep at N GPUs.
# Usage (through gpurun --gpus N): tools/gpu_multi.sh N tag
cd "$(dirname "$0")/.."
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
timeout 300 python tools/microbench/dma_aggregate.py $N 512 20 > gpurun_out/dma_aggregate_${TAG}_${N}gpu.txt 2>&1; cat gpurun_out/dma_aggregate_${TAG}_${N}gpu.txt | grep aggregate
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-stream \
    > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_${N}gpu.json"))
print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "pcm16", d["e2e_pcm16"]["value"], "gather", d.get("host_gather"), "parity", d.get("parity", {}).get("ok"))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --warmup 3 --no-stream --no-cpu --sweep-utterances 65536 \
    > gpurun_out/bench_${TAG}_sweep65536_${N}gpu.json 2> gpurun_out/bench_${TAG}_sweep65536_${N}gpu.err; echo "sweep rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_sweep65536_${N}gpu.json"))
print("sweep N=$N value", d["value"], "steps", d["steps"], "e2e", d["e2e"]["value"], "gather", d.get("host_gather"))
PY
