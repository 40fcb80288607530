#!/bin/bash
# first GPU validation: offline parity tests, a memcheck pass on a tiny case, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_offline.py -q -s -m gpu 2>&1 | tail -80 > gpurun_out/pytest_offline.log
tail -40 gpurun_out/pytest_offline.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/memcheck.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, 'oracle')
import gtcrn_oracle as O, gtcrn_micro_b200 as G
m = G.GTCRNMicro(); m.load_state_dict(O.load_weights_npz('tests/golden/weights_dns3.npz'))
x = O.synth_batch(3, 256 * 53 + 9, seed=1).cuda()
y = m.enhance(x); torch.cuda.synchronize(); print('ok', y.shape)
PY
echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-stream > gpurun_out/bench_r1_first.json 2> gpurun_out/bench_r1_first.err
echo "bench rc=$?"; cat gpurun_out/bench_r1_first.json; tail -5 gpurun_out/bench_r1_first.err
