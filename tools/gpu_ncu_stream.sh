#!/bin/bash
# ncu --set full on the kernels of ONE streaming hop at N concurrent streams.  Usage: tools/gpu_ncu_stream.sh tag [N] [wav|spec]
cd "$(dirname "$0")/.."
TAG=$1; N=${2:-16384}; KIND=${3:-spec}
mkdir -p gpurun_out /tmp/ncu
cat > gpurun_out/stream_hops.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
from gtcrn_micro_b200 import pack
import numpy as np
N, kind, hops = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
z = np.load('tests/golden/weights_dns3.npz')
m = G.GTCRNMicro(); m.load_state_dict({k: torch.from_numpy(np.array(z[k])) for k in z.files})
s = G.StreamGTCRNMicro(); G.convert_to_stream(s, m)
st = s.new_state(N, 'cuda:0')
g = torch.Generator(device='cuda:0').manual_seed(1)
if kind == 'wav':
    xs = [torch.randn(N, 256, device='cuda:0', generator=g) * 0.1 for _ in range(4)]; out = torch.empty(N, 256, device='cuda:0')
    for i in range(hops): s.step_wav(xs[i % 4], st, out)
else:
    xs = [torch.randn(N, 257, 2, device='cuda:0', generator=g) for _ in range(4)]; out = torch.empty(N, 257, 2, device='cuda:0')
    for i in range(hops): s.step(xs[i % 4], st, out)
torch.cuda.synchronize()
PY
# hop = 10 launches (spec) / 12 (wav): skip 6 warm hops, capture one whole hop
PER=10; [ "$KIND" = wav ] && PER=12
GTCRNB_STREAM_GRAPH=0 timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_front|ks_|k_back|k_gtconv_dec_tc" -s $((6 * PER)) -c $PER -f \
    -o /tmp/ncu/prof_$TAG python gpurun_out/stream_hops.py $N $KIND 8 > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
