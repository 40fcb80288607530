#!/bin/bash
# device time of one streaming hop (spectra resident on the device) at N streams, CUDA events over 200 hops.  Usage: tools/gpu_stream_hop.sh "ENV=.." ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
env $cfg python - "$cfg" <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import gtcrn_micro_b200 as G
z = np.load('tests/golden/weights_dns3.npz')
m = G.GTCRNMicro(); m.load_state_dict({k: torch.from_numpy(np.array(z[k])) for k in z.files})
s = G.StreamGTCRNMicro(); G.convert_to_stream(s, m)
for N in (4096, 16384, 65536):
    st = s.new_state(N, 'cuda:0')
    g = torch.Generator(device='cuda:0').manual_seed(1)
    xs = [torch.randn(N, 257, 2, device='cuda:0', generator=g) for _ in range(4)]; out = torch.empty(N, 257, 2, device='cuda:0')
    for i in range(20): s.step(xs[i % 4], st, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(200): s.step(xs[i % 4], st, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"{sys.argv[1]!r:40s} N={N:6d}: {ms*1e3:8.1f} us per hop  {ms*1e6/N:6.1f} ns per frame")
PY
done
