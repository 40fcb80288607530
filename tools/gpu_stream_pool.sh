#!/bin/bash
# wav-in / wav-out hop time through pinned host buffers against the number of stream groups of StreamPool.  Usage: tools/gpu_stream_pool.sh N [groups ...]
cd "$(dirname "$0")/.."
python - "$@" <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench, gtcrn_micro_b200 as G
m = G.GTCRNMicro(); s = G.StreamGTCRNMicro(); G.convert_to_stream(s, m)
N = int(sys.argv[1])
for g in [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]:
    ms, ok = bench._stream_wav_hops(s, N, torch.device('cuda:0'), 100, groups=g)
    st = bench._hop_stats(ms)
    print(f"N {N} groups {g}: {st}  ok {ok}", flush=True)
    torch.cuda.empty_cache()
PY
