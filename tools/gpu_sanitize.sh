#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > gpurun_out/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import gtcrn_oracle as O, gtcrn_micro_b200 as G
W = O.load_weights_npz('tests/golden/weights_dns3.npz')
m = G.GTCRNMicro(); m.load_state_dict(W)
x = O.synth_batch(14, 256 * 70 + 9, seed=1).cuda()          # 14 utterances: ragged tensor-core tiles (12 per CTA, 4 columns per warpgroup)
y = m.enhance(x); s = m.stft(x); z = m.istft(m(s)); torch.cuda.synchronize()
sm = G.StreamGTCRNMicro(); G.convert_to_stream(sm, m)
st = sm.new_state(5, 'cuda:0')                                # 5 streams: one full and one ragged 3-stream tile of the TC decoder
for k in range(6):
    o = sm.step_wav(x[:1, 256 * k:256 * (k + 1)].repeat(5, 1).contiguous(), st)
pm = m.pair_metrics(x[:, :y.shape[1]].contiguous(), y)
yh = m.enhance_host(x.cpu().pin_memory())
pool = G.StreamPool(sm, 7, 'cuda:0', groups=2); hin = x[:7, :256].cpu().pin_memory(); hout = torch.empty(7, 256).pin_memory()
for k in range(3): pool.step_wav_host(hin, hout)
torch.cuda.synchronize(); print('ok', y.shape, float((y - z).abs().max()), bool(torch.equal(yh, y.cpu())), pm.shape, 'tc mask', m.tc_mask())
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python gpurun_out/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitize_$tool.log | tail -3
done
