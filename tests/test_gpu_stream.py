"""Parity of the CUDA streaming path (frame by frame, ring-buffer state) through the C ABI.

Anchors: the reference's whole-model stream-vs-offline comparison (gtcrn_micro_stream.py:577-657, printed there,
asserted here), its cache layout (:618-623) and the golden reference streaming outputs in tests/golden.
"""
import numpy as np
import pytest
import torch

import gtcrn_micro_b200 as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_TOL = 1e-4
WAV_TOL = 1e-4


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def models(W):
    m = G.GTCRNMicro(); m.load_state_dict(W)
    s = G.StreamGTCRNMicro(); G.convert_to_stream(s, m)
    return m, s


def run_stream(smodel, spec):
    """spec [B,257,T,2] on DEV -> [B,257,T,2] via the native state API"""
    B, _, T, _ = spec.shape
    st = smodel.new_state(B, DEV)
    frames = spec.permute(2, 0, 1, 3).contiguous()
    outs = [smodel.step(frames[t], st) for t in range(T)]
    return torch.stack(outs, dim=2)


def test_stream_equals_offline_and_reference(oracle, weights, ref_outputs, b200_lib):
    tag, W = weights
    m, s = models(W)
    spec = torch.from_numpy(ref_outputs["spec"]).to(DEV)                    # [2,257,41,2]
    ys = run_stream(s, spec)
    y_off = m(spec)
    bitwise = bool(torch.equal(ys, y_off.contiguous()))
    e_off = rel_err(ys, y_off)
    e_ref = rel_err(ys, torch.from_numpy(ref_outputs[f"{tag}.stream_spec_out"]))
    with torch.inference_mode():
        e_or = rel_err(ys, oracle.forward_spec(W, spec.cpu()))
    print(f"[{tag}] stream vs own offline {e_off:.2e} (bitwise {bitwise}); vs reference stream {e_ref:.2e}; vs oracle offline {e_or:.2e}")
    assert e_off <= 1e-5 and e_ref <= REL_TOL and e_or <= REL_TOL
    assert s.launch_count() == 41 * 8                                      # front, fused encoder stack, GTCNs, 3 decoder blocks, back, advance


def test_reference_cache_api(oracle, weights, ref_outputs, b200_lib):
    # drive StreamGTCRNMicro.forward exactly like the reference loop (gtcrn_micro_stream.py:618-636)
    tag, W = weights
    _, s = models(W)
    spec = torch.from_numpy(ref_outputs["spec"]).to(DEV)
    B = spec.shape[0]
    conv_cache = torch.zeros(2, B, 16, 6, 33, device=DEV)
    tra_cache = torch.zeros(2, 3, B, 8, 2, device=DEV)
    tcn_cache = [[torch.zeros(B, 16, 2 * d, 33, device=DEV) for d in (1, 2, 4, 8)] for _ in range(2)]
    ys = []
    for t in range(25):
        yi, conv_cache, tra_cache, tcn_cache = s(spec[:, :, t:t + 1], conv_cache, tra_cache, tcn_cache)
        assert yi.shape == (B, 257, 1, 2)
        ys.append(yi)
        if t == 20:
            assert rel_err(conv_cache, torch.from_numpy(ref_outputs[f"{tag}.conv_cache20"])) <= REL_TOL
            assert rel_err(tra_cache, torch.from_numpy(ref_outputs[f"{tag}.tra_cache20"])) <= REL_TOL
            for g in range(2):
                for k in range(4):
                    assert rel_err(tcn_cache[g][k], torch.from_numpy(ref_outputs[f"{tag}.tcn_cache20.{g}.{k}"])) <= REL_TOL, (g, k)
    ys = torch.cat(ys, dim=2)
    assert rel_err(ys, torch.from_numpy(ref_outputs[f"{tag}.stream_spec_out"][:, :, :25])) <= REL_TOL
    with pytest.raises(ValueError):
        s(spec[:, :, :2], conv_cache, tra_cache, tcn_cache)                 # T = 1 per call, like the reference
    with pytest.raises(AssertionError):
        s(spec[:, :, :1], conv_cache[:, :, :, :4], tra_cache, tcn_cache)    # cache shape asserts (gtcrn_micro_stream.py:131-138)


def test_stream_wav_equals_offline(oracle, weights_dns3, b200_lib):
    m, s = models(weights_dns3)
    N, hops = 5, 40
    x = oracle.synth_batch(N, 256 * hops, seed=31).to(DEV)
    off = m.enhance(x)                                                        # [N, 256*40]
    st = s.new_state(N, DEV)
    outs = [s.step_wav(x[:, 256 * k:256 * (k + 1)].contiguous(), st) for k in range(hops)]
    y = torch.cat(outs, dim=1)
    assert float(y[:, :512].abs().max()) == 0.0                               # two hops of latency
    # call k emits hop k-2; the last offline hop depends on the end-of-file reflection and is never streamed
    got, want = y[:, 512:], off[:, :256 * (hops - 2)]
    err = float((got - want).abs().max())
    print(f"stream wav vs offline: max-abs {err:.2e}")
    assert err <= 1e-6
    ref = oracle.enhance(weights_dns3, x.cpu())
    assert float((got.cpu() - ref[:, :256 * (hops - 2)]).abs().max()) <= WAV_TOL
    # reset -> identical replay
    s.reset(st)
    y2 = torch.cat([s.step_wav(x[:, 256 * k:256 * (k + 1)].contiguous(), st) for k in range(hops)], dim=1)
    assert torch.equal(y, y2)


def test_many_streams_independent(oracle, weights_dns3, b200_lib):
    _, s = models(weights_dns3)
    N, T = 1003, 20                                                           # not a multiple of any tile size
    g = torch.Generator(device=DEV).manual_seed(5)
    spec = torch.randn(N, 257, T, 2, device=DEV, generator=g)
    ys = run_stream(s, spec)
    assert bool(torch.isfinite(ys).all())
    for i in (0, 7, 500, 1002):
        alone = run_stream(s, spec[i:i + 1])
        assert torch.equal(alone, ys[i:i + 1]), i
    with torch.inference_mode():
        ref = oracle.forward_spec(weights_dns3, spec[1000:1003].cpu())
    assert rel_err(ys[1000:1003], ref) <= REL_TOL


def test_long_stream_ring_wraparound(oracle, weights_dns3, b200_lib):
    # 200 frames: every ring (depth 2..16) wraps many times
    m, s = models(weights_dns3)
    spec = m.stft(oracle.synth_batch(2, 256 * 199 + 3, seed=8).to(DEV))
    ys = run_stream(s, spec)
    assert rel_err(ys, m(spec)) <= 1e-5


def test_hop_graphs_replay_bitwise(oracle, weights_dns3, b200_lib, monkeypatch):
    # With static in/out buffers the library replays a CUDA graph of the hop from the third call on (the 10-12 launches with
    # by-value weight blocks make small hops launch-bound).  Graph replay, plain launches and changing buffers must agree bit for bit,
    # for both step flavours, also after the state was reset and after an offline call grew the shared workspace.
    m, s = models(weights_dns3)
    s.step(torch.zeros(1, 257, 2, device=DEV), s.new_state(1, DEV))    # create the handle (reads the env) while graphs are on
    monkeypatch.setenv("GTCRNB_STREAM_GRAPH", "0")
    m0, s0 = models(weights_dns3)
    s0.step(torch.zeros(1, 257, 2, device=DEV), s0.new_state(1, DEV))
    monkeypatch.delenv("GTCRNB_STREAM_GRAPH")
    N, HOPS = 5, 24
    g = torch.Generator().manual_seed(12)
    wav = (0.1 * torch.randn(HOPS, N, 256, generator=g)).to(DEV)
    spec = torch.randn(HOPS, N, 257, 2, generator=g).to(DEV)
    st_g, st_p, st_v = s.new_state(N, DEV), s0.new_state(N, DEV), s.new_state(N, DEV)
    hop_in, hop_out = torch.empty(N, 256, device=DEV), torch.empty(N, 256, device=DEV)
    l0 = s.launch_count()
    for k in range(HOPS):
        hop_in.copy_(wav[k])
        s.step_wav(hop_in, st_g, hop_out)                              # stable pointers -> graph
        ref = s0.step_wav(wav[k].contiguous(), st_p)                   # graphs disabled
        var = s.step_wav(wav[k].clone(), st_v)                         # graphs enabled, pointers change every hop
        assert torch.equal(hop_out, ref) and torch.equal(var, ref), k
        if k == 11:                                                    # a bigger stream set grows the workspace under the cached graph
            s.step_wav(torch.zeros(64, 256, device=DEV), s.new_state(64, DEV))
    assert s.launch_count() - l0 >= 2 * HOPS * 10                      # replayed hops are counted like launched ones (10 kernels per hop)
    st_g2, st_p2 = s.new_state(N, DEV), s0.new_state(N, DEV)
    f_in, f_out = torch.empty(N, 257, 2, device=DEV), torch.empty(N, 257, 2, device=DEV)
    for k in range(HOPS):
        f_in.copy_(spec[k])
        s.step(f_in, st_g2, f_out)
        assert torch.equal(f_out, s0.step(spec[k].contiguous(), st_p2)), k
        if k == 9:
            s.reset(st_g2); s0.reset(st_p2)


@pytest.mark.parametrize("tail", [0, 1, 77, 255])
def test_stream_flush_completes_offline_output(oracle, weights_dns3, b200_lib, tail):
    # SURVEY 8(f)1 "OLA tail flush": step_wav + flush_wav == enhance of the whole signal (infer.py:60-76), including the
    # end-of-signal reflect padding and trailing samples that do not fill a hop
    m, s = models(weights_dns3)
    N, hops = 3, 11
    x = oracle.synth_batch(N, 256 * hops + tail, seed=32 + tail).to(DEV)
    off = m.enhance(x)                                                        # [N, 256*hops]
    st = s.new_state(N, DEV)
    outs = [s.step_wav(x[:, 256 * k:256 * (k + 1)].contiguous(), st) for k in range(hops)]
    outs.append(s.flush_wav(st, x[:, 256 * hops:].contiguous() if tail else None))
    y = torch.cat(outs, dim=1)[:, 512:]
    assert y.shape == off.shape
    err = float((y - off).abs().max())
    print(f"stream wav + flush (tail {tail}) vs offline: max-abs {err:.2e}")
    assert err <= 1e-6
    assert float((y.cpu() - oracle.enhance(weights_dns3, x.cpu())).abs().max()) <= WAV_TOL
    with pytest.raises(ValueError):
        s.flush_wav(st, torch.zeros(N, 256, device=DEV))


def test_reference_forward_reuses_native_state(oracle, weights_dns3, ref_outputs, b200_lib):
    # the drop-in forward keeps its ring-buffer state between calls when the caller feeds the returned caches back
    # (one import in total), and re-imports when the caches were replaced or modified through torch
    _, s = models(weights_dns3)
    spec = torch.from_numpy(ref_outputs["spec"]).to(DEV)
    B = spec.shape[0]

    def fresh():
        return (torch.zeros(2, B, 16, 6, 33, device=DEV), torch.zeros(2, 3, B, 8, 2, device=DEV),
                [[torch.zeros(B, 16, 2 * d, 33, device=DEV) for d in (1, 2, 4, 8)] for _ in range(2)])
    cc, tc, nc = fresh()
    l0 = s.launch_count()
    ys = []
    for t in range(12):
        yi, cc, tc, nc = s(spec[:, :, t:t + 1], cc, tc, nc)
        ys.append(yi)
    assert s.launch_count() - l0 == 1 + 12 * (8 + 1)              # one import, then step (8 launches) + export per frame
    want = torch.from_numpy(ref_outputs["dns3.stream_spec_out"][:, :, :12])
    assert rel_err(torch.cat(ys, dim=2), want) <= REL_TOL
    # restart from zero caches (new tensors): must re-import and reproduce frame 0
    cc2, tc2, nc2 = fresh()
    y0, cc2, tc2, nc2 = s(spec[:, :, 0:1], cc2, tc2, nc2)
    assert torch.equal(y0, ys[0])
    # in-place edit of a returned cache through torch is honoured
    cc2.zero_(); tc2.zero_()
    for g in range(2):
        for k in range(4):
            nc2[g][k].zero_()
    y0b, *_ = s(spec[:, :, 0:1], cc2, tc2, nc2)
    assert torch.equal(y0b, ys[0])


def test_stream_parity_at_scale_16384(oracle, weights_dns3, b200_lib):
    # BASELINE configs[2] size: 16,384 concurrent streams x 100 hops; a 16-stream subset is compared with the oracle's
    # frame-by-frame loop and every stream with its bitwise twin (streams 0..15 are repeated at the end of the batch)
    _, s = models(weights_dns3)
    N, T, K = 16384, 100, 16
    g = torch.Generator(device=DEV).manual_seed(77)
    st = s.new_state(N, DEV)
    sub_in, sub_out = [], []
    for t in range(T):
        f = torch.randn(N, 257, 2, device=DEV, generator=g)
        f[N - K:] = f[:K]
        o = s.step(f, st)
        assert torch.equal(o[N - K:], o[:K]), t
        sub_in.append(f[:K].cpu()); sub_out.append(o[:K].cpu())
    spec = torch.stack(sub_in, dim=2)                                 # [K,257,T,2]
    with torch.inference_mode():
        ref = oracle.stream_spec(weights_dns3, spec)
    e = rel_err(torch.stack(sub_out, dim=2), ref)
    print(f"16384 streams x {T} hops: subset rel err vs oracle stream loop {e:.2e}")
    assert e <= REL_TOL


@pytest.mark.parametrize("groups", [1, 3])
def test_stream_pool_host_buffers_equal_single_state(oracle, weights_dns3, b200_lib, groups):
    # StreamPool.step_wav_host: pinned host buffers in / out, streams split into groups whose copies overlap the other groups'
    # kernels.  Streams are independent, so every hop must equal one step_wav over all streams bit for bit (graphs kick in on the
    # second hop of every group: rotating keys).
    m, s = models(weights_dns3)
    N, HOPS = 37, 12
    pool = G.StreamPool(s, N, DEV, groups=groups)
    assert sum(pool.sizes) == N and len(pool.sizes) == groups
    st = s.new_state(N, DEV)
    g = torch.Generator().manual_seed(5)
    hin, hout = torch.empty(N, 256).pin_memory(), torch.empty(N, 256).pin_memory()
    for k in range(HOPS):
        hin.copy_(0.1 * torch.randn(N, 256, generator=g))
        pool.step_wav_host(hin, hout)
        want = s.step_wav(hin.to(DEV), st)
        torch.cuda.synchronize()
        assert torch.equal(hout, want.cpu()), k
    with pytest.raises(ValueError):
        pool.step_wav_host(hin[:5], hout)
