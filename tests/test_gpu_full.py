"""BASELINE configs[3]: the original-GTCRN-style variant (unfold SFE, groups = 2, dilated depthwise GTConv, GRU-TRA, DPGRNN) on the
GPU, through the C ABI (gtcrnb_create_variant(..., GTCRNB_VARIANT_FULL)).

PARITY UNPINNED: the reference has no implementation of this variant (only commented-out lines, gtcrn_micro.py:187,199,346,361,
381,391,417-418,429,448,494), so every comparison here is against the builder-written oracle/gtcrn_full_oracle.py (itself
cross-checked against a stock-torch-layer model in tests/test_full_oracle.py) with seeded random weights.  Tolerances as for the
Micro path: waveform max-abs <= 1e-4, SI-SNR delta <= 0.01 dB, activations <= 1e-4 relative."""
import pytest
import torch

import gtcrn_full_oracle as FO
import gtcrn_micro_b200 as G
import gtcrn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


@pytest.fixture(scope="module", params=[3, 11])
def full(request, b200_lib):
    W = FO.random_weights(request.param)
    m = G.GTCRN(); m.load_state_dict(W)
    s = G.StreamGTCRN(); G.convert_to_stream(s, m)
    return W, m, s


def test_unpinned_full_layer_by_layer(full):
    W, m, _ = full
    spec = O.stft(O.synth_batch(2, 256 * 40 + 9, seed=4))
    taps = {}
    with torch.inference_mode():
        want = FO.forward_spec(W, spec, taps=taps)
    got = m(spec.to(DEV))
    B, _, T, _ = spec.shape
    rep = []
    for mine, theirs in (("en0", "en0"), ("en1", "en1"), ("en2", "en2"), ("en3", "en3"), ("en4", "en4"), ("dp1", "dp1"), ("dp2", "dp2"),
                         ("de0", "de0"), ("de1", "de1"), ("de2", "de2")):
        rep.append((mine, rel_err(m.debug_activation(mine, B, T), taps[theirs].permute(0, 2, 3, 1))))
    rep.append(("out", rel_err(got, want)))
    print("full variant, layer by layer:", " ".join(f"{n}={e:.1e}" for n, e in rep))
    assert all(e <= TOL for _, e in rep), rep


@pytest.mark.parametrize("T", [1, 2, 3, 7, 8, 9, 16, 17, 33])
def test_unpinned_full_ragged_lengths(full, T):
    W, m, _ = full
    spec = torch.randn(3, 257, T, 2, generator=torch.Generator().manual_seed(T))
    with torch.inference_mode():
        want = FO.forward_spec(W, spec)
    assert rel_err(m(spec.to(DEV)), want) <= TOL


def test_unpinned_full_enhance_waveform_and_sisnr(full):
    W, m, _ = full
    noisy, clean = O.synth_batch(5, 16000 * 3 + 77, seed=8, return_clean=True)
    want = FO.enhance(W, noisy)
    got = m.enhance(noisy.to(DEV)).cpu()
    assert got.shape == want.shape
    err = float((got - want).abs().max())
    n = got.shape[1]
    d = max(abs(O.si_snr(clean[i, :n].numpy(), got[i].numpy()) - O.si_snr(clean[i, :n].numpy(), want[i].numpy())) for i in range(5))
    print(f"full variant enhance: max-abs {err:.2e}, SI-SNR delta {d:.2e} dB")
    assert err <= TOL and d <= 0.01
    # host pipeline goes through the same kernels
    assert torch.equal(m.enhance_host(noisy.pin_memory()), got)


def test_unpinned_full_host_pipeline_groups(full, monkeypatch):
    # several utterance groups and copy chunks (front / back per chunk, persistent GTConv / DPGRNN kernels per group): bitwise equal
    for k, v in {"CHUNK": "2", "G1": "1", "GLAST": "1", "COPY": "1"}.items():
        monkeypatch.setenv("GTCRNB_HOST_" + k, v)
    W, m, _ = full
    noisy = O.synth_batch(29, 16000 * 3 + 77, seed=9)
    assert torch.equal(m.enhance_host(noisy.pin_memory()), m.enhance(noisy.to(DEV)).cpu())


def test_unpinned_full_batch_invariance_and_causality_bitwise(full):
    W, m, _ = full
    x = O.synth_batch(9, 256 * 50, seed=2).to(DEV)
    y = m.enhance(x)
    assert torch.equal(m.enhance(x[4:5]), y[4:5])
    x2 = x.clone()
    x2[:, 256 * 30:] = 0.1 * torch.randn(9, 256 * 20, device=DEV)
    y2 = m.enhance(x2)
    # frames up to 28 see only unchanged samples (frame t covers samples 256(t-1) .. 256(t+1)); output hop k needs frames k, k+1
    assert torch.equal(y[:, :256 * 28], y2[:, :256 * 28])
    assert not torch.equal(y[:, 256 * 31:], y2[:, 256 * 31:])


def test_unpinned_full_stream_equals_offline(full):
    W, m, s = full
    x = O.synth_batch(4, 256 * 45, seed=6)
    spec = O.stft(x).to(DEV)
    off = m(spec)
    st = s.new_state(4, DEV)
    frames = spec.permute(2, 0, 1, 3).contiguous()
    ys = torch.stack([s.step(frames[t], st) for t in range(frames.shape[0])], dim=2)
    bitwise = bool(torch.equal(ys, off.contiguous()))
    e_off = rel_err(ys, off)
    with torch.inference_mode():
        e_or = rel_err(ys, FO.stream_spec(W, spec.cpu()))
    print(f"full variant stream vs offline {e_off:.2e} (bitwise {bitwise}); vs oracle stream loop {e_or:.2e}")
    assert e_off <= 1e-5 and e_or <= TOL
    # reset -> identical replay; streams independent of their batch neighbours
    s.reset(st)
    ys2 = torch.stack([s.step(frames[t], st) for t in range(frames.shape[0])], dim=2)
    assert torch.equal(ys, ys2)
    st1 = s.new_state(1, DEV)
    y1 = torch.stack([s.step(frames[t][2:3].contiguous(), st1) for t in range(frames.shape[0])], dim=2)
    assert torch.equal(y1, ys[2:3])


@pytest.mark.parametrize("tail", [0, 100])
def test_unpinned_full_stream_wav_and_flush_equal_enhance(full, tail):
    W, m, s = full
    N, hops = 3, 14
    x = O.synth_batch(N, 256 * hops + tail, seed=13).to(DEV)
    off = m.enhance(x)
    st = s.new_state(N, DEV)
    outs = [s.step_wav(x[:, 256 * k:256 * (k + 1)].contiguous(), st) for k in range(hops)]
    outs.append(s.flush_wav(st, x[:, 256 * hops:].contiguous() if tail else None))
    y = torch.cat(outs, dim=1)[:, 512:]
    assert y.shape == off.shape
    assert float((y - off).abs().max()) <= 1e-6
    assert float((y.cpu() - FO.enhance(W, x.cpu())).abs().max()) <= TOL


def test_unpinned_full_variant_handles_are_separate(full, weights_dns3):
    # a Micro handle and a Full handle side by side; wrong-variant blobs are rejected by the library
    import ctypes as C
    from gtcrn_micro_b200 import lib, pack, pack_full
    W, m, _ = full
    mm = G.GTCRNMicro(); mm.load_state_dict(weights_dns3)
    x = O.synth_batch(2, 256 * 20, seed=1).to(DEV)
    a, b = m.enhance(x), mm.enhance(x)
    assert not torch.equal(a, b)
    L = lib.load()
    h = C.c_void_p(None)
    blob = pack.pack_state_dict(weights_dns3)
    assert L.gtcrnb_create_variant(blob.ctypes.data_as(C.c_void_p), blob.size, 1, C.byref(h)) != 0
    blobf = pack_full.pack_state_dict(W)
    assert L.gtcrnb_create_variant(blobf.ctypes.data_as(C.c_void_p), blobf.size, 0, C.byref(h)) != 0
    assert L.gtcrnb_create_variant(blobf.ctypes.data_as(C.c_void_p), blobf.size, 7, C.byref(h)) != 0
