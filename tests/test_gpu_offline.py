"""Parity of the CUDA offline path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): enhanced waveform max-abs <= 1e-4 in fp32, SI-SNR delta <= 0.01 dB.
Spectra / activations are compared relative to their max magnitude (<= 1e-4 * max(1, |ref|_max)).
"""
import numpy as np
import pytest
import torch

import gtcrn_micro_b200 as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
WAV_TOL = 1e-4
REL_TOL = 1e-4


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def make_model(W, b200_lib):
    m = G.GTCRNMicro()
    m.load_state_dict(W)
    return m


def test_stft_istft_vs_oracle(oracle, weights_dns3, b200_lib):
    m = make_model(weights_dns3, b200_lib)
    g = torch.Generator().manual_seed(11)
    for L in (257, 300, 256 * 20 + 77, 16000):
        x = 0.1 * torch.randn(3, L, generator=g)
        s_gpu = m.stft(x.to(DEV))
        s_ref = oracle.stft(x)
        assert tuple(s_gpu.shape) == tuple(s_ref.shape)
        assert rel_err(s_gpu, s_ref) <= 2e-6, L
        w_gpu = m.istft(s_ref.to(DEV))
        w_ref = oracle.istft(s_ref)
        assert tuple(w_gpu.shape) == tuple(w_ref.shape)
        assert float((w_gpu.cpu() - w_ref).abs().max()) <= 2e-6, L


def test_forward_spec_layer_by_layer(oracle, weights, ref_outputs, b200_lib):
    tag, W = weights
    m = make_model(W, b200_lib)
    spec = torch.from_numpy(ref_outputs["spec"])                     # [2,257,41,2]
    taps = {}
    with torch.inference_mode():
        y_or = oracle.forward_spec(W, spec, taps=taps)
    y = m(spec.to(DEV))
    torch.cuda.synchronize()
    B, _, T, _ = spec.shape
    report = []
    names = {"en0": "en0", "en1": "en1", "en2": "en2", "en3": "en3", "en4": "en4", "gtcn1": "gtcn1.3", "gtcn2": "gtcn2.3",
             "de0": "de0", "de1": "de1", "de2": "de2"}
    for mine, theirs in names.items():
        a = m.debug_activation(mine, B, T)                           # [B,T,F,16]
        o = taps[theirs].permute(0, 2, 3, 1)
        report.append((mine, rel_err(a, o)))
    report.append(("out", rel_err(y, y_or)))
    print(f"[{tag}] " + "  ".join(f"{n}:{e:.1e}" for n, e in report))
    for n, e in report:
        assert e <= REL_TOL, (tag, n, e, report)
    # and against the reference's own output (golden fixture)
    assert rel_err(y, torch.from_numpy(ref_outputs[f"{tag}.spec_out"])) <= REL_TOL
    assert m.launch_count() == 10


def test_enhance_golden_pairs(weights_dns3, pairs, b200_lib):
    m = make_model(weights_dns3, b200_lib)
    x = torch.stack([torch.from_numpy(pairs[f"noisy{i}"].astype(np.float32) / 32768.0) for i in range(1, 6)])
    out = m.enhance(x.to(DEV)).cpu().numpy()
    for i in range(5):
        enh = pairs[f"enh{i + 1}"]
        n = len(enh)
        err = np.abs(out[i, :n] - enh.astype(np.float32) / 32768.0).max()
        exact = np.mean(np.floor(out[i, :n] * 32768.0) == enh)
        print(f"pair {i + 1}: max err {err:.3e}  floor-exact {100 * exact:.3f}%")
        assert err <= 3.3e-5                                          # 1 PCM16 LSB (3.05e-5) + fp32 noise
        assert exact > 0.995


@pytest.mark.parametrize("kind", ["speech", "randn"])
def test_enhance_waveform_tolerance(oracle, weights, b200_lib, kind):
    tag, W = weights
    m = make_model(W, b200_lib)
    if kind == "speech":
        x, clean = oracle.synth_batch(4, 32000 + 100, seed=1234, return_clean=True)
    else:
        x = 0.1 * torch.randn(4, 32000 + 100, generator=torch.Generator().manual_seed(0)); clean = None
    ref = oracle.enhance(W, x)
    out = m.enhance(x.to(DEV)).cpu()
    assert out.shape == ref.shape
    err = float((out - ref).abs().max())
    print(f"[{tag}/{kind}] waveform max-abs err {err:.3e} (|ref|max {float(ref.abs().max()):.3f})")
    if tag == "dns3":
        assert err <= WAV_TOL
    else:                                                             # random weights: larger dynamic range
        assert err <= WAV_TOL * max(1.0, float(ref.abs().max()))
    if clean is not None:
        for i in range(4):
            c = clean[i, :ref.shape[1]].numpy()
            d = abs(oracle.si_snr(c, out[i].numpy()) - oracle.si_snr(c, ref[i].numpy()))
            assert d <= 0.01, d


def test_closer_to_fp64_truth_than_tolerance(oracle, weights_dns3, b200_lib):
    # the fp32 reference itself is ~1e-7..1e-6 away from exact arithmetic; so must we be
    m = make_model(weights_dns3, b200_lib)
    x = oracle.synth_batch(2, 16000 + 37, seed=99)
    truth = oracle.enhance(oracle.cast_weights(weights_dns3, torch.float64), x.double())
    out = m.enhance(x.to(DEV)).cpu().double()
    ref32 = oracle.enhance(weights_dns3, x).double()
    e_gpu, e_ref = float((out - truth).abs().max()), float((ref32 - truth).abs().max())
    print(f"vs fp64 truth: gpu {e_gpu:.2e}  torch-fp32 {e_ref:.2e}")
    assert e_gpu <= 2e-5


def test_causality_bitwise_random_init(b200_lib):
    # reference tests/models/test_gtcrn_micro.py:6-39: random-init weights, two 2 s signals sharing the first second
    torch.manual_seed(123)
    m = G.GTCRNMicro()
    g = torch.Generator().manual_seed(4)
    a, b, c = (torch.randn(1, 16000, generator=g) for _ in range(3))
    y1 = m.enhance(torch.cat([a, b], 1).to(DEV))
    y2 = m.enhance(torch.cat([a, c], 1).to(DEV))
    assert float((y1[:, :16000 - 512] - y2[:, :16000 - 512]).abs().max()) == 0.0
    assert float((y1[:, 16000:] - y2[:, 16000:]).abs().max()) > 0.0
    # same through the module-level API exactly as the reference test does it (stft -> model -> istft)
    s1, s2 = m.stft(torch.cat([a, b], 1).to(DEV)), m.stft(torch.cat([a, c], 1).to(DEV))
    z1, z2 = m.istft(m(s1)), m.istft(m(s2))
    assert float((z1[:, :16000 - 512] - z2[:, :16000 - 512]).abs().max()) == 0.0
    # the fused enhance() overlap-adds (a*ie + b*ie) instead of (a + b)*ie: same to rounding
    assert float((z1 - y1).abs().max()) <= 1e-6 * max(1.0, float(y1.abs().max()))


def test_segmentation_and_batch_invariance_bitwise(oracle, weights_dns3, b200_lib):
    # B = 1 splits the utterance into many time segments (with warm-up); B = 300 uses one segment per CTA.
    # The carried-state walk must give bit-identical results either way, and per-utterance results must not
    # depend on what else is in the batch.
    m = make_model(weights_dns3, b200_lib)
    x = oracle.synth_batch(3, 256 * 400 + 13, seed=5).to(DEV)
    alone = torch.cat([m.enhance(x[i:i + 1]) for i in range(3)])
    batched = m.enhance(x.repeat(100, 1))
    assert torch.equal(batched[:3], alone)
    assert torch.equal(batched[297:], alone)
    ref = oracle.enhance(weights_dns3, x[:1].cpu())
    assert float((alone[:1].cpu() - ref).abs().max()) <= WAV_TOL


@pytest.mark.parametrize("T", [1, 2, 3, 15, 16, 17, 47, 48, 49, 97])
def test_ragged_lengths(oracle, weights_dns3, b200_lib, T):
    m = make_model(weights_dns3, b200_lib)
    g = torch.Generator().manual_seed(T)
    spec = torch.randn(3, 257, T, 2, generator=g)
    with torch.inference_mode():
        ref = oracle.forward_spec(weights_dns3, spec)
    out = m(spec.to(DEV))
    assert rel_err(out, ref) <= REL_TOL


def test_strided_inputs_and_empty(oracle, weights_dns3, b200_lib):
    m = make_model(weights_dns3, b200_lib)
    g = torch.Generator().manual_seed(8)
    big = torch.randn(4, 300, 30, 4, generator=g).to(DEV)
    view = big[1:3, 20:277, 3:25, 1:3]                                # non-contiguous in every dim
    with torch.inference_mode():
        ref = oracle.forward_spec(weights_dns3, view.cpu().contiguous())
    assert rel_err(m(view), ref) <= REL_TOL
    # torch.stft-style layout ([B][T][F][2] memory, permuted view)
    s = torch.randn(2, 22, 257, 2, generator=g).to(DEV).permute(0, 2, 1, 3)
    with torch.inference_mode():
        ref = oracle.forward_spec(weights_dns3, s.cpu())
    assert rel_err(m(s), ref) <= REL_TOL
    assert m(torch.zeros(0, 257, 5, 2, device=DEV)).shape == (0, 257, 5, 2)
    assert m.enhance(torch.zeros(0, 1000, device=DEV)).shape == (0, 768)
    with pytest.raises(G.GtcrnB200Error):
        m.enhance(torch.zeros(2, 256, device=DEV))                    # reflect padding needs L > 256 (torch.stft raises too)


@pytest.mark.parametrize("knobs", [{}, {"CHUNK": "1", "RAMP": "0", "COPY": "1"}, {"CHUNK": "2", "G1": "1", "GLAST": "1", "COPY": "1"},
                                   {"CHUNK": "4", "G1": "1", "GLAST": "2", "COPY": "2", "GROWTH": "150", "DECAY": "150"},
                                   {"CHUNK": "8", "G1": "2", "GLAST": "1", "COPY": "1", "PRIO": "0"}])
def test_enhance_host_matches_device(oracle, weights_dns3, b200_lib, monkeypatch, knobs):
    # the pipelined host path (utterance groups rotating over three workspaces, front / back kernels per copy chunk on their own
    # streams) must equal one device call bit for bit, whatever the schedule.  Knobs are in 10 s units: 1 -> 3 utterances of 3 s.
    for k, v in knobs.items():
        monkeypatch.setenv("GTCRNB_HOST_" + k, v)
    m = make_model(weights_dns3, b200_lib)
    x = oracle.synth_batch(70, 16000 * 3 + 5, seed=21)
    xp = x.pin_memory()
    out_h = m.enhance_host(xp)
    out_d = m.enhance(x.to(DEV)).cpu()
    assert torch.equal(out_h, out_d)
    assert torch.equal(m.enhance_host(xp[:1]), out_d[:1])              # a single utterance: one group, one copy chunk
    assert torch.equal(m.enhance_host(xp), out_d)                      # and again on the warm pipeline (event / buffer reuse)


def test_enhance_host_unfused_fft(oracle, weights_dns3, b200_lib, monkeypatch):
    # GTCRNB_FUSE_FFT=0: groups go through the separate STFT / iSTFT kernels as single copy chunks
    monkeypatch.setenv("GTCRNB_FUSE_FFT", "0")
    monkeypatch.setenv("GTCRNB_HOST_CHUNK", "2")
    monkeypatch.setenv("GTCRNB_HOST_G1", "1")
    m = make_model(weights_dns3, b200_lib)
    x = oracle.synth_batch(23, 16000 * 3 + 5, seed=22)
    assert torch.equal(m.enhance_host(x.pin_memory()), m.enhance(x.to(DEV)).cpu())


def test_full_size_properties(oracle, weights_dns3, b200_lib):
    # BASELINE config-2 shape at reduced batch (64 x 10 s): properties that do not need the slow oracle
    m = make_model(weights_dns3, b200_lib)
    x = oracle.synth_batch(64, 160000, seed=77).to(DEV)
    y = m.enhance(x)
    assert y.shape == (64, 160000) and bool(torch.isfinite(y).all())
    # (1) each utterance independent of its batch neighbours (bitwise)
    assert torch.equal(m.enhance(x[5:6]), y[5:6])
    # (2) STFT -> iSTFT round trip reconstructs the input (COLA of the sqrt-Hann pair) at full size
    rt = m.istft(m.stft(x))
    assert float((rt - x[:, :rt.shape[1]]).abs().max()) <= 1e-5
    # (3) mask is bounded (tanh): |Y| <= sqrt(2) * |X| bin by bin
    s = m.stft(x[:4]); o = m(s)
    assert bool((o.pow(2).sum(-1) <= 2.0 * s.pow(2).sum(-1) + 1e-6).all())
    # (4) oracle spot check on one utterance
    ref = oracle.enhance(weights_dns3, x[7:8].cpu())
    assert float((y[7:8].cpu() - ref).abs().max()) <= WAV_TOL


@pytest.mark.parametrize("chunk", [None, "1"])
def test_enhance_host_pcm16_matches_float_path(oracle, weights_dns3, pairs, b200_lib, monkeypatch, chunk):
    if chunk is not None:
        monkeypatch.setenv("GTCRNB_HOST_COPY", "1")
        monkeypatch.setenv("GTCRNB_HOST_G1", "1")
    # PCM16 on the wire: decode x/32768 and floor(y*32768) encode on the device must equal doing both on the host around the
    # float path, bit for bit; and it reproduces the reference's shipped enhanced example
    if chunk is not None:
        monkeypatch.setenv("GTCRNB_HOST_CHUNK", chunk)
    m = make_model(weights_dns3, b200_lib)
    g = torch.Generator().manual_seed(3)
    for B, L in ((5, 16000 * 2 + 3), (37, 256 * 90 + 201), (2, 257)):
        x16 = torch.randint(-20000, 20000, (B, L), generator=g, dtype=torch.int16)
        x16[0, :10] = torch.tensor([-32768, 32767, 0, 1, -1, 12345, -12345, 300, -300, 7], dtype=torch.int16)
        got = m.enhance_host_pcm16(x16.pin_memory())
        y = m.enhance((x16.float() / 32768.0).to(DEV)).cpu()
        want = torch.floor(y * 32768.0).clamp_(-32768, 32767).to(torch.int16)
        assert got.dtype == torch.int16 and got.shape == want.shape
        assert torch.equal(got, want), (B, L)
    n = torch.from_numpy(pairs["noisy1"].copy())[None]
    out = m.enhance_host_pcm16(n)[0].numpy()
    enh = pairs["enh1"]
    assert np.abs(out[:len(enh)].astype(np.int32) - enh).max() <= 1 and np.mean(out[:len(enh)] == enh) > 0.995
    assert m.enhance_host_pcm16(torch.zeros(0, 1000, dtype=torch.int16)).shape == (0, 768)
    with pytest.raises(ValueError):
        m.enhance_host_pcm16(torch.zeros(2, 1000))


def test_two_handles_and_streams_on_one_device(oracle, weights, weights_dns3, b200_lib):
    # a handle owns its workspace: two handles driven from two CUDA streams at the same time must not disturb each other
    tag, W = weights
    ma, mb = make_model(W, b200_lib), make_model(weights_dns3, b200_lib)
    xa = oracle.synth_batch(24, 16000 * 2 + 9, seed=1).to(DEV)
    xb = oracle.synth_batch(17, 16000 * 3 + 100, seed=2).to(DEV)
    ra, rb = ma.enhance(xa), mb.enhance(xb)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(4):
        with torch.cuda.stream(sa):
            ya = ma.enhance(xa)
        with torch.cuda.stream(sb):
            yb = mb.enhance(xb)
    torch.cuda.synchronize()
    assert torch.equal(ya, ra) and torch.equal(yb, rb)


def test_device_call_then_host_pipeline_is_ordered(oracle, weights_dns3, b200_lib):
    # enhance() is asynchronous on the caller's stream and shares the handle's workspace with the host pipeline's internal
    # streams: the pipeline must wait for it (silent corruption without the ordering event)
    m = make_model(weights_dns3, b200_lib)
    big = oracle.synth_batch(96, 160000, seed=5).to(DEV)
    small = oracle.synth_batch(6, 16000, seed=6)
    want_big, want_small = m.enhance(big), m.enhance(small.to(DEV)).cpu()
    torch.cuda.synchronize()
    sp = small.pin_memory()
    for _ in range(3):
        yb = m.enhance(big)                       # ~2 ms of queued kernels on the current stream
        ys = m.enhance_host(sp)                   # must not start on the workspace before they finish
        torch.cuda.synchronize()
        assert torch.equal(ys, want_small) and torch.equal(yb, want_big)


def test_bench_batch_parity_1024x10s(oracle, weights_dns3, b200_lib):
    # the configuration the headline number is quoted on (BASELINE configs[1]: 1024 x 10 s): a fixed 16-utterance subset of
    # the batch is compared with the oracle, waveform max-abs <= 1e-4 and SI-SNR delta <= 0.01 dB
    m = make_model(weights_dns3, b200_lib)
    B, L = 1024, 160000
    idx = [0, 1, 63, 64, 127, 128, 255, 256, 511, 512, 700, 767, 768, 900, 1022, 1023]
    noisy, clean = oracle.synth_batch(len(idx), L, seed=2024, return_clean=True)
    x = torch.empty(B, L, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(9)
    for b0 in range(0, B, 128):
        x[b0:b0 + 128] = 0.05 * torch.randn(128, L, device=DEV, generator=g)
    x[idx] = noisy.to(DEV)
    y = m.enhance(x)
    got = y[idx].cpu()
    ref = oracle.enhance(weights_dns3, noisy)
    err = float((got - ref).abs().max())
    d = max(abs(oracle.si_snr(clean[i, :got.shape[1]].numpy(), got[i].numpy()) - oracle.si_snr(clean[i, :got.shape[1]].numpy(), ref[i].numpy())) for i in range(len(idx)))
    print(f"1024 x 10 s batch, 16-utterance subset: max-abs {err:.2e}, SI-SNR delta {d:.2e} dB")
    assert err <= WAV_TOL and d <= 0.01
    assert bool(torch.isfinite(y).all())
