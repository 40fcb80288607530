"""The oracle against the committed golden fixtures (made from the real reference by make_golden.py)."""
import numpy as np
import torch


def test_golden_pairs_within_one_lsb(oracle, weights_dns3, pairs):
    # reference: examples/gtcrn_micro/noisy{i}.wav -> enh{i}.wav through infer.py:60-76 with best_model_dns3.tar
    for i in range(1, 6):
        x = torch.from_numpy(pairs[f"noisy{i}"].astype(np.float32) / 32768.0)[None]
        out = oracle.enhance(weights_dns3, x)[0].numpy()
        enh = pairs[f"enh{i}"]
        n = len(enh)                                     # 256 * 311: frames that never see the cropped tail
        err = np.abs(out[:n] - enh.astype(np.float32) / 32768.0).max()
        assert err <= 3.1e-5, (i, err)                   # 1 PCM16 LSB = 3.05e-5
        assert np.mean(np.floor(out[:n] * 32768.0) == enh) > 0.999


def test_reference_outputs(oracle, weights, ref_outputs):
    tag, W = weights
    wav = torch.from_numpy(ref_outputs["wav"])
    spec = torch.from_numpy(ref_outputs["spec"])
    scale = max(1.0, float(np.abs(ref_outputs[f"{tag}.spec_out"]).max()))
    assert (oracle.stft(wav) - spec).abs().max() <= 2e-5
    with torch.inference_mode():
        y = oracle.forward_spec(W, spec)
        assert np.abs(y.numpy() - ref_outputs[f"{tag}.spec_out"]).max() / scale <= 2e-5
        ys = oracle.stream_spec(W, spec)
        assert np.abs(ys.numpy() - ref_outputs[f"{tag}.stream_spec_out"]).max() / scale <= 2e-5
        assert np.abs(oracle.enhance(W, wav).numpy() - ref_outputs[f"{tag}.wav_out"]).max() <= 2e-6
        # streaming caches in the reference's own tensor layout after 21 frames
        st = oracle.StreamState(wav.shape[0])
        for t in range(21):
            oracle.forward_spec(W, spec[:, :, t:t + 1], st)
        cc, tc, nc = st.to_reference()
        rel = lambda a, b: np.abs(a.numpy() - b).max() / max(1.0, np.abs(b).max())
        assert rel(cc, ref_outputs[f"{tag}.conv_cache20"]) <= 2e-5
        assert rel(tc, ref_outputs[f"{tag}.tra_cache20"]) <= 2e-5
        for g in range(2):
            for k in range(4):
                assert rel(nc[g][k], ref_outputs[f"{tag}.tcn_cache20.{g}.{k}"]) <= 2e-5


def test_oracle_causality_bitwise(oracle, weights_dns3):
    # tests/models/test_gtcrn_micro.py:6-39 of the reference: identical 1 s prefix => bit-identical output prefix
    g = torch.Generator().manual_seed(3)
    a, b, c = (torch.randn(1, 16000, generator=g) for _ in range(3))
    y1 = oracle.enhance(weights_dns3, torch.cat([a, b], 1))
    y2 = oracle.enhance(weights_dns3, torch.cat([a, c], 1))
    assert (y1[:, :16000 - 512] - y2[:, :16000 - 512]).abs().max().item() == 0
    assert (y1[:, 16000:] - y2[:, 16000:]).abs().max().item() > 0


def test_stream_conv_primitives(oracle):
    # tests/streaming/conversion/test_convolution.py:35-85 of the reference, restated on the oracle's blocks:
    # a TCN / GTConv evaluated frame by frame with caches equals the zero-padded offline evaluation (atol 1e-6)
    g = torch.Generator().manual_seed(5)
    W = oracle.load_weights_npz(__import__("os").path.join(__import__("conftest").GOLDEN, "weights_rand7.npz"))
    x = torch.randn(1, 16, 10, 33, generator=g)
    with torch.inference_mode():
        for d, p in ((1, "gtcn1.blocks.0"), (8, "gtcn2.blocks.3")):
            off, _ = oracle.tcn_block(x, W, p, d)
            cache, outs = None, []
            for t in range(10):
                o, cache = oracle.tcn_block(x[:, :, t:t + 1], W, p, d, cache)
                outs.append(o)
            assert (torch.cat(outs, 2) - off).abs().max() <= 1e-5
        for dec, p in ((False, "encoder.en_convs.2"), (True, "decoder.de_convs.0")):
            off, _, _ = oracle.gt_conv_block(x, W, p, dec)
            cc = tc = None
            outs = []
            for t in range(10):
                o, cc, tc = oracle.gt_conv_block(x[:, :, t:t + 1], W, p, dec, cc, tc)
                outs.append(o)
            assert (torch.cat(outs, 2) - off).abs().max() <= 1e-5


def test_metrics_match_reference_values(oracle, metrics_golden):
    # oracle restatement of eval_intrusive_metrics.py:74-94 against values computed by the reference's own functions
    for i in range(4):
        r, x = metrics_golden[f"ref{i}"], metrics_golden[f"inf{i}"]
        assert abs(oracle.si_snr(r, x) - float(metrics_golden[f"si_snr{i}"])) <= 1e-9
        assert abs(oracle.sdr(r, x) - float(metrics_golden[f"sdr{i}"])) <= 1e-9
