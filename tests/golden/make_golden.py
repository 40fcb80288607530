"""Generate the committed golden fixtures from the REAL reference (dev container only).

    python tests/golden/make_golden.py

Writes into tests/golden/:
  weights_dns3.npz     ckpt["model"] of gtcrn_micro/ckpts/best_model_dns3.tar (388 keys, offline naming)
  weights_rand7.npz    a randomised weight set (non-trivial BN stats / PReLU slopes), seed 7
  pairs.npz            first 80,000 samples (5 s) of noisy{1..5}.wav and the first 79,616 samples of
                       enh{1..5}.wav (int16).  The model is causal and frame t only sees samples
                       < 256(t+1), so output[:256*311] of the cropped input equals the full-file output.
  ref_outputs.npz      seeded input + the reference's own outputs (offline spec, streaming spec, wav)
                       for both weight sets.
  metrics.npz          waveform pairs with the reference's own SI-SNR / SDR values (eval_intrusive_metrics.py:74-94)
Nothing from the reference's sources is copied; only its weights / audio data and its outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from check_against_reference import REF, import_reference, randomise, read_wav_int16, reference_metric_functions  # noqa: E402


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def make_metrics():
    """metrics.npz: float32 (reference, degraded) pairs and the reference's own SI-SNR / SDR for them."""
    sisnr_ref, sdr_ref = reference_metric_functions()
    g = np.random.default_rng(20261017)
    out = {}
    for i, (n, gain, noise, dc) in enumerate(((4000, 0.8, 0.3, 0.01), (16000, 1.0, 0.02, 0.0), (16000, 0.5, 1.0, -0.02), (257, 1.2, 0.1, 0.0))):
        t = np.arange(n) / 16000.0
        r = (0.1 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.01 * g.standard_normal(n)).astype(np.float32)
        x = (gain * r + noise * 0.05 * g.standard_normal(n) + dc).astype(np.float32)
        out[f"ref{i}"], out[f"inf{i}"] = r, x
        out[f"si_snr{i}"] = np.float64(sisnr_ref(r.astype(np.float64), x.astype(np.float64)))
        out[f"sdr{i}"] = np.float64(sdr_ref(r.astype(np.float64), x.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)


def main():
    if "--only-metrics" in sys.argv:
        make_metrics()
        return
    make_metrics()
    GTCRNMicro, StreamGTCRNMicro, convert_to_stream = import_reference()
    ckpt = torch.load(os.path.join(REF, "gtcrn_micro/ckpts/best_model_dns3.tar"), map_location="cpu")
    win = torch.hann_window(512).pow(0.5)
    models = {}
    m = GTCRNMicro().eval(); m.load_state_dict(ckpt["model"]); models["dns3"] = m
    m = GTCRNMicro().eval(); randomise(m, 7); models["rand7"] = m
    for tag, m in models.items():
        np.savez_compressed(os.path.join(HERE, f"weights_{tag}.npz"), **sd_to_np(m.state_dict()))

    pairs = {}
    for i in range(1, 6):
        noisy = read_wav_int16(os.path.join(REF, f"gtcrn_micro/examples/gtcrn_micro/noisy{i}.wav"))
        enh = read_wav_int16(os.path.join(REF, f"gtcrn_micro/examples/gtcrn_micro/enh{i}.wav"))
        pairs[f"noisy{i}"] = noisy[:80000].copy()
        pairs[f"enh{i}"] = enh[:256 * 311].copy()
    np.savez_compressed(os.path.join(HERE, "pairs.npz"), **pairs)

    out = {}
    g = torch.Generator().manual_seed(20261017)
    wav = 0.1 * torch.randn(2, 256 * 40 + 77, generator=g)
    out["wav"] = wav.numpy()
    with torch.no_grad():
        spec = torch.stft(wav, 512, 256, 512, win, return_complex=False)
        out["spec"] = spec.contiguous().numpy()
        for tag, m in models.items():
            y = m(spec)
            out[f"{tag}.spec_out"] = y.contiguous().numpy()
            out[f"{tag}.wav_out"] = torch.istft(torch.view_as_complex(y.contiguous()), 512, 256, 512, win).numpy()
            sm = StreamGTCRNMicro().eval()
            convert_to_stream(sm, m)
            B = wav.shape[0]
            cc = torch.zeros(2, B, 16, 6, 33)
            tc = torch.zeros(2, 3, B, 8, 2)
            nc = [[torch.zeros(B, 16, 2 * d, 33) for d in (1, 2, 4, 8)] for _ in range(2)]
            ys = []
            for t in range(spec.shape[2]):
                yi, cc, tc, nc = sm(spec[:, :, t:t + 1], cc, tc, nc)
                ys.append(yi)
                if t == 20:   # cache snapshot in the reference's own layout
                    out[f"{tag}.conv_cache20"] = cc.numpy().copy()
                    out[f"{tag}.tra_cache20"] = tc.numpy().copy()
                    for gi in range(2):
                        for k in range(4):
                            out[f"{tag}.tcn_cache20.{gi}.{k}"] = nc[gi][k].numpy().copy()
            out[f"{tag}.stream_spec_out"] = torch.cat(ys, dim=2).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"{f:24s} {os.path.getsize(os.path.join(HERE, f)) / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
