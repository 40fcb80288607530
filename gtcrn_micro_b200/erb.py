"""ERB triangular filter bank (fixed, non-trainable weights of the model's ERB module).

Restates the constructor arithmetic of gtcrn_micro/models/gtcrn_micro.py:27-61 (`ERB.hz2erb`, `erb2hz`,
`erb_filter_banks`) so that a freshly constructed `GTCRNMicro` carries bit-identical `erb.erb_fc.weight` /
`erb.ierb_fc.weight` buffers (the shipped checkpoint stores exactly these values).
"""
from __future__ import annotations

import numpy as np


def erb_filter_bank(low_bins: int = 65, bands: int = 64, nfft: int = 512, high_hz: float = 8000.0,
                    fs: float = 16000.0) -> np.ndarray:
    """Returns W [bands, nfft/2 + 1 - low_bins] float32; rows overlap by 50 %, columns sum to 1."""
    def to_erb(hz):
        return 21.4 * np.log10(0.00437 * hz + 1)

    def to_hz(erb):
        return (10 ** (erb / 21.4) - 1) / 0.00437

    centres = np.linspace(to_erb(low_bins / nfft * fs), to_erb(high_hz), bands)
    edge = np.round(to_hz(centres) / fs * nfft).astype(np.int32)
    W = np.zeros((bands, nfft // 2 + 1), dtype=np.float32)
    # first band: falling ramp only
    W[0, edge[0]:edge[1]] = (edge[1] - np.arange(edge[0], edge[1]) + 1e-12) / (edge[1] - edge[0] + 1e-12)
    for b in range(1, bands - 1):
        lo, mid, hi = edge[b - 1], edge[b], edge[b + 1]
        W[b, lo:mid] = (np.arange(lo, mid) - lo + 1e-12) / (mid - lo + 1e-12)
        W[b, mid:hi] = (hi - np.arange(mid, hi) + 1e-12) / (hi - mid + 1e-12)
    # last band: complement of the previous one up to and including the Nyquist bin
    W[-1, edge[-2]:edge[-1] + 1] = 1 - W[-2, edge[-2]:edge[-1] + 1]
    return np.abs(W[:, low_bins:])
