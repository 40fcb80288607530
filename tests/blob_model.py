"""NumPy model of the arithmetic the CUDA kernels implement, driven ONLY by the packed weight blob.

Test helper (not shipped, not the oracle): it executes the same formulas and index maps as
`gtcrn_micro_b200/csrc/*.cu` -- folded BN, [k][ci][co] weights, gather-form transposed convs, sparse
ERB, radix-4 Stockham FFT with the packed twiddles -- so packer/layout/formula bugs are caught on CPU
by comparing with the oracle, before any GPU time is spent.  float64 throughout.
"""
from __future__ import annotations

import numpy as np

from gtcrn_micro_b200 import pack as P


class Blob:
    def __init__(self, blob: np.ndarray):
        assert blob.dtype == np.float32 and blob.size == P.blob_size()
        self.b = blob

    def f(self, chunk, field, idx=0, shape=None, as_int=False):
        off, n = P.chunk_fields(chunk)[field]
        base = P.chunk_offset(chunk, idx) + off
        v = self.b[base:base + n]
        if as_int:
            return v.view(np.int32).copy()
        v = v.astype(np.float64)
        if shape is not None:
            v = v[:int(np.prod(shape))].reshape(shape)
        return v


def prelu(x, a):
    return np.maximum(x, 0) + a * np.minimum(x, 0)


# ---------------------------------------------------------------------------------- FFT
def fft256_stockham(z, tw):
    """Forward 256-point complex FFT, radix-4 Stockham autosort, exactly the kernel's index map.
    z [..., 256] complex; tw [256] complex = the blob's stage-major twiddle table."""
    N, R = 256, 4
    a = z.copy()
    Ns = 1
    while Ns < N:
        out = np.empty_like(a)
        for j in range(N // R):
            k = j % Ns
            base = {1: 0, 4: 0, 16: 12, 64: 60}[Ns]               # stage-major twiddle table, 3 per butterfly (pack._fft_tables)
            v = [a[..., j + r * (N // R)] * (tw[base + 3 * k + (r - 1)] if (r > 0 and Ns > 1) else 1.0) for r in range(R)]
            # radix-4 butterfly (forward)
            s0, s1 = v[0] + v[2], v[0] - v[2]
            s2, s3 = v[1] + v[3], (v[1] - v[3]) * (-1j)
            y = [s0 + s2, s1 + s3, s0 - s2, s1 - s3]
            j0 = (j // Ns) * Ns * R + k
            for r in range(R):
                out[..., j0 + r * Ns] = y[r]
        a = out
        Ns *= R
    return a


def rfft512(x, tw256, tw512):
    """x [..., 512] real -> X [..., 257] complex via one 256-point complex FFT."""
    z = x[..., 0::2] + 1j * x[..., 1::2]
    Z = fft256_stockham(z, tw256)
    k = np.arange(257)
    Zk = Z[..., k % 256]
    Zc = np.conj(Z[..., (256 - k) % 256])
    return 0.5 * (Zk + Zc) + (-0.5j) * tw512[k] * (Zk - Zc)


def irfft512(X, tw256, tw512):
    """X [..., 257] complex -> x [..., 512] real (imag of DC / Nyquist ignored, 1/512 scaling).
    Inverse FFT computed with the forward kernel:  ifft(Z) = conj(fft(conj(Z))) / 256."""
    X = X.copy()
    X[..., 0] = X[..., 0].real
    X[..., 256] = X[..., 256].real
    k = np.arange(256)
    Xk = X[..., k]
    Xc = np.conj(X[..., 256 - k])
    Z = 0.5 * (Xk + Xc) + 0.5j * np.conj(tw512[k]) * (Xk - Xc)
    z = np.conj(fft256_stockham(np.conj(Z), tw256)) / 256.0
    x = np.empty(X.shape[:-1] + (512,), np.float64)
    x[..., 0::2] = z.real
    x[..., 1::2] = z.imag
    return x


def stft(B: Blob, wav):
    """wav [nb, L] -> spec [nb, T, 257] complex"""
    nb, L = wav.shape
    T = 1 + L // 256
    win = B.f("FFT", "WIN")
    tw256 = B.f("FFT", "TW256", shape=(256, 2)); tw256 = tw256[:, 0] + 1j * tw256[:, 1]
    tw512 = B.f("FFT", "TW512", shape=(257, 2)); tw512 = tw512[:, 0] + 1j * tw512[:, 1]
    idx = (np.arange(T)[:, None] * 256 + np.arange(512)[None, :]) - 256
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, 2 * (L - 1) - idx, idx)
    frames = wav[:, idx] * win
    return rfft512(frames, tw256, tw512)


def istft(B: Blob, spec):
    """spec [nb, T, 257] complex -> wav [nb, 256 (T-1)]"""
    nb, T, _ = spec.shape
    win = B.f("FFT", "WIN")
    inv_env = B.f("FFT", "INV_ENV")
    tw256 = B.f("FFT", "TW256", shape=(256, 2)); tw256 = tw256[:, 0] + 1j * tw256[:, 1]
    tw512 = B.f("FFT", "TW512", shape=(257, 2)); tw512 = tw512[:, 0] + 1j * tw512[:, 1]
    fr = irfft512(spec, tw256, tw512) * win
    out = (fr[:, :-1, 256:] + fr[:, 1:, :256]) * inv_env
    return out.reshape(nb, -1)


# ---------------------------------------------------------------------------------- model stages
def front(B: Blob, spec):
    """spec [nb,T,257] complex -> en0 [nb,T,65,16], en1 [nb,T,33,16]  (channel-last)."""
    re, im = spec.real, spec.imag
    mag = np.sqrt(re * re + im * im + 1e-12)
    feat = np.stack([mag, re, im], axis=-1)                     # [nb,T,257,3]
    st = B.f("FRONT", "ERB_START", as_int=True); cn = B.f("FRONT", "ERB_CNT", as_int=True)
    ew = B.f("FRONT", "ERB_W", shape=(64, P.ERB_MAXTAP))
    erb = np.zeros(feat.shape[:2] + (129, 3))
    erb[:, :, :65] = feat[:, :, :65]
    for b in range(64):
        for j in range(cn[b]):
            erb[:, :, 65 + b] += ew[b, j] * feat[:, :, 65 + st[b] + j]
    sw = B.f("FRONT", "SFE_W", shape=(3, 3))                    # [c][k]
    ep = np.pad(erb, ((0, 0), (0, 0), (1, 1), (0, 0)))
    sfe = sum(ep[:, :, k:k + 129] * sw[:, k] for k in range(3))
    a = B.f("FRONT", "SLOPE")
    w0 = B.f("FRONT", "EN0_W", shape=(5, 3, 16)); b0 = B.f("FRONT", "EN0_B")
    sp = np.pad(sfe, ((0, 0), (0, 0), (2, 2), (0, 0)))
    en0 = b0 + sum(np.einsum("ntfi,io->ntfo", sp[:, :, k:k + 129:2][:, :, :65], w0[k]) for k in range(5))
    en0 = prelu(en0, a[0])
    w1 = B.f("FRONT", "EN1_W", shape=(5, 16, 16)); b1 = B.f("FRONT", "EN1_B")
    xp = np.pad(en0, ((0, 0), (0, 0), (2, 2), (0, 0)))
    en1 = b1 + sum(np.einsum("ntfi,io->ntfo", xp[:, :, k:k + 65:2][:, :, :33], w1[k]) for k in range(5))
    en1 = prelu(en1, a[1])
    return en0, en1


def tra_gate(B: Blob, chunk, idx, q):
    """q [nb,T,33,8] -> gated q (zero history)."""
    e = (q * q).sum(axis=2) * (1.0 / 33.0)                      # [nb,T,8]
    wt = B.f(chunk, "TRA_DW_W", idx, (3, 8)); bt = B.f(chunk, "TRA_DW_B", idx)
    wp = B.f(chunk, "TRA_PW_W", idx, (8, 8)); bp = B.f(chunk, "TRA_PW_B", idx)
    ep = np.pad(e, ((0, 0), (2, 0), (0, 0)))
    T = q.shape[1]
    y = bt + sum(wt[k] * ep[:, k:k + T] for k in range(3))
    g = 1.0 / (1.0 + np.exp(-(y @ wp + bp)))
    return q * g[:, :, None, :]


def gtconv(B: Blob, x, idx, dec):
    """x [nb,T,33,16] -> [nb,T,33,16]"""
    chunk = "DECGT" if dec else "ENCGT"
    a = B.f(chunk, "SLOPE", idx)
    x1, x2 = x[..., :8], x[..., 8:]
    p = prelu(x1 @ B.f(chunk, "PC1_W", idx, (8, 16)) + B.f(chunk, "PC1_B", idx), a[0])
    T = x.shape[1]
    pp = np.pad(p, ((0, 0), (2, 0), (1, 1), (0, 0)))             # 2 zero frames at t<0, f zero pad
    if dec:
        wf = B.f(chunk, "FULL_W", idx, (3, 3, 16, 16))
        d = B.f(chunk, "FULL_B", idx) + sum(
            np.einsum("ntfi,io->ntfo", pp[:, kt:kt + T, kf:kf + 33], wf[kt, kf]) for kt in range(3) for kf in range(3))
    else:
        wd = B.f(chunk, "DW_W", idx, (3, 3, 16))
        d = B.f(chunk, "DW_B", idx) + sum(pp[:, kt:kt + T, kf:kf + 33] * wd[kt, kf] for kt in range(3) for kf in range(3))
    d = prelu(d, a[1])
    q = d @ B.f(chunk, "PC2_W", idx, (16, 8)) + B.f(chunk, "PC2_B", idx)
    q = tra_gate(B, chunk, idx, q)
    out = np.empty_like(x)
    out[..., 0::2] = q
    out[..., 1::2] = x2
    return out


def tcn(B: Blob, x, idx, d):
    a = B.f("TCN", "SLOPE", idx)
    y1 = prelu(x @ B.f("TCN", "W1", idx, (16, 16)) + B.f("TCN", "B1", idx), a[0])
    T = x.shape[1]
    yp = np.pad(y1, ((0, 0), (2 * d, 0), (0, 0), (0, 0)))
    w2 = B.f("TCN", "W2", idx, (3, 16))
    y2 = prelu(B.f("TCN", "B2", idx) + sum(w2[k] * yp[:, k * d:k * d + T] for k in range(3)), a[1])
    return prelu(y2 @ B.f("TCN", "W3", idx, (16, 16)) + B.f("TCN", "B3", idx) + x, a[2])


def deconv_s2(x, w, b, Fout):
    """gather-form ConvTranspose (1,5), stride 2, pad 2:  out[g] = sum_{k: (g+2-k) even} w[k] x[(g+2-k)/2]"""
    nb, T, Fin, Ci = x.shape
    out = np.zeros((nb, T, Fout, w.shape[2])) + b
    for g in range(Fout):
        for k in range(5):
            n = g + 2 - k
            if n % 2 == 0 and 0 <= n // 2 < Fin:
                out[:, :, g] += x[:, :, n // 2] @ w[k]
    return out


def back(B: Blob, x, en1, en0, spec):
    """x = de2 out [nb,T,33,16]; returns enhanced spec [nb,T,257] complex and the mask [nb,T,257,2]"""
    a = B.f("BACK", "SLOPE")
    y = prelu(deconv_s2(x + en1, B.f("BACK", "DE3_W", shape=(5, 16, 16)), B.f("BACK", "DE3_B"), 65), a[0])
    m = np.tanh(deconv_s2(y + en0, B.f("BACK", "DE4_W", shape=(5, 16, 2)), B.f("BACK", "DE4_B")[:2], 129))
    b0 = B.f("BACK", "IERB_B0", as_int=True); iw = B.f("BACK", "IERB_W", shape=(192, 2))
    mask = np.zeros(m.shape[:2] + (257, 2))
    mask[:, :, :65] = m[:, :, :65]
    for i in range(192):
        mask[:, :, 65 + i] = iw[i, 0] * m[:, :, 65 + b0[i]] + iw[i, 1] * m[:, :, 65 + b0[i] + 1]
    yr = spec.real * mask[..., 0] - spec.imag * mask[..., 1]
    yi = spec.imag * mask[..., 0] + spec.real * mask[..., 1]
    return yr + 1j * yi, mask


def forward(B: Blob, spec, taps=None):
    """spec [nb,T,257] complex -> enhanced spec [nb,T,257] complex"""
    en = [None] * 5
    en[0], en[1] = front(B, spec)
    x = en[1]
    for i in range(3):
        x = gtconv(B, x, i, False); en[2 + i] = x
    for g in range(2):
        for k, d in enumerate((1, 2, 4, 8)):
            x = tcn(B, x, g * 4 + k, d)
    g2 = x
    de = []
    for i in range(3):
        x = gtconv(B, x + en[4 - i], i, True); de.append(x)
    y, mask = back(B, x, en[1], en[0], spec)
    if taps is not None:
        taps.update(en0=en[0], en1=en[1], en2=en[2], en3=en[3], en4=en[4], gtcn2=g2, de0=de[0], de1=de[1], de2=de[2], mask=mask)
    return y


def enhance(B: Blob, wav):
    return istft(B, forward(B, stft(B, wav.astype(np.float64))))
