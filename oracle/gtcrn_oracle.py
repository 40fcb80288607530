"""CPU oracle for the GTCRN-Micro inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (no nn.Module, no autograd, explicit weights dict) of the
reference algorithm `wav -> STFT -> GTCRNMicro.forward -> iSTFT -> wav`, offline and frame-by-frame
streaming.  It exists so that the CUDA path can be checked on a GPU box where /root/reference is
absent.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import it; the product package `gtcrn_micro_b200` never does (it fails loudly without its
CUDA extension).

Parity status: PINNED.  `oracle/check_against_reference.py` (run in the dev container, where the
reference is importable) verifies every function here against the reference's own modules, and
`tests/golden/*.npz` (made by `tests/golden/make_golden.py` from the reference itself) pin it on the
GPU box: the five shipped noisy->enh example pairs (<= 1 PCM16 LSB), reference forward / streaming
outputs on seeded inputs with the shipped checkpoint and with a randomised weight set.

The arithmetic lives in ATen (torch CPU ops), exactly as in the reference, whose hot path is pure
PyTorch (torch>=2.5.1 per its pyproject.toml:37); nothing in /root/reference implements conv/FFT
itself.  All functions take `dtype` from the inputs, so the same code gives the fp64 "truth" used
to judge which fp32 implementation is closer to exact arithmetic.

Reference anchors (paths relative to /root/reference):
  stft / istft            gtcrn_micro/infer.py:60-76
  features, forward       gtcrn_micro/models/gtcrn_micro.py:506-532
  ERB bm / bs             gtcrn_micro/models/gtcrn_micro.py:63-73
  SFE_Lite                gtcrn_micro/models/gtcrn_micro.py:77-90
  TRALite                 gtcrn_micro/models/gtcrn_micro.py:122-139
  ConvBlock               gtcrn_micro/models/gtcrn_micro.py:142-164
  GTConvBlock             gtcrn_micro/models/gtcrn_micro.py:167-253
  TCN / GTCN              gtcrn_micro/models/gtcrn_micro.py:256-336
  Encoder / Decoder       gtcrn_micro/models/gtcrn_micro.py:339-469
  Mask                    gtcrn_micro/models/gtcrn_micro.py:472-482
  streaming blocks        gtcrn_micro/streaming/gtcrn_micro_stream.py:182-574
  stream conv primitives  gtcrn_micro/streaming/conversion/convolution.py:62-253
  weight conversion       gtcrn_micro/streaming/conversion/convert.py:7-56
  SI-SNR                  gtcrn_micro/eval/eval_intrusive_metrics.py:74-82
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 512
HOP = 256
N_BINS = 257          # N_FFT // 2 + 1
ERB_LOW = 65          # bins passed through unchanged
ERB_BANDS = 64        # triangular bands covering bins 65..256
N_ERB = ERB_LOW + ERB_BANDS   # 129
BN_EPS = 1e-5
TCN_DILATIONS = (1, 2, 4, 8)

Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def load_weights_npz(path: str, dtype=torch.float32) -> Weights:
    """Load a state_dict saved as .npz (keys = reference state_dict keys, offline naming)."""
    z = np.load(path)
    out = {}
    for k in z.files:
        a = z[k]
        t = torch.from_numpy(np.array(a))
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out


def cast_weights(w: Weights, dtype) -> Weights:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in w.items()}


def stream_to_offline_keys(sd: Weights) -> Weights:
    """Inverse of convert.py:7-56: accept the streaming twin's state_dict (28 renamed keys, decoder
    3x3 weights stored as flipped [out,in,kt,kf] conv weights) and return offline-named weights."""
    out = {}
    for k, v in sd.items():
        if ".ConvTranspose2d." in k:
            nk = k.replace("ConvTranspose2d.", "")
            if nk.endswith("weight"):
                v = torch.flip(v, dims=[-2, -1]).permute(1, 0, 2, 3).contiguous()
            out[nk] = v
        elif ".Conv2d." in k:
            out[k.replace("Conv2d.", "")] = v
        else:
            out[k] = v
    return out


def erb_filter_bank(low: int = ERB_LOW, bands: int = ERB_BANDS, nfft: int = N_FFT,
                    high_hz: float = 8000.0, fs: float = 16000.0) -> np.ndarray:
    """Triangular ERB bank W[bands, nfft/2+1-low] (gtcrn_micro.py:35-61), float32."""
    def hz2erb(f):
        return 21.4 * np.log10(0.00437 * f + 1)

    def erb2hz(e):
        return (10 ** (e / 21.4) - 1) / 0.00437

    pts = np.linspace(hz2erb(low / nfft * fs), hz2erb(high_hz), bands)
    b = np.round(erb2hz(pts) / fs * nfft).astype(np.int32)
    W = np.zeros((bands, nfft // 2 + 1), dtype=np.float32)
    W[0, b[0]:b[1]] = (b[1] - np.arange(b[0], b[1]) + 1e-12) / (b[1] - b[0] + 1e-12)
    for i in range(bands - 2):
        W[i + 1, b[i]:b[i + 1]] = (np.arange(b[i], b[i + 1]) - b[i] + 1e-12) / (b[i + 1] - b[i] + 1e-12)
        W[i + 1, b[i + 1]:b[i + 2]] = (b[i + 2] - np.arange(b[i + 1], b[i + 2]) + 1e-12) / (b[i + 2] - b[i + 1] + 1e-12)
    W[-1, b[-2]:b[-1] + 1] = 1 - W[-2, b[-2]:b[-1] + 1]
    return np.abs(W[:, low:])


# --------------------------------------------------------------------------------------------
# STFT / iSTFT  (infer.py:60-76; torch.stft center=True, reflect, onesided, sqrt-Hann)
# --------------------------------------------------------------------------------------------
def sqrt_hann(dtype=torch.float32) -> torch.Tensor:
    # the reference builds the window in fp32 (torch.hann_window(512).pow(0.5)) -- keep that
    # rounding, then widen, so that fp64 runs use the same window values.
    return torch.hann_window(N_FFT).pow(0.5).to(dtype)


def stft(wav: torch.Tensor) -> torch.Tensor:
    """wav [B, L] -> spec [B, 257, T, 2] with T = 1 + L // 256.  Memory order is [B][T][F][2]
    (the tensor returned is the permuted view, like torch.stft(return_complex=False))."""
    B, L = wav.shape
    w = sqrt_hann(wav.dtype)
    xp = F.pad(wav[:, None, :], (HOP, HOP), mode="reflect")[:, 0]
    frames = xp.unfold(1, N_FFT, HOP) * w                    # [B, T, 512]
    X = torch.fft.rfft(frames, dim=-1)                      # [B, T, 257]
    return torch.view_as_real(X).permute(0, 2, 1, 3)        # [B, 257, T, 2]


def istft(spec: torch.Tensor) -> torch.Tensor:
    """spec [B, 257, T, 2] -> wav [B, 256 * (T - 1)]  (torch.istft with the same window)."""
    B, nb, T, _ = spec.shape
    w = sqrt_hann(spec.dtype)
    X = torch.view_as_complex(spec.permute(0, 2, 1, 3).contiguous())   # [B, T, 257]
    frames = torch.fft.irfft(X, n=N_FFT, dim=-1) * w                    # imag of DC / Nyquist ignored
    total = N_FFT + HOP * (T - 1)
    out = torch.zeros(B, total, dtype=spec.dtype)
    env = torch.zeros(total, dtype=spec.dtype)
    w2 = w * w
    for t in range(T):                                                  # plain overlap-add
        out[:, t * HOP:t * HOP + N_FFT] += frames[:, t]
        env[t * HOP:t * HOP + N_FFT] += w2
    out = out[:, HOP:total - HOP]
    env = env[HOP:total - HOP]
    return out / env


# --------------------------------------------------------------------------------------------
# layer helpers
# --------------------------------------------------------------------------------------------
def _bn(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"],
                        False, 0.0, BN_EPS)


def _prelu(x: torch.Tensor, w: Weights, p: str) -> torch.Tensor:
    return F.prelu(x, w[p + ".weight"])


def erb_compress(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """[..., 257] -> [..., 129]"""
    return torch.cat([x[..., :ERB_LOW], F.linear(x[..., ERB_LOW:], w["erb.erb_fc.weight"])], dim=-1)


def erb_expand(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """[..., 129] -> [..., 257]"""
    return torch.cat([x[..., :ERB_LOW], F.linear(x[..., ERB_LOW:], w["erb.ierb_fc.weight"])], dim=-1)


def features(spec: torch.Tensor) -> torch.Tensor:
    """spec [B,257,T,2] -> (mag, re, im) [B,3,T,257]"""
    re = spec[..., 0].permute(0, 2, 1)
    im = spec[..., 1].permute(0, 2, 1)
    mag = torch.sqrt(re ** 2 + im ** 2 + 1e-12)
    return torch.stack([mag, re, im], dim=1)


def sfe_lite(x: torch.Tensor, w: Weights) -> torch.Tensor:
    return F.conv2d(x, w["sfe.depth_conv.weight"], None, padding=(0, 1), groups=3)


def conv_block(x: torch.Tensor, w: Weights, p: str, deconv: bool, last: bool) -> torch.Tensor:
    """ConvBlock: (1,5) kernel, stride (1,2), pad (0,2); BN; PReLU or Tanh."""
    if deconv:
        y = F.conv_transpose2d(x, w[p + ".conv.weight"], w[p + ".conv.bias"], stride=(1, 2), padding=(0, 2))
    else:
        y = F.conv2d(x, w[p + ".conv.weight"], w[p + ".conv.bias"], stride=(1, 2), padding=(0, 2))
    y = _bn(y, w, p + ".bn")
    return torch.tanh(y) if last else _prelu(y, w, p + ".act")


def tra_lite(q: torch.Tensor, w: Weights, p: str, cache: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """q [B,8,T,F]; cache [B,8,2] = energies of the two previous frames (zeros at t<0)."""
    e = (q * q).mean(dim=3)                                             # [B,8,T]
    if cache is None:
        cache = torch.zeros(q.shape[0], q.shape[1], 2, dtype=q.dtype)
    ec = torch.cat([cache, e], dim=2)
    y = F.conv1d(ec, w[p + ".depth_conv.weight"], w[p + ".depth_conv.bias"], groups=q.shape[1])
    g = torch.sigmoid(F.conv1d(y, w[p + ".point_conv.weight"], w[p + ".point_conv.bias"]))
    return q * g[..., None], ec[:, :, -2:].contiguous()


def gt_conv_block(x: torch.Tensor, w: Weights, p: str, deconv: bool,
                  conv_cache: Optional[torch.Tensor] = None, tra_cache: Optional[torch.Tensor] = None):
    """GTConvBlock on T frames, continuing from caches (None = zeros = start of utterance).

    conv_cache [B,16,2,33]: point-conv-1 activations of the two previous frames.
    The decoder's ConvTranspose2d(16,16,(3,3),pad (0,1)) is evaluated as the equivalent causal
    convolution (weights transposed and flipped in kt,kf); the two extra trailing frames the
    reference produces and crops (gtcrn_micro.py:245-249) are never computed.
    Returns (out [B,16,T,33], conv_cache', tra_cache')."""
    B, C, T, Fq = x.shape
    x1, x2 = x[:, :8], x[:, 8:]
    if deconv:
        w1 = w[p + ".point_conv1.weight"].permute(1, 0, 2, 3)           # [in,out,1,1] -> [out,in,1,1]
    else:
        w1 = w[p + ".point_conv1.weight"]
    h = F.conv2d(x1, w1, w[p + ".point_conv1.bias"])
    h = _prelu(_bn(h, w, p + ".point_bn1"), w, p + ".point_act")
    if conv_cache is None:
        conv_cache = torch.zeros(B, 16, 2, Fq, dtype=x.dtype)
    hc = torch.cat([conv_cache, h], dim=2)                              # [B,16,T+2,F]
    if deconv:
        wd = torch.flip(w[p + ".depth_conv.weight"], dims=[-2, -1]).permute(1, 0, 2, 3).contiguous()
        d = F.conv2d(hc, wd, w[p + ".depth_conv.bias"], padding=(0, 1))
    else:
        d = F.conv2d(hc, w[p + ".depth_conv.weight"], w[p + ".depth_conv.bias"], padding=(0, 1), groups=16)
    d = _prelu(_bn(d, w, p + ".depth_bn"), w, p + ".depth_act")
    if deconv:
        w2 = w[p + ".point_conv2.weight"].permute(1, 0, 2, 3)
    else:
        w2 = w[p + ".point_conv2.weight"]
    q = _bn(F.conv2d(d, w2, w[p + ".point_conv2.bias"]), w, p + ".point_bn2")
    q, tra_cache = tra_lite(q, w, p + ".tra", tra_cache)
    out = torch.stack([q, x2], dim=2).reshape(B, 16, T, Fq)            # channel interleave [q0,x0,q1,x1,..]
    return out, hc[:, :, -2:].contiguous(), tra_cache


def tcn_block(x: torch.Tensor, w: Weights, p: str, d: int, cache: Optional[torch.Tensor] = None):
    """TCN with dilation d.  cache [B,16,2d,33] = conv1 activations of the 2d previous frames."""
    B, C, T, Fq = x.shape
    y1 = _prelu(_bn(F.conv2d(x, w[p + ".conv1.weight"], w[p + ".conv1.bias"]), w, p + ".bn1"), w, p + ".act1")
    if cache is None:
        cache = torch.zeros(B, C, 2 * d, Fq, dtype=x.dtype)
    yc = torch.cat([cache, y1], dim=2)
    y2 = F.conv2d(yc, w[p + ".conv2.weight"], w[p + ".conv2.bias"], dilation=(d, 1), groups=C)
    y2 = _prelu(_bn(y2, w, p + ".bn2"), w, p + ".act2")
    y3 = _bn(F.conv2d(y2, w[p + ".conv3.weight"], w[p + ".conv3.bias"]), w, p + ".bn3")
    return _prelu(y3 + x, w, p + ".act3"), yc[:, :, -2 * d:].contiguous()


# --------------------------------------------------------------------------------------------
# stream state (this oracle's own layout) + conversion from/to the reference's cache tensors
# --------------------------------------------------------------------------------------------
class StreamState:
    """gt[i]  [B,16,2,33]  i = 0..2 encoder blocks en2..en4, 3..5 decoder blocks de0..de2
       tra[i] [B,8,2]      same indexing
       tcn[g][k] [B,16,2*d_k,33]   g = 0,1 (gtcn1, gtcn2), k = 0..3"""

    def __init__(self, B: int, dtype=torch.float32):
        self.gt = [torch.zeros(B, 16, 2, 33, dtype=dtype) for _ in range(6)]
        self.tra = [torch.zeros(B, 8, 2, dtype=dtype) for _ in range(6)]
        self.tcn = [[torch.zeros(B, 16, 2 * d, 33, dtype=dtype) for d in TCN_DILATIONS] for _ in range(2)]

    # reference layout (gtcrn_micro_stream.py:618-623, slice map :416-429 / :490-500)
    @staticmethod
    def from_reference(conv_cache, tra_cache, tcn_cache) -> "StreamState":
        B = conv_cache.shape[1]
        s = StreamState(B, conv_cache.dtype)
        for i in range(3):
            s.gt[i] = conv_cache[0][:, :, 2 * i:2 * i + 2].clone()
            s.tra[i] = tra_cache[0][i].clone()
        for i, sl in enumerate((4, 2, 0)):                             # de0 uses [4:6], de1 [2:4], de2 [:2]
            s.gt[3 + i] = conv_cache[1][:, :, sl:sl + 2].clone()
            s.tra[3 + i] = tra_cache[1][i].clone()
        s.tcn = [[c.clone() for c in tcn_cache[g]] for g in range(2)]
        return s

    def to_reference(self):
        B = self.gt[0].shape[0]
        dt = self.gt[0].dtype
        conv_cache = torch.zeros(2, B, 16, 6, 33, dtype=dt)
        tra_cache = torch.zeros(2, 3, B, 8, 2, dtype=dt)
        for i in range(3):
            conv_cache[0][:, :, 2 * i:2 * i + 2] = self.gt[i]
            tra_cache[0][i] = self.tra[i]
        for i, sl in enumerate((4, 2, 0)):
            conv_cache[1][:, :, sl:sl + 2] = self.gt[3 + i]
            tra_cache[1][i] = self.tra[3 + i]
        tcn_cache = [[c.clone() for c in self.tcn[g]] for g in range(2)]
        return conv_cache, tra_cache, tcn_cache


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
def forward_spec(w: Weights, spec: torch.Tensor, state: Optional[StreamState] = None,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """GTCRNMicro.forward: spec [B,257,T,2] -> enhanced spec [B,257,T,2].

    state=None: offline (zero history).  With a StreamState the T frames continue the stream and
    the state is updated in place (T = 1 reproduces StreamGTCRNMicro.forward).
    taps: optional dict filled with intermediate activations (for kernel debugging)."""
    st = state

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    feat = features(spec)                                              # [B,3,T,257]
    feat = tap("erb", erb_compress(feat, w))                           # [B,3,T,129]
    feat = tap("sfe", sfe_lite(feat, w))
    en_outs: List[torch.Tensor] = []
    x = tap("en0", conv_block(feat, w, "encoder.en_convs.0", False, False)); en_outs.append(x)
    x = tap("en1", conv_block(x, w, "encoder.en_convs.1", False, False)); en_outs.append(x)
    for i in range(3):
        cc, tc = (st.gt[i], st.tra[i]) if st is not None else (None, None)
        x, cc, tc = gt_conv_block(x, w, f"encoder.en_convs.{2 + i}", False, cc, tc)
        if st is not None:
            st.gt[i], st.tra[i] = cc, tc
        tap(f"en{2 + i}", x); en_outs.append(x)
    for g in range(2):
        for k, d in enumerate(TCN_DILATIONS):
            c = st.tcn[g][k] if st is not None else None
            x, c = tcn_block(x, w, f"gtcn{g + 1}.blocks.{k}", d, c)
            if st is not None:
                st.tcn[g][k] = c
            tap(f"gtcn{g + 1}.{k}", x)
    for i in range(3):
        cc, tc = (st.gt[3 + i], st.tra[3 + i]) if st is not None else (None, None)
        x, cc, tc = gt_conv_block(x + en_outs[4 - i], w, f"decoder.de_convs.{i}", True, cc, tc)
        if st is not None:
            st.gt[3 + i], st.tra[3 + i] = cc, tc
        tap(f"de{i}", x)
    x = tap("de3", conv_block(x + en_outs[1], w, "decoder.de_convs.3", True, False))
    m = tap("de4", conv_block(x + en_outs[0], w, "decoder.de_convs.4", True, True))   # [B,2,T,129]
    m = tap("mask", erb_expand(m, w))                                  # [B,2,T,257]
    sr = spec[..., 0].permute(0, 2, 1)
    si = spec[..., 1].permute(0, 2, 1)
    yr = sr * m[:, 0] - si * m[:, 1]
    yi = si * m[:, 0] + sr * m[:, 1]
    return torch.stack([yr, yi], dim=-1).permute(0, 2, 1, 3)           # [B,257,T,2]


def stream_spec(w: Weights, spec: torch.Tensor, state: Optional[StreamState] = None) -> torch.Tensor:
    """Frame-by-frame evaluation (T = 1 per step), like the loop at gtcrn_micro_stream.py:626-636."""
    B, _, T, _ = spec.shape
    st = state if state is not None else StreamState(B, spec.dtype)
    outs = [forward_spec(w, spec[:, :, t:t + 1], st) for t in range(T)]
    return torch.cat(outs, dim=2)


def enhance(w: Weights, wav: torch.Tensor, streaming: bool = False) -> torch.Tensor:
    """wav [B,L] -> enhanced wav [B, 256*(L//256)]   (infer.py:60-76)."""
    with torch.inference_mode():
        spec = stft(wav)
        y = stream_spec(w, spec) if streaming else forward_spec(w, spec)
        return istft(y)


# --------------------------------------------------------------------------------------------
# metrics
# --------------------------------------------------------------------------------------------
def si_snr(ref: np.ndarray, inf: np.ndarray) -> float:
    """eval_intrusive_metrics.py:74-82"""
    ref = np.asarray(ref, dtype=np.float64)
    inf = np.asarray(inf, dtype=np.float64)
    inf = inf - inf.mean()
    ref = ref - ref.mean()
    a = np.sum(inf * ref) / np.sum(ref ** 2 + 1e-8)
    tgt = a * ref
    res = inf - tgt
    return float(10 * np.log10((np.sum(tgt ** 2) + 1e-8) / (np.sum(res ** 2) + 1e-8)))


def sdr(ref: np.ndarray, inf: np.ndarray) -> float:
    """eval_intrusive_metrics.py:85-94"""
    ref = np.asarray(ref, dtype=np.float64)
    inf = np.asarray(inf, dtype=np.float64)
    inf = inf - inf.mean()
    ref = ref - ref.mean()
    res = inf - ref
    return float(10 * np.log10((np.sum(ref ** 2) + 1e-8) / (np.sum(res ** 2) + 1e-8)))


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def synth_batch(B: int, L: int, seed: int = 1234, return_clean: bool = False):
    """Seeded 'speech-like' noisy batch: 5 harmonics of f0~U(90,250) Hz with a 4 Hz syllable
    envelope (RMS ~0.05) + white noise at SNR~U(0,15) dB, clipped to [-1,1].  float32 [B,L]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float64) / 16000.0
    f0 = 90 + 160 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    ph = 2 * math.pi * torch.rand(B, 5, generator=g, dtype=torch.float64)
    clean = torch.zeros(B, L, dtype=torch.float64)
    for h in range(5):
        clean += torch.sin(2 * math.pi * f0 * (h + 1) * t + ph[:, h:h + 1]) / (h + 1)
    env = 0.5 * (1 + torch.sin(2 * math.pi * 4.0 * t + 2 * math.pi * torch.rand(B, 1, generator=g, dtype=torch.float64)))
    clean = clean * env
    clean = clean * (0.05 / clean.pow(2).mean(dim=1, keepdim=True).sqrt())
    snr = 15 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    noise = torch.randn(B, L, generator=g, dtype=torch.float64)
    noise = noise * (0.05 / 10 ** (snr / 20))
    noisy = (clean + noise).clamp(-1, 1).float()
    if return_clean:
        return noisy, clean.float()
    return noisy
