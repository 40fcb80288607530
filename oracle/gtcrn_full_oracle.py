"""CPU oracle for BASELINE configs[3]: the ORIGINAL-GTCRN-style architecture (unfold SFE, grouped convs, dilated depthwise
GTConv, GRU-based TRA, DPGRNN).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED -- no reference code / checkpoint / test exists for this variant.
/root/reference (bglid/GTCRN-Micro) replaced these blocks ("Replacing DPGRNN approach with grouped TCNs",
gtcrn_micro/models/gtcrn_micro.py:5) and keeps them only as commented-out lines, which are the sole corroboration available:
  gtcrn_micro.py:187   `# self.point_conv1 = conv_module(in_channels // 2 * 3, hidden_channels, 1)`   -> SFE inside GTConvBlock (24 inputs)
  gtcrn_micro.py:199,210 `# groups=hidden_channels`                                                   -> depthwise 3x3 in conv AND deconv
  gtcrn_micro.py:346   `# 3 * 3,  # SFE`                                                              -> en0 takes 9 channels
  gtcrn_micro.py:361,448 `# groups=2`                                                                 -> en1 / de3 grouped
  gtcrn_micro.py:381,391 `# dilation=(2, 1)` / `# dilation=(5, 1)`                                    -> encoder dilations 1, 2, 5
  gtcrn_micro.py:417-418,429 `# padding=(2 * 5, 1)`, `# dilation=(5, 1)`, `# dilation=(2, 1)`         -> decoder dilations 5, 2, 1
  gtcrn_micro.py:494   `# self.sfe = SFE(3, 1)`                                                       -> unfold-based SFE on the input
  gtcrn_micro.py:5, gtcrn_micro/README.md:4  "Swapped GRU blocks + DPGRNN for a psuedo Grouped-TCN"   -> DPGRNN x2 where gtcn1/gtcn2 are
The block definitions themselves (GRU-TRA, GRNN, DPGRNN, layer sizes) follow the public upstream GTCRN model the reference
derives from, as recalled in SURVEY.md Appendix C; they cannot be verified from /root/reference.  Weights are seeded random
(`random_weights`).  What IS checked: this functional restatement against an independent nn.Module construction of the same
architecture built from stock torch layers (nn.GRU, nn.ConvTranspose2d with the upstream padding / dilation / groups, nn.Unfold,
nn.LayerNorm) in tests/full_reference_model.py, and the GRU cell equations against torch.nn.GRU.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / parity legs may import this file.

Shapes follow gtcrn_oracle.py: spec [B,257,T,2], activations [B,C,T,F].  All ops causal in T.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from gtcrn_oracle import (BN_EPS, ERB_LOW, _bn, _prelu, erb_compress, erb_expand, erb_filter_bank, features, istft, stft)  # noqa: F401

ENC_DIL = (1, 2, 5)            # gtcrn_micro.py:381,391 (commented upstream values)
DEC_DIL = (5, 2, 1)            # gtcrn_micro.py:417-418,429
LN_EPS = 1e-8
Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# parameter tree (upstream key names) and seeded random weights
# --------------------------------------------------------------------------------------------
def _bn_spec(p, n):
    return [(f"{p}.weight", (n,), "bn_w"), (f"{p}.bias", (n,), "bn_b"), (f"{p}.running_mean", (n,), "bn_m"), (f"{p}.running_var", (n,), "bn_v"),
            (f"{p}.num_batches_tracked", (), "bn_n")]


def _gru_spec(p, n_in, n_h, bidirectional=False):
    out = []
    for suf in ("", "_reverse") if bidirectional else ("",):
        out += [(f"{p}.weight_ih_l0{suf}", (3 * n_h, n_in), "w"), (f"{p}.weight_hh_l0{suf}", (3 * n_h, n_h), "w"),
                (f"{p}.bias_ih_l0{suf}", (3 * n_h,), "b"), (f"{p}.bias_hh_l0{suf}", (3 * n_h,), "b")]
    return out


def param_spec():
    """(key, shape, kind) of the full model's state_dict, in upstream naming."""
    s = [("erb.erb_fc.weight", (64, 192), "erb"), ("erb.ierb_fc.weight", (192, 64), "erbT")]
    s += [("encoder.en_convs.0.conv.weight", (16, 9, 1, 5), "w"), ("encoder.en_convs.0.conv.bias", (16,), "b")] + _bn_spec("encoder.en_convs.0.bn", 16) + [("encoder.en_convs.0.act.weight", (1,), "prelu")]
    s += [("encoder.en_convs.1.conv.weight", (16, 8, 1, 5), "w"), ("encoder.en_convs.1.conv.bias", (16,), "b")] + _bn_spec("encoder.en_convs.1.bn", 16) + [("encoder.en_convs.1.act.weight", (1,), "prelu")]

    def gt(p, deconv):
        o = [(f"{p}.point_conv1.weight", (24, 16, 1, 1) if deconv else (16, 24, 1, 1), "w"), (f"{p}.point_conv1.bias", (16,), "b")] + _bn_spec(f"{p}.point_bn1", 16)
        o += [(f"{p}.point_act.weight", (1,), "prelu"), (f"{p}.depth_conv.weight", (16, 1, 3, 3), "w"), (f"{p}.depth_conv.bias", (16,), "b")] + _bn_spec(f"{p}.depth_bn", 16)
        o += [(f"{p}.depth_act.weight", (1,), "prelu"), (f"{p}.point_conv2.weight", (16, 8, 1, 1) if deconv else (8, 16, 1, 1), "w"), (f"{p}.point_conv2.bias", (8,), "b")]
        o += _bn_spec(f"{p}.point_bn2", 8) + _gru_spec(f"{p}.tra.att_gru", 8, 16) + [(f"{p}.tra.att_fc.weight", (8, 16), "w"), (f"{p}.tra.att_fc.bias", (8,), "b")]
        return o
    for i in (2, 3, 4):
        s += gt(f"encoder.en_convs.{i}", False)
    for g in (1, 2):
        p = f"dpgrnn{g}"
        for r in (1, 2):
            s += _gru_spec(f"{p}.intra_rnn.rnn{r}", 8, 4, bidirectional=True)
        s += [(f"{p}.intra_fc.weight", (16, 16), "w"), (f"{p}.intra_fc.bias", (16,), "b"), (f"{p}.intra_ln.weight", (33, 16), "ln_w"), (f"{p}.intra_ln.bias", (33, 16), "ln_b")]
        for r in (1, 2):
            s += _gru_spec(f"{p}.inter_rnn.rnn{r}", 8, 8)
        s += [(f"{p}.inter_fc.weight", (16, 16), "w"), (f"{p}.inter_fc.bias", (16,), "b"), (f"{p}.inter_ln.weight", (33, 16), "ln_w"), (f"{p}.inter_ln.bias", (33, 16), "ln_b")]
    for i in (0, 1, 2):
        s += gt(f"decoder.de_convs.{i}", True)
    s += [("decoder.de_convs.3.conv.weight", (16, 8, 1, 5), "w"), ("decoder.de_convs.3.conv.bias", (16,), "b")] + _bn_spec("decoder.de_convs.3.bn", 16) + [("decoder.de_convs.3.act.weight", (1,), "prelu")]
    s += [("decoder.de_convs.4.conv.weight", (16, 2, 1, 5), "w"), ("decoder.de_convs.4.conv.bias", (2,), "b")] + _bn_spec("decoder.de_convs.4.bn", 2)
    return s


def random_weights(seed: int = 0, dtype=torch.float32) -> Weights:
    """Seeded weights with non-trivial BN statistics, PReLU slopes and LayerNorm affine terms (no checkpoint exists)."""
    g = torch.Generator().manual_seed(seed)
    W = {}
    for key, shape, kind in param_spec():
        if kind == "erb":
            t = torch.from_numpy(erb_filter_bank())
        elif kind == "erbT":
            t = torch.from_numpy(np.ascontiguousarray(erb_filter_bank().T))
        elif kind == "bn_n":
            t = torch.tensor(1000, dtype=torch.long)
        elif kind == "bn_v":
            t = 0.3 + 1.5 * torch.rand(shape, generator=g)
        elif kind in ("bn_w", "ln_w"):
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif kind in ("bn_b", "bn_m", "ln_b"):
            t = 0.3 * torch.randn(shape, generator=g)
        elif kind == "prelu":
            t = torch.tensor([0.05 + 0.5 * float(torch.rand((), generator=g))])
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            if key.endswith("depth_conv.weight"):
                fan_in = 9
            t = (torch.rand(shape, generator=g) * 2 - 1) / np.sqrt(max(1, fan_in)) * (1.7 if kind == "w" else 1.0)
        W[key] = t.to(dtype) if t.is_floating_point() else t
    return W


def n_learned_params(W: Weights) -> int:
    return sum(v.numel() for k, v in W.items() if v.is_floating_point() and not k.startswith("erb.") and "running_" not in k)


def cast_weights(W: Weights, dtype) -> Weights:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in W.items()}


def save_weights_npz(path: str, W: Weights) -> None:
    np.savez_compressed(path, **{k: v.numpy() for k, v in W.items()})


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
def sfe_unfold(x: torch.Tensor) -> torch.Tensor:
    """SFE(kernel 3, stride 1) = nn.Unfold((1,3), padding (0,1)) reshaped: [B,C,T,F] -> [B,3C,T,F], channel c*3+k = x[c, :, f+k-1]
    (zeros outside).  gtcrn_micro.py:494 (commented `SFE(3, 1)`)."""
    xp = F.pad(x, (1, 1))
    return torch.stack([xp[..., 0:-2], xp[..., 1:-1], xp[..., 2:]], dim=2).reshape(x.shape[0], x.shape[1] * 3, x.shape[2], x.shape[3])


def gru_cell(x: torch.Tensor, h: torch.Tensor, w: Weights, p: str, suf: str = "") -> torch.Tensor:
    """One torch.nn.GRU step (gate order r, z, n):  r = s(Wir x + bir + Whr h + bhr), z likewise,
    n = tanh(Win x + bin + r * (Whn h + bhn)), h' = (1 - z) * n + z * h."""
    H = h.shape[-1]
    gi = F.linear(x, w[f"{p}.weight_ih_l0{suf}"], w[f"{p}.bias_ih_l0{suf}"])
    gh = F.linear(h, w[f"{p}.weight_hh_l0{suf}"], w[f"{p}.bias_hh_l0{suf}"])
    r = torch.sigmoid(gi[..., :H] + gh[..., :H])
    z = torch.sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H])
    n = torch.tanh(gi[..., 2 * H:] + r * gh[..., 2 * H:])
    return (1 - z) * n + z * h


def gru_seq(x: torch.Tensor, h0: torch.Tensor, w: Weights, p: str, reverse: bool = False):
    """x [N,S,I], h0 [N,H] -> (y [N,S,H], h_last)."""
    S = x.shape[1]
    h, ys = h0, [None] * S
    for s in (range(S - 1, -1, -1) if reverse else range(S)):
        h = gru_cell(x[:, s], h, w, p, "_reverse" if reverse else "")
        ys[s] = h
    return torch.stack(ys, dim=1), h


def tra_gru(q: torch.Tensor, w: Weights, p: str, h: Optional[torch.Tensor]):
    """TRA: zt = mean_F(q^2) [B,C,T] -> GRU(C -> 2C) over T -> Linear(2C -> C) -> sigmoid -> gate.  h [B,2C] carried state."""
    B, C = q.shape[0], q.shape[1]
    zt = (q * q).mean(dim=3).transpose(1, 2)                            # [B,T,C]
    if h is None:
        h = torch.zeros(B, 2 * C, dtype=q.dtype)
    at, h = gru_seq(zt, h, w, p + ".att_gru")
    at = torch.sigmoid(F.linear(at, w[p + ".att_fc.weight"], w[p + ".att_fc.bias"])).transpose(1, 2)   # [B,C,T]
    return q * at[..., None], h


def conv_block(x, w, p, deconv, last, groups=1):
    if deconv:
        y = F.conv_transpose2d(x, w[p + ".conv.weight"], w[p + ".conv.bias"], stride=(1, 2), padding=(0, 2), groups=groups)
    else:
        y = F.conv2d(x, w[p + ".conv.weight"], w[p + ".conv.bias"], stride=(1, 2), padding=(0, 2), groups=groups)
    y = _bn(y, w, p + ".bn")
    return torch.tanh(y) if last else _prelu(y, w, p + ".act")


def gt_conv_block(x, w, p, deconv: bool, d: int, conv_cache=None, tra_h=None):
    """Full GTConvBlock on T frames continuing from caches.  conv_cache [B,16,2d,33] = point-conv-1 activations of the 2d
    previous frames (zeros = causal padding); tra_h [B,16] = TRA GRU state.
    The decoder's depthwise ConvTranspose2d(k (3,3), padding (2d,1), dilation (d,1)) applied to the causally padded input equals
    the causal depthwise convolution with the kernel flipped in both axes (checked against the stock layer in the tests)."""
    B, C, T, Fq = x.shape
    x1, x2 = x[:, :8], x[:, 8:]
    w1 = w[p + ".point_conv1.weight"]
    h = F.conv2d(sfe_unfold(x1), w1.permute(1, 0, 2, 3) if deconv else w1, w[p + ".point_conv1.bias"])
    h = _prelu(_bn(h, w, p + ".point_bn1"), w, p + ".point_act")
    if conv_cache is None:
        conv_cache = torch.zeros(B, 16, 2 * d, Fq, dtype=x.dtype)
    hc = torch.cat([conv_cache, h], dim=2)
    wd = w[p + ".depth_conv.weight"]
    if deconv:
        wd = torch.flip(wd, dims=[-2, -1])
    dd = F.conv2d(hc, wd, w[p + ".depth_conv.bias"], padding=(0, 1), dilation=(d, 1), groups=16)
    dd = _prelu(_bn(dd, w, p + ".depth_bn"), w, p + ".depth_act")
    w2 = w[p + ".point_conv2.weight"]
    q = _bn(F.conv2d(dd, w2.permute(1, 0, 2, 3) if deconv else w2, w[p + ".point_conv2.bias"]), w, p + ".point_bn2")
    q, tra_h = tra_gru(q, w, p + ".tra", tra_h)
    out = torch.stack([q, x2], dim=2).reshape(B, 16, T, Fq)
    return out, hc[:, :, -2 * d:].contiguous(), tra_h


def dpgrnn(x, w, p, inter_h=None, taps=None):
    """DPGRNN(16, 33, 16).  x [B,16,T,33]; inter_h [B*33,16] carried state of the inter (time) GRUs (rnn1 | rnn2 halves)."""
    B, C, T, Fq = x.shape
    xt = x.permute(0, 2, 3, 1)                                          # [B,T,F,C]
    a = xt.reshape(B * T, Fq, C)
    ys = []
    for r in (1, 2):                                                    # GRNN: channel halves -> separate bidirectional GRUs (hidden 4)
        xa = a[..., 8 * (r - 1):8 * r]
        z0 = torch.zeros(B * T, 4, dtype=x.dtype)
        yf, _ = gru_seq(xa, z0, w, f"{p}.intra_rnn.rnn{r}")
        yb, _ = gru_seq(xa, z0, w, f"{p}.intra_rnn.rnn{r}", reverse=True)
        ys += [yf, yb]
    a = F.linear(torch.cat(ys, dim=-1), w[p + ".intra_fc.weight"], w[p + ".intra_fc.bias"]).reshape(B, T, Fq, C)
    a = F.layer_norm(a, (Fq, C), w[p + ".intra_ln.weight"], w[p + ".intra_ln.bias"], LN_EPS)
    intra = xt + a
    if taps is not None:
        taps[p + ".intra"] = intra.permute(0, 3, 1, 2)
    e = intra.permute(0, 2, 1, 3).reshape(B * Fq, T, C)                 # [B*F,T,C]
    if inter_h is None:
        inter_h = torch.zeros(B * Fq, 16, dtype=x.dtype)
    y1, h1 = gru_seq(e[..., :8], inter_h[:, :8], w, f"{p}.inter_rnn.rnn1")
    y2, h2 = gru_seq(e[..., 8:], inter_h[:, 8:], w, f"{p}.inter_rnn.rnn2")
    e = F.linear(torch.cat([y1, y2], dim=-1), w[p + ".inter_fc.weight"], w[p + ".inter_fc.bias"])
    e = e.reshape(B, Fq, T, C).permute(0, 2, 1, 3)                      # [B,T,F,C]
    e = F.layer_norm(e, (Fq, C), w[p + ".inter_ln.weight"], w[p + ".inter_ln.bias"], LN_EPS)
    return (intra + e).permute(0, 3, 1, 2), torch.cat([h1, h2], dim=-1)


class StreamState:
    """gt[i] [B,16,2d_i,33] (i = 0..2 encoder d = 1,2,5; 3..5 decoder d = 5,2,1), tra[i] [B,16], inter[g] [B*33,16]."""

    def __init__(self, B: int, dtype=torch.float32):
        self.gt = [torch.zeros(B, 16, 2 * d, 33, dtype=dtype) for d in ENC_DIL + DEC_DIL]
        self.tra = [torch.zeros(B, 16, dtype=dtype) for _ in range(6)]
        self.inter = [torch.zeros(B * 33, 16, dtype=dtype) for _ in range(2)]


def forward_spec(w: Weights, spec: torch.Tensor, state: Optional[StreamState] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """spec [B,257,T,2] -> enhanced spec [B,257,T,2]; `state` continues a stream (updated in place), None = offline."""
    st = state

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v
    feat = erb_compress(features(spec), w)                              # [B,3,T,129]
    feat = tap("sfe", sfe_unfold(feat))                                 # [B,9,T,129]
    en_outs: List[torch.Tensor] = []
    x = tap("en0", conv_block(feat, w, "encoder.en_convs.0", False, False)); en_outs.append(x)
    x = tap("en1", conv_block(x, w, "encoder.en_convs.1", False, False, groups=2)); en_outs.append(x)
    for i in range(3):
        cc, th = (st.gt[i], st.tra[i]) if st is not None else (None, None)
        x, cc, th = gt_conv_block(x, w, f"encoder.en_convs.{2 + i}", False, ENC_DIL[i], cc, th)
        if st is not None:
            st.gt[i], st.tra[i] = cc, th
        tap(f"en{2 + i}", x); en_outs.append(x)
    for g in range(2):
        x, hh = dpgrnn(x, w, f"dpgrnn{g + 1}", st.inter[g] if st is not None else None, taps)
        if st is not None:
            st.inter[g] = hh
        tap(f"dp{g + 1}", x)
    for i in range(3):
        cc, th = (st.gt[3 + i], st.tra[3 + i]) if st is not None else (None, None)
        x, cc, th = gt_conv_block(x + en_outs[4 - i], w, f"decoder.de_convs.{i}", True, DEC_DIL[i], cc, th)
        if st is not None:
            st.gt[3 + i], st.tra[3 + i] = cc, th
        tap(f"de{i}", x)
    x = tap("de3", conv_block(x + en_outs[1], w, "decoder.de_convs.3", True, False, groups=2))
    m = tap("de4", conv_block(x + en_outs[0], w, "decoder.de_convs.4", True, True))
    m = erb_expand(m, w)
    sr, si = spec[..., 0].permute(0, 2, 1), spec[..., 1].permute(0, 2, 1)
    return torch.stack([sr * m[:, 0] - si * m[:, 1], si * m[:, 0] + sr * m[:, 1]], dim=-1).permute(0, 2, 1, 3)


def stream_spec(w: Weights, spec: torch.Tensor, state: Optional[StreamState] = None) -> torch.Tensor:
    B, _, T, _ = spec.shape
    st = state if state is not None else StreamState(B, spec.dtype)
    return torch.cat([forward_spec(w, spec[:, :, t:t + 1], st) for t in range(T)], dim=2)


def enhance(w: Weights, wav: torch.Tensor, streaming: bool = False) -> torch.Tensor:
    with torch.inference_mode():
        spec = stft(wav)
        return istft(stream_spec(w, spec) if streaming else forward_spec(w, spec))
