"""nn.Module surface of the ORIGINAL-GTCRN-style variant (BASELINE configs[3]): unfold SFE, grouped en1 / de3, dilated depthwise
GTConv blocks, GRU-based TRA and two DPGRNNs instead of SFE-Lite / TRALite / GTCN.

PARITY UNPINNED.  The reference (bglid/GTCRN-Micro) removed this architecture and only keeps it as commented-out lines --
gtcrn_micro/models/gtcrn_micro.py:187 (point_conv1 on 3x the channels), :199,210 (depthwise groups), :346 (en0 on 3*3 channels),
:361,448 (groups=2), :381,391 (dilations 2 and 5), :417-418,429 (decoder dilations / padding), :494 (SFE(3, 1)), and :5 /
gtcrn_micro/README.md:4 (GRU blocks + DPGRNN swapped for the grouped TCN) -- so there is no reference module, checkpoint or test
to mirror.  The state_dict below uses the key names of the public model the reference derives from, the
kernels are checked against the builder-written oracle/gtcrn_full_oracle.py, and weights are seeded random unless loaded.

Same call surface as GTCRNMicro / StreamGTCRNMicro (forward on spectra, enhance, enhance_host, stft / istft, pair_metrics, native
streaming state with step / step_wav / flush_wav); the C ABI handle is created with GTCRNB_VARIANT_FULL.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as _lib
from . import pack_full as _pack_full
from .modules import GTCRNMicro, StreamGTCRNMicro, _bn


def _gru(p, n_in, n_h, bidirectional=False):
    out = []
    for suf in ("", "_reverse") if bidirectional else ("",):
        out += [(f"{p}.weight_ih_l0{suf}", (3 * n_h, n_in), "w"), (f"{p}.weight_hh_l0{suf}", (3 * n_h, n_h), "w"),
                (f"{p}.bias_ih_l0{suf}", (3 * n_h,), "b"), (f"{p}.bias_hh_l0{suf}", (3 * n_h,), "b")]
    return out


def _gt_block(p, deconv):
    o = [(f"{p}.point_conv1.weight", (24, 16, 1, 1) if deconv else (16, 24, 1, 1), "w"), (f"{p}.point_conv1.bias", (16,), "b")] + _bn(f"{p}.point_bn1", 16)
    o += [(f"{p}.point_act.weight", (1,), "prelu"), (f"{p}.depth_conv.weight", (16, 1, 3, 3), "w"), (f"{p}.depth_conv.bias", (16,), "b")] + _bn(f"{p}.depth_bn", 16)
    o += [(f"{p}.depth_act.weight", (1,), "prelu"), (f"{p}.point_conv2.weight", (16, 8, 1, 1) if deconv else (8, 16, 1, 1), "w"), (f"{p}.point_conv2.bias", (8,), "b")]
    o += _bn(f"{p}.point_bn2", 8) + _gru(f"{p}.tra.att_gru", 8, 16) + [(f"{p}.tra.att_fc.weight", (8, 16), "w"), (f"{p}.tra.att_fc.bias", (8,), "b")]
    return o


def param_spec_full():
    s = [("erb.erb_fc.weight", (64, 192), "erb"), ("erb.ierb_fc.weight", (192, 64), "erbT")]
    s += [("encoder.en_convs.0.conv.weight", (16, 9, 1, 5), "w"), ("encoder.en_convs.0.conv.bias", (16,), "b")] + _bn("encoder.en_convs.0.bn", 16) + [("encoder.en_convs.0.act.weight", (1,), "prelu")]
    s += [("encoder.en_convs.1.conv.weight", (16, 8, 1, 5), "w"), ("encoder.en_convs.1.conv.bias", (16,), "b")] + _bn("encoder.en_convs.1.bn", 16) + [("encoder.en_convs.1.act.weight", (1,), "prelu")]
    for i in (2, 3, 4):
        s += _gt_block(f"encoder.en_convs.{i}", False)
    for g in (1, 2):
        p = f"dpgrnn{g}"
        for r in (1, 2):
            s += _gru(f"{p}.intra_rnn.rnn{r}", 8, 4, bidirectional=True)
        s += [(f"{p}.intra_fc.weight", (16, 16), "w"), (f"{p}.intra_fc.bias", (16,), "b"), (f"{p}.intra_ln.weight", (33, 16), "ln_w"), (f"{p}.intra_ln.bias", (33, 16), "ln_b")]
        for r in (1, 2):
            s += _gru(f"{p}.inter_rnn.rnn{r}", 8, 8)
        s += [(f"{p}.inter_fc.weight", (16, 16), "w"), (f"{p}.inter_fc.bias", (16,), "b"), (f"{p}.inter_ln.weight", (33, 16), "ln_w"), (f"{p}.inter_ln.bias", (33, 16), "ln_b")]
    for i in (0, 1, 2):
        s += _gt_block(f"decoder.de_convs.{i}", True)
    s += [("decoder.de_convs.3.conv.weight", (16, 8, 1, 5), "w"), ("decoder.de_convs.3.conv.bias", (16,), "b")] + _bn("decoder.de_convs.3.bn", 16) + [("decoder.de_convs.3.act.weight", (1,), "prelu")]
    s += [("decoder.de_convs.4.conv.weight", (16, 2, 1, 5), "w"), ("decoder.de_convs.4.conv.bias", (2,), "b")] + _bn("decoder.de_convs.4.bn", 2)
    return s


class _FullMixin:
    _variant = 1

    @classmethod
    def _param_spec(cls):
        return param_spec_full()

    def _pack(self) -> np.ndarray:
        return _pack_full.pack_state_dict(self.state_dict())


class GTCRN(_FullMixin, GTCRNMicro):
    """Offline model of the original-GTCRN-style variant (parity unpinned, see the module docstring).
    Activations for `debug_activation`: en0 (F=65), en1..en4, dp1, dp2, de0..de2."""


GTCRNFull = GTCRN


class StreamGTCRN(_FullMixin, StreamGTCRNMicro):
    """Frame-by-frame twin of `GTCRN`: native ring-buffer / GRU state for N concurrent streams (new_state, step, step_wav,
    flush_wav, reset).  The state per stream: 32 rows of GTConv history, 6 TRA GRU states, 2 x 33 inter-GRU states."""

    def forward(self, *args, **kwargs):
        raise _lib.GtcrnB200Error("StreamGTCRN: the reference defines no cache-tensor layout for this variant; use new_state() / step() / step_wav()")


StreamGTCRNFull = StreamGTCRN
