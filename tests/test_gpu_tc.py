"""The tensor-core (tcgen05) kernels of the offline path: decoder GTConv 3x3 on fp16 hi/lo operands, GTCN pointwise convs as 3xTF32,
both with the A operand in tensor memory (gtcrn_micro_b200/csrc/gtcrn_tc.cuh), and en1 inside k_front with both operands in shared
memory (fp16 hi'/lo with a 2^8 offset, mask bit 2).  They are the default (GTCRNB_TC unset = 7); every other GPU test therefore
already exercises them.  Here they are compared with the FFMA kernels (GTCRNB_TC=0) and with the oracle
at the points where an operand split could go wrong: every layer tap, large and tiny signal levels, ragged batch sizes that leave
warpgroup rows and whole groups empty, and time segmentation.

Tolerances (BASELINE.json north_star): waveform max-abs <= 1e-4, SI-SNR delta <= 0.01 dB; activations relative 1e-4 of |ref|max.
"""
import pytest
import torch

import gtcrn_micro_b200 as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def model_with(W, monkeypatch, tc):
    """A fresh handle whose GTCRNB_TC mask is `tc` (the library reads the variable when the handle is created)."""
    monkeypatch.setenv("GTCRNB_TC", str(tc))
    m = G.GTCRNMicro()
    m.load_state_dict(W)
    m.enhance(torch.zeros(1, 1024, device=DEV))          # creates the handle under this environment
    assert m.tc_mask() == tc
    return m


def test_tc_is_the_default(weights_dns3, b200_lib, monkeypatch):
    monkeypatch.delenv("GTCRNB_TC", raising=False)
    m = G.GTCRNMicro(); m.load_state_dict(weights_dns3)
    m.enhance(torch.zeros(1, 1024, device=DEV))
    assert m.tc_mask() == 15


@pytest.mark.parametrize("tc", [1, 2, 3, 4, 8, 15])
def test_tc_layers_match_ffma_and_oracle(oracle, weights, ref_outputs, b200_lib, monkeypatch, tc):
    tag, W = weights
    spec = torch.from_numpy(ref_outputs["spec"])                     # [2,257,41,2]
    taps = {}
    with torch.inference_mode():
        y_or = oracle.forward_spec(W, spec, taps=taps)
    B, _, T, _ = spec.shape
    m_tc = model_with(W, monkeypatch, tc)
    y_tc = m_tc(spec.to(DEV))
    acts_tc = {n: m_tc.debug_activation(n, B, T).cpu() for n in ("en1", "gtcn1", "gtcn2", "de0", "de1", "de2")}
    m_ff = model_with(W, monkeypatch, 0)
    y_ff = m_ff(spec.to(DEV))
    names = {"en1": "en1", "gtcn1": "gtcn1.3", "gtcn2": "gtcn2.3", "de0": "de0", "de1": "de1", "de2": "de2"}
    rep = []
    for mine, theirs in names.items():
        o = taps[theirs].permute(0, 2, 3, 1)
        e_or = rel_err(acts_tc[mine], o)
        e_ff = rel_err(acts_tc[mine], m_ff.debug_activation(mine, B, T))
        rep.append((mine, e_or, e_ff))
        assert e_or <= 1e-4 and e_ff <= 1e-5, (tag, tc, mine, e_or, e_ff)
    print(f"[{tag} tc={tc}] " + " ".join(f"{n}:{a:.1e}/{b:.1e}" for n, a, b in rep) + f"  out vs oracle {rel_err(y_tc, y_or):.2e} vs ffma {rel_err(y_tc, y_ff):.2e}")
    assert rel_err(y_tc, y_or) <= 1e-4 and rel_err(y_tc, y_ff) <= 1e-5


@pytest.mark.parametrize("gain", [1e-4, 1.0, 30.0, 1000.0])
def test_tc_signal_level(oracle, weights_dns3, b200_lib, monkeypatch, gain):
    """fp16 hi/lo operands: tiny signals sit in the fp16 subnormal range of the lo part (absolute error stays ~1e-8), loud ones
    (30x full scale) far below the 65,504 saturation point of the point-conv activations; en1 sees spectral magnitudes that grow
    with the input level, which is what its 2^8 offset between the hi' and lo halves is for (1000x full scale still exact)."""
    x = oracle.synth_batch(3, 256 * 40 + 31, seed=5)
    x = gain * x / x.abs().max()
    m = model_with(weights_dns3, monkeypatch, 15 if gain <= 30.0 else 13)      # the decoder's fp16 operands are specified up to 30x
    y = m.enhance(x.to(DEV)).cpu()
    ref = oracle.enhance(weights_dns3, x)
    err = float((y - ref).abs().max())
    scale = max(1.0, float(ref.abs().max()))
    print(f"gain {gain}: max-abs err {err:.2e} (|ref|max {float(ref.abs().max()):.3g})")
    assert torch.isfinite(y).all() and err <= 1e-4 * scale


@pytest.mark.parametrize("B,L", [(1, 256 * 9), (2, 256 * 33 + 5), (10, 256 * 17), (13, 4000), (37, 256 * 12 + 100)])
def test_tc_ragged_batches(oracle, weights_dns3, b200_lib, monkeypatch, B, L):
    """B not a multiple of the 3 groups per warpgroup / 12 per CTA (decoder) nor of the 4 columns per warpgroup (GTCN): the padded
    rows must neither be stored nor disturb their neighbours."""
    x = oracle.synth_batch(B, L, seed=B)
    m = model_with(weights_dns3, monkeypatch, 15)
    y = m.enhance(x.to(DEV)).cpu()
    sel = sorted({0, B // 2, B - 1})
    ref = oracle.enhance(weights_dns3, x[sel])
    assert float((y[sel] - ref).abs().max()) <= 1e-4
    # bitwise independence of the batch composition
    alone = m.enhance(x[B - 1:B].to(DEV)).cpu()
    assert torch.equal(alone, y[B - 1:B])


def test_tc_time_segmentation_is_bitwise(oracle, weights_dns3, b200_lib, monkeypatch):
    """One long utterance is split into many time segments with warm-up (few utterances, many SMs); the same utterance inside a big
    batch runs unsegmented.  Row-wise MMA results do not depend on the tile they are computed in."""
    x = oracle.synth_batch(1, 256 * 600, seed=2)
    m = model_with(weights_dns3, monkeypatch, 15)
    y1 = m.enhance(x.to(DEV)).cpu()
    xb = torch.cat([x, oracle.synth_batch(299, 256 * 600, seed=3)])
    yb = m.enhance(xb.to(DEV)).cpu()
    assert torch.equal(y1, yb[:1])
    assert float((y1 - oracle.enhance(weights_dns3, x)).abs().max()) <= 1e-4
