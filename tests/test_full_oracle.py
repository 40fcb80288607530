"""configs[3] (original-GTCRN-style variant) oracle checks on CPU.  PARITY UNPINNED: there is no reference implementation of
this variant (only commented-out lines, gtcrn_micro.py:187,199,346,361,381,391,417-418,429,448,494); what is pinned here is the
functional oracle against an independent stock-torch-layer construction of the same architecture."""
import pytest
import torch

import gtcrn_full_oracle as FO
import gtcrn_oracle as O
from full_reference_model import GTCRNFullModule


@pytest.fixture(scope="module")
def W():
    return FO.random_weights(3)


def test_unpinned_state_dict_matches_stock_module_and_size(W):
    m = GTCRNFullModule().eval()
    sd = m.state_dict()
    assert set(sd.keys()) == set(W.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(W[k].shape), k
    # the public model this architecture follows is quoted at ~23.7 k parameters
    n = FO.n_learned_params(W)
    assert 23_000 <= n <= 24_500, n


def test_unpinned_oracle_equals_stock_layer_model(W):
    m = GTCRNFullModule().eval()
    m.load_state_dict(W)
    spec = O.stft(O.synth_batch(2, 256 * 40 + 17, seed=5))
    taps = {}
    with torch.inference_mode():
        want = m(spec)
        got = FO.forward_spec(W, spec, taps=taps)
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= 2e-6, err


def test_unpinned_gru_cell_equals_torch_gru():
    g = torch.Generator().manual_seed(1)
    gru = torch.nn.GRU(8, 16, batch_first=True).eval()
    w = {f"g.{k}": v.detach() for k, v in gru.state_dict().items()}
    x, h0 = torch.randn(5, 30, 8, generator=g), torch.randn(5, 16, generator=g)
    with torch.inference_mode():
        want, hl = gru(x, h0[None])
        got, hl2 = FO.gru_seq(x, h0, w, "g")
    assert float((got - want).abs().max()) <= 1e-6 and float((hl2 - hl[0]).abs().max()) <= 1e-6


def test_unpinned_stream_equals_offline_and_causal(W):
    x = O.synth_batch(2, 256 * 30, seed=9)
    spec = O.stft(x)
    with torch.inference_mode():
        off = FO.forward_spec(W, spec)
        st = FO.stream_spec(W, spec)
        spec2 = spec.clone()
        spec2[:, :, 20:] = torch.randn_like(spec2[:, :, 20:])
        off2 = FO.forward_spec(W, spec2)
    assert float((st - off).abs().max() / off.abs().max()) <= 1e-5
    assert torch.equal(off[:, :, :20], off2[:, :, :20])          # zero look-ahead
    assert not torch.equal(off[:, :, 20:], off2[:, :, 20:])
