"""Host-side logic: weight packer, blob layout header, module key set, C-ABI symbol table (no GPU)."""
import os
import re

import numpy as np
import pytest
import torch

import blob_model as M
import gtcrn_micro_b200 as G
from conftest import GOLDEN, ROOT
from gtcrn_micro_b200 import pack


def test_header_in_sync():
    with open(os.path.join(ROOT, "gtcrn_micro_b200", "csrc", "blob_layout.h")) as f:
        assert f.read() == pack.emit_header(), "run `python -m gtcrn_micro_b200.pack --emit-header`"


def test_blob_model_matches_fp64_oracle(oracle, weights, ref_outputs):
    tag, W = weights
    B = M.Blob(pack.pack_state_dict(W))
    W64 = oracle.cast_weights(W, torch.float64)
    x64 = torch.from_numpy(ref_outputs["wav"]).double()
    with torch.inference_mode():
        s64 = oracle.stft(x64)
        taps = {}
        y64 = oracle.forward_spec(W64, s64, taps=taps)
        w64 = oracle.istft(y64).numpy()
    so = torch.view_as_complex(s64.permute(0, 2, 1, 3).contiguous()).numpy()
    assert np.abs(M.stft(B, ref_outputs["wav"].astype(np.float64)) - so).max() < 1e-6
    mt = {}
    ym = M.forward(B, so, taps=mt)
    for k in ("en0", "en1", "en2", "en3", "en4", "de0", "de1", "de2"):
        o = taps[k].permute(0, 2, 3, 1).numpy()
        assert np.abs(mt[k] - o).max() <= 2e-6 * max(1.0, np.abs(o).max()), k
    yo = torch.view_as_complex(y64.permute(0, 2, 1, 3).contiguous()).numpy()
    assert np.abs(ym - yo).max() <= 2e-6 * max(1.0, np.abs(yo).max())
    assert np.abs(M.istft(B, yo) - w64).max() < 1e-6
    assert np.abs(M.enhance(B, ref_outputs["wav"]) - w64).max() < 1e-6


def test_module_state_dict_keys_match_checkpoint(weights_dns3):
    m = G.GTCRNMicro()
    sd = m.state_dict()
    assert set(sd) == set(weights_dns3) and len(sd) == 388
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(weights_dns3[k].shape) and v.dtype == weights_dns3[k].dtype, k
    m.load_state_dict(weights_dns3)                          # strict
    assert torch.equal(G.GTCRNMicro().state_dict()["erb.erb_fc.weight"], weights_dns3["erb.erb_fc.weight"])
    bad = dict(weights_dns3); bad.pop("sfe.depth_conv.weight")
    with pytest.raises(RuntimeError):
        G.GTCRNMicro().load_state_dict(bad)


def test_convert_to_stream_and_stream_keys_pack_identically(weights):
    _, W = weights
    m = G.GTCRNMicro(); m.load_state_dict(W)
    s = G.StreamGTCRNMicro()
    G.convert_to_stream(s, m)
    ssd = s.state_dict()
    renamed = [k for k in ssd if k not in W]
    assert len(renamed) == 28                               # 22 Conv2d + 6 ConvTranspose2d keys
    k = "decoder.de_convs.0.depth_conv.ConvTranspose2d.weight"
    ref = torch.flip(W["decoder.de_convs.0.depth_conv.weight"].permute(1, 0, 2, 3), dims=[-2, -1])
    assert torch.equal(ssd[k], ref)
    assert np.array_equal(pack.pack_state_dict(ssd), pack.pack_state_dict(W))
    assert np.array_equal(pack.pack_state_dict({"module." + a: b for a, b in W.items()}), pack.pack_state_dict(W))

    class Other(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.foo = torch.nn.Linear(2, 2)
    with pytest.raises(ValueError, match="Key error!"):
        G.convert_to_stream(Other(), m)


def test_packer_rejects_unsupported(weights_dns3):
    W = dict(weights_dns3)
    W["encoder.en_convs.0.act.weight"] = torch.ones(16)     # per-channel PReLU is not the reference architecture
    with pytest.raises(ValueError):
        pack.pack_state_dict(W)
    W = dict(weights_dns3)
    W["erb.erb_fc.weight"] = torch.rand(64, 192)            # dense ERB: not a triangular bank
    with pytest.raises(ValueError):
        pack.pack_state_dict(W)
    W = dict(weights_dns3); del W["gtcn1.blocks.0.conv1.weight"]
    with pytest.raises(KeyError):
        pack.pack_state_dict(W)


def test_no_cpu_fallback():
    m = G.GTCRNMicro()
    with pytest.raises(G.GtcrnB200Error):
        m(torch.zeros(1, 257, 3, 2))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 256, 3, 2))


def test_capi_library_exports_every_declared_symbol(b200_lib):
    from gtcrn_micro_b200 import lib
    with open(os.path.join(ROOT, "include", "gtcrn_b200.h")) as f:
        declared = sorted(set(re.findall(r"\b(gtcrnb_[a-z_0-9]+)\s*\(", f.read())))
    assert declared == sorted(lib.SYMBOLS)
    for name in declared:
        assert hasattr(b200_lib, name), name
    # calls that need no GPU
    assert b200_lib.gtcrnb_abi_version() == 1
    assert b200_lib.gtcrnb_blob_floats() == pack.blob_size()
    assert b200_lib.gtcrnb_stream_state_bytes(4) > 4 * 38112 * 4


def test_load_checkpoint_layouts(tmp_path, weights_dns3):
    # reference checkpoint dict {"epoch","optimizer","scheduler","model"} (train.py:206-211; consumer infer.py:38-39), plus the
    # variants its export scripts tolerate (stream_onnx.py:38-47): bare state_dict, "state_dict" key, DDP "module." prefix
    ref = pack.pack_state_dict(weights_dns3)
    p1 = str(tmp_path / "model_157.tar")
    torch.save({"epoch": 157, "optimizer": {}, "scheduler": {}, "model": dict(weights_dns3)}, p1)
    p2 = str(tmp_path / "bare.pt")
    torch.save(dict(weights_dns3), p2)
    p3 = str(tmp_path / "ddp.pt")
    torch.save({"state_dict": {"module." + k: v for k, v in weights_dns3.items()}}, p3)
    for path in (p1, p2, p3):
        m = G.load_checkpoint(path)
        assert isinstance(m, G.GTCRNMicro)
        assert np.array_equal(pack.pack_state_dict(m.state_dict()), ref)
    s = G.load_checkpoint(p1, stream=True)
    assert isinstance(s, G.StreamGTCRNMicro)
    assert np.array_equal(pack.pack_state_dict(s.state_dict()), ref)
