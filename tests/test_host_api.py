"""Host-side argument handling of the Python surface (no GPU needed): the checks that guard raw pointers handed to the C ABI,
the `-C` config reader of the file pipeline, the checkpoint loader's safe default."""
import os
import textwrap

import pytest
import torch

import gtcrn_micro_b200 as G
from gtcrn_micro_b200 import infer, lib, modules

REF_CFG = "/root/reference/gtcrn_micro/conf/cfg_infer.yaml"


def test_enhance_host_rejects_bad_out_before_touching_the_library():
    m = G.GTCRNMicro()
    wav = torch.zeros(2, 1000)
    for bad in (torch.zeros(2, 768, dtype=torch.float64), torch.zeros(2, 767), torch.zeros(3, 768), torch.zeros(768, 2).t()):
        with pytest.raises(ValueError):
            m.enhance_host(wav, bad)
    wav16 = torch.zeros(2, 1000, dtype=torch.int16)
    for bad in (torch.zeros(2, 768), torch.zeros(2, 700, dtype=torch.int16), torch.zeros(768, 2, dtype=torch.int16).t()):
        with pytest.raises(ValueError):
            m.enhance_host_pcm16(wav16, bad)


def test_pinned_empty_failure_raises_the_documented_error(monkeypatch):
    class FakeLib:
        @staticmethod
        def gtcrnb_host_alloc(n):
            return None

        gtcrnb_host_alloc_wc = gtcrnb_host_alloc

    monkeypatch.setattr(modules._lib, "load", lambda: FakeLib)
    with pytest.raises(lib.GtcrnB200Error):
        G.pinned_empty((4, 4))
    with pytest.raises(lib.GtcrnB200Error):
        G.pinned_empty((4, 4), write_combined=True)


def test_config_interpolation(tmp_path):
    p = tmp_path / "cfg.yaml"
    p.write_text(textwrap.dedent("""\
        ---
        test_dataset:
          noisy_dir: data/noisy
          clean_dir: data/clean
        network:
          exp_path: data/exp_1
          ckpt_name: best_model_157
          checkpoint: ${network.exp_path}/checkpoints/${network.ckpt_name}.tar
          enh_folder: ${network.exp_path}/${network.ckpt_name}/enhanced
          alias: ${network.checkpoint}
          n: 3
          n_alias: ${network.n}
        """))
    cfg = infer.load_config(str(p))
    assert cfg["network"]["checkpoint"] == "data/exp_1/checkpoints/best_model_157.tar"
    assert cfg["network"]["enh_folder"] == "data/exp_1/best_model_157/enhanced"
    assert cfg["network"]["alias"] == cfg["network"]["checkpoint"]
    assert cfg["network"]["n_alias"] == 3
    p.write_text("a: ${b}\nb: ${a}\n")
    with pytest.raises(ValueError):
        infer.load_config(str(p))
    p.write_text("a: ${nope.x}\n")
    with pytest.raises(KeyError):
        infer.load_config(str(p))


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not mounted")
def test_reference_infer_config_resolves():
    cfg = infer.load_config(REF_CFG)
    exp = cfg["network"]["exp_path"]
    assert cfg["network"]["checkpoint"] == f"{exp}/checkpoints/best_model_157.tar"
    assert cfg["network"]["enh_folder"] == f"{exp}/best_model_157/enhanced"
    assert "${" not in repr(cfg)


@pytest.mark.skipif(not os.path.exists("/root/reference/gtcrn_micro/ckpts/best_model_dns3.tar"), reason="reference tree not mounted")
def test_load_checkpoint_uses_safe_unpickling(monkeypatch):
    seen = {}
    real = torch.load

    def spy(*a, **k):
        seen.update(k)
        return real(*a, **k)

    monkeypatch.setattr(torch, "load", spy)
    m = G.load_checkpoint("/root/reference/gtcrn_micro/ckpts/best_model_dns3.tar")
    assert seen.get("weights_only") is True
    assert isinstance(m, G.GTCRNMicro)
