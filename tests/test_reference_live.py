"""Checks against the live reference; only run where /root/reference is mounted (the dev container)."""
import os

import pytest
import torch

import gtcrn_micro_b200 as G

REF = os.environ.get("GTCRN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")


def test_state_dict_key_sets_equal_reference_modules():
    from check_against_reference import import_reference
    RefModel, RefStream, ref_convert = import_reference()
    ref, mine = RefModel().state_dict(), G.GTCRNMicro().state_dict()
    assert list(ref.keys()) == list(mine.keys())
    assert all(ref[k].shape == mine[k].shape and ref[k].dtype == mine[k].dtype for k in ref)
    refs, mines = RefStream().state_dict(), G.StreamGTCRNMicro().state_dict()
    assert list(refs.keys()) == list(mines.keys())
    assert all(refs[k].shape == mines[k].shape for k in refs)
    # convert_to_stream: identical result to the reference's converter
    torch.manual_seed(0)
    m = RefModel(); s_ref = RefStream(); ref_convert(s_ref, m)
    mine_m = G.GTCRNMicro(); mine_m.load_state_dict(m.state_dict())
    s_mine = G.StreamGTCRNMicro(); G.convert_to_stream(s_mine, mine_m)
    for k, v in s_ref.state_dict().items():
        assert torch.equal(v, s_mine.state_dict()[k]), k
    # and a reference streaming module can be filled from OUR offline module (drop-in both ways)
    s_ref2 = RefStream(); G.convert_to_stream(s_ref2, mine_m)
    for k, v in s_ref.state_dict().items():
        assert torch.equal(v, s_ref2.state_dict()[k]), k
