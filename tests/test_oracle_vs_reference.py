"""Live check of the oracle restatement against the imported reference class (build container only)."""
import pytest
import torch

from oracle import casync_oracle as O


def test_spec_matches_reference_state_dict(reference_model_cls):
    ref = reference_model_cls(6, "hubert").state_dict()
    spec = O.state_spec()
    assert [n for n, *_ in spec] == list(ref.keys())
    for n, shape, *_ in spec:
        assert tuple(ref[n].shape) == tuple(shape), n


@pytest.mark.parametrize("regime", ["R0", "R1", "default"])
def test_forward_equals_reference(reference_model_cls, regime):
    torch.manual_seed(1)
    net = reference_model_cls(6, "hubert").eval()
    if regime != "default":
        net.load_state_dict(O.make_state_dict(3, regime), strict=True)
    sd = net.state_dict()
    x, a = O.make_inputs(2, 5)
    with torch.no_grad():
        ref = net(x, a)
    got = O.forward(sd, x, a)
    assert float((ref - got).abs().max()) <= 1e-6


def test_window_audio_equals_reference_arithmetic():
    """Same windows as FrameSynthesizer._get_audio_features (infer_api.py:99-145), restated inline."""
    import numpy as np
    feats = np.random.RandomState(0).randn(30, 2, 1024).astype(np.float32)
    idxs = [0, 3, 15, 25, 29]
    got = O.window_audio(torch.from_numpy(feats), idxs)
    for n, idx in enumerate(idxs):
        left, right = max(idx - 8, 0), min(idx + 8, 30)
        auds = torch.from_numpy(feats[left:right])
        if idx - 8 < 0:
            auds = torch.cat([torch.zeros_like(auds[: 8 - idx]), auds], 0)
        if idx + 8 > 30:
            auds = torch.cat([auds, torch.zeros_like(auds[: idx + 8 - 30])], 0)
        assert torch.equal(got[n], auds.reshape(32, 32, 32))
