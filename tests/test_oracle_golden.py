"""oracle/ vs the golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import casync_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def probe(t):
    flat = t.reshape(-1)
    return flat[torch.linspace(0, flat.numel() - 1, 256).long()].numpy()


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_oracle_reproduces_reference_output(regime):
    g = np.load(os.path.join(GOLD, "unet_%s_seed0_b2.npz" % regime))
    sd = O.make_state_dict(0, regime)
    x, a = O.make_inputs(2, 0)
    out, st = O.forward(sd, x, a, return_stages=True)
    # same torch ops on the same machine class: agreement to fp32 round-off
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-6)
    for name in O.STAGE_NAMES:
        if "probe_" + name not in g:
            continue
        ref = g["probe_" + name]
        tol = 1e-5 * max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(probe(st[name]), ref, rtol=0, atol=tol, err_msg=name)
        assert abs(float(st[name].double().norm()) - float(g["norm_" + name])) <= 1e-5 * float(g["norm_" + name]) + 1e-9


def test_inputs_follow_reference_layout():
    x, a = O.make_inputs(3, 7)
    assert x.shape == (3, 6, 160, 160) and a.shape == (3, 32, 32, 32)
    assert float(x.min()) >= 0 and float(x.max()) <= 1
    assert float(x[:, 3:6, 5:150, 5:155].abs().max()) == 0          # cv2.rectangle((5,5,150,145)) mask
    assert float(x[:, 3:6, 4, :].abs().max()) > 0 and float(x[:, 3:6, 150, :].abs().max()) > 0
    assert float(x[:, 3:6, :, 155].abs().max()) > 0
    # frame i depends only on (seed, index): shard-local generation equals global generation
    x2, a2 = O.make_inputs(1, 7, frame_offset=2)
    assert torch.equal(x2[0], x[2]) and torch.equal(a2[0], a[2])


def test_window_audio_zero_pads_clip_ends():
    feats = torch.arange(20 * 2 * 1024, dtype=torch.float32).reshape(20, 2, 1024) + 1
    w = O.window_audio(feats, [0, 10, 19]).reshape(3, 16, 2, 1024)
    assert float(w[0, :8].abs().max()) == 0 and torch.equal(w[0, 8:], feats[0:8])
    assert torch.equal(w[1], feats[2:18])
    assert torch.equal(w[2, :9], feats[11:20]) and float(w[2, 9:].abs().max()) == 0


def test_metrics():
    a = torch.zeros(4)
    b = torch.full((4,), 1 / 255)
    assert abs(O.max_abs_255(a, b) - 1) < 1e-6
    assert abs(O.psnr_db(a, b) - 20 * np.log10(255)) < 1e-6
