"""The unmodified caller (image_infer_v1/tools/frame_synthesizer/infer_api.py::FrameSynthesizer) around the drop-in
class.  Needs /root/reference (build container only): the scene builder and runner live in
tests/golden/make_caller_golden.py, which also wrote tests/golden/caller_batch.npz for the GPU box."""
import importlib.util
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import have_reference
from oracle import casync_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not have_reference(), reason="reference checkout not present on this machine")


def _gen():
    spec = importlib.util.spec_from_file_location("make_caller_golden", os.path.join(HERE, "golden", "make_caller_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fixture_is_what_the_unmodified_caller_produces_with_the_reference_model():
    """process_batch of the reference, reference Model on CPU: the committed fixture is reproducible bit for bit."""
    g = _gen()
    from image_infer_v1.models.unet import Model as RefModel
    with tempfile.TemporaryDirectory() as root:
        feats = g.build_scene(root)
        ckpt = os.path.join(root, "unet.pth")
        torch.save(O.make_state_dict(5, "R1"), ckpt)
        images, hub, results, calls = g.run_caller(RefModel, "cpu", root, feats, ckpt)
    gold = np.load(os.path.join(HERE, "golden", "caller_batch.npz"))
    x, a, out = calls[0]
    assert torch.equal(O.assemble_x(gold["crops"]), x)                      # the caller's float input = cat(crop, masked)/255
    assert np.array_equal(gold["audio"], a.numpy()) and np.array_equal(a.numpy(), hub)
    assert torch.equal(O.window_audio(torch.from_numpy(feats), [0, 5, g.T_FEAT - 1]), a)
    assert np.array_equal(gold["pred_u8"], np.array(out.numpy().transpose(0, 2, 3, 1) * 255, dtype=np.uint8))
    for img, res in zip(images, results):                                   # the blend really pasted something
        assert res.shape == img.shape and not np.array_equal(res, img)


def test_drop_in_class_goes_through_the_callers_constructor_and_error_path():
    """FrameSynthesizer.__init__ unmodified with calipsync_b200.Model swapped in: Model(6, "hubert").to(device),
    load_state_dict(torch.load(ckpt)), eval() (infer_api.py:41-43).  There is no CPU path, so on this GPU-less machine
    the forward raises RuntimeError -- which the caller catches and answers with the original frames
    (infer_api.py:352-357), the reference's own failure behaviour."""
    import calipsync_b200
    g = _gen()
    with tempfile.TemporaryDirectory() as root:
        feats = g.build_scene(root)
        ckpt = os.path.join(root, "unet.pth")
        sd = O.make_state_dict(5, "R1")
        torch.save(sd, ckpt)
        images, hub, results, calls = g.run_caller(calipsync_b200.Model, "cpu", root, feats, ckpt)
    assert calls == []                                                      # the forward raised before returning
    assert len(results) == len(images) and all(np.array_equal(r, i) for r, i in zip(results, images))


def test_oracle_blend_is_the_callers_blend(monkeypatch):
    """oracle.blend_paste against the reference's own paste-back: cv2.resize / cv2.bitwise_and are wrapped (recording
    their results while the UNMODIFIED process_batch runs), which yields every frame's re-sized crop and final face mask;
    the oracle's blend of those must reproduce the frames the caller returned, bit for bit."""
    import cv2
    g = _gen()
    from image_infer_v1.models.unet import Model as RefModel
    rec = {"resize": [], "and": []}
    real_resize, real_and = cv2.resize, cv2.bitwise_and

    def resize(src, dsize, *a, **k):
        out = real_resize(src, dsize, *a, **k)
        rec["resize"].append((tuple(dsize), out.copy()))
        return out

    def bitwise_and(a, b, *r, **k):
        out = real_and(a, b, *r, **k)
        rec["and"].append(out.copy())
        return out

    monkeypatch.setattr(cv2, "resize", resize)
    monkeypatch.setattr(cv2, "bitwise_and", bitwise_and)
    with tempfile.TemporaryDirectory() as root:
        feats = g.build_scene(root)
        ckpt = os.path.join(root, "unet.pth")
        torch.save(O.make_state_dict(5, "R1"), ckpt)
        images, hub, results, calls = g.run_caller(RefModel, "cpu", root, feats, ckpt)
        lms = [np.loadtxt(os.path.join(root, "positions", "%06d.txt" % i)) for i in range(g.N_FRAMES)]
    n = g.N_FRAMES
    assert len(rec["and"]) == n and len(rec["resize"]) == 2 * n        # per frame: crop -> 168x168, then back to (w, w)
    back = [r for r in rec["resize"] if r[0] != (168, 168)]
    for i in range(n):
        xmin, ymin, xmax = int(lms[i][1][0]), int(lms[i][52][1]), int(lms[i][31][0])      # crop box (infer_api.py:206-209)
        rect = (ymin, ymin + (xmax - xmin), xmin, xmax)
        got = O.blend_paste(images[i], back[i][1], rec["and"][i], rect)
        assert np.array_equal(got, results[i]), i
