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


def test_input_assembly_equals_reference_code():
    """oracle.assemble_x / window_audio against the reference's own code: FrameSynthesizer._get_audio_features called
    unbound, and the image lines of process_batch (infer_api.py:238-245) executed verbatim with cv2."""
    import os
    import sys

    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from conftest import REFERENCE_DIR, have_reference
    if not have_reference():
        pytest.skip("reference checkout not present on this machine")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    from image_infer_v1.tools.frame_synthesizer.infer_api import FrameSynthesizer

    rs = np.random.RandomState(3)
    feats = rs.randn(40, 2, 1024).astype(np.float32)
    idxs = [0, 1, 7, 8, 20, 32, 33, 39, 45, 60]
    ref = FrameSynthesizer._get_audio_features(None, feats, idxs)
    assert torch.equal(O.window_audio(torch.from_numpy(feats), idxs), torch.from_numpy(ref))

    crops168 = rs.randint(0, 256, size=(3, 168, 168, 3), dtype=np.uint8)
    want = []
    for crop_img in crops168:                      # infer_api.py:238-245
        img_real_ex = crop_img[4:164, 4:164].copy()
        img_masked = cv2.rectangle(img_real_ex.copy(), (5, 5, 150, 145), (0, 0, 0), -1)
        img_real_ex = img_real_ex.transpose(2, 0, 1).astype(np.float32) / 255.0
        img_masked = img_masked.transpose(2, 0, 1).astype(np.float32) / 255.0
        want.append(np.concatenate([img_real_ex, img_masked]))
    got = O.assemble_x(crops168[:, 4:164, 4:164])
    assert torch.equal(got, torch.from_numpy(np.stack(want)))


def test_hubert_extraction_restatement_equals_reference():
    """BASELINE config 4's upstream stage: tools/e2e_hubert.extract_features restates HubertExtractor.extract_features
    (utils/hubert_extractor.py:18-58: 1000-step clips, pad / trim to expected_T, drop an odd frame, [-1, 2, 1024]).
    Checked live against the reference method on a small random HuBERT (the pretrained weights are not on disk)."""
    import importlib.util
    import sys
    import types
    from conftest import REFERENCE_DIR, ROOT, have_reference
    if not have_reference():
        pytest.skip("reference checkout not present on this machine")
    transformers = pytest.importorskip("transformers")
    fe_cls, hubert_cls, cfg_cls = transformers.Wav2Vec2FeatureExtractor, transformers.HubertModel, transformers.HubertConfig
    _ = transformers.Wav2Vec2Processor                                          # resolve the lazy imports first
    if "soundfile" not in sys.modules:                                          # imported at the reference module's top,
        import importlib.machinery                                              # never used here
        stub = types.ModuleType("soundfile")
        stub.__spec__ = importlib.machinery.ModuleSpec("soundfile", None)
        sys.modules["soundfile"] = stub
    spec = importlib.util.spec_from_file_location("ref_hubert_extractor", REFERENCE_DIR + "/utils/hubert_extractor.py")
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    sys.modules.setdefault("bench", types.ModuleType("bench"))                  # e2e_hubert only needs bench in main()
    spec2 = importlib.util.spec_from_file_location("e2e_hubert", ROOT + "/tools/e2e_hubert.py")
    ours = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(ours)

    cfg = cfg_cls(hidden_size=1024, num_hidden_layers=1, num_attention_heads=16, intermediate_size=64,
                                    feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True)
    torch.manual_seed(0)
    model = hubert_cls(cfg).eval()
    ext = object.__new__(ref_mod.HubertExtractor)       # no checkpoint to load: wire the members by hand
    ext.device = "cpu"
    ext.model = model
    ext.processor = fe_cls(feature_size=1, sampling_rate=16000, padding_value=0.0,
                                                          do_normalize=True, return_attention_mask=False)
    for seconds in (3.0, 20.5, 41.03):                   # below one clip, one clip + remainder, two clips + short tail
        speech = (torch.randn(int(seconds * 16000), generator=torch.Generator().manual_seed(int(seconds))) * 0.1)
        want = ext.extract_features(speech.numpy())
        got = ours.extract_features(model, speech)
        assert tuple(got.shape) == tuple(want.shape)
        assert float((got - want).abs().max()) <= 2e-4 * max(1.0, float(want.abs().max()))
    ext.model = ext.processor = None
