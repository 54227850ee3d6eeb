"""The drop-in boundary: class contract, state_dict layout, C-ABI exports, error behaviour (no GPU needed)."""
import ctypes
import os
import re

import pytest
import torch

import calipsync_b200
from calipsync_b200 import Model, _lib
from oracle import casync_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_spec():
    sd = Model(6, "hubert").state_dict()
    spec = O.state_spec()
    assert len(sd) == 582
    assert list(sd.keys()) == [n for n, *_ in spec]
    for n, shape, kind, _ in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
        assert sd[n].dtype == (torch.int64 if kind == "bn_nbt" else torch.float32), n


def test_seeded_init_identical_to_reference(reference_model_cls):
    torch.manual_seed(0)
    ours = Model(6, "hubert").state_dict()
    torch.manual_seed(0)
    ref = reference_model_cls(6, "hubert").state_dict()
    assert list(ours.keys()) == list(ref.keys())
    assert all(torch.equal(ours[k], ref[k]) for k in ref)


def test_load_state_dict_strict_both_ways(reference_model_cls):
    ref = reference_model_cls(6, "hubert")
    ours = Model(6, "hubert")
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict(O.make_state_dict(0, "R1"), strict=True)


def test_constructor_contract():
    Model()                      # defaults n_channels=6, mode='hubert', n_blocks=4
    Model(6, "hubert")           # positional form used by infer_api.py:41 and step2_train_unet.py:72
    with pytest.raises(NotImplementedError):
        Model(6, "wenet")
    m = Model(6, "hubert")
    assert m.n_channels == 6 and sum(p.numel() for p in m.parameters()) == 19793937


def test_forward_refuses_everything_but_the_cuda_path():
    m = Model(6, "hubert")
    x, a = O.make_inputs(1, 0)
    with pytest.raises(RuntimeError, match="inference-only"):
        m(x, a)                  # training mode
    m.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, a)                  # CPU tensors: there is no CPU path
    with pytest.raises(RuntimeError, match=r"\[B,6,160,160\]"):
        m(x[:, :3], a)
    with pytest.raises(RuntimeError, match="float32"):
        m(x.half(), a)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "casync_b200.h")).read()
    declared = set(re.findall(r"CASYNC_API[^;]*?\b(casync_\w+)\s*\(", header))
    assert declared == {n for n, _, _ in _lib.SYMBOLS}, declared ^ {n for n, _, _ in _lib.SYMBOLS}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load().casync_version()


def test_schema_and_ir_table_are_consistent():
    schema = _lib.weight_schema()
    assert len(schema) == len({n for n, _ in schema})
    irs = _lib.ir_table()
    assert len(irs) == 26 and irs[0]["name"] == "inc.inconv.0"
    names = {n for n, *_ in O.state_spec()}
    for d in irs:
        assert d["name"] + ".conv.0.weight" in names
    assert calipsync_b200.shard_sizes(1500, 8) == [188] * 7 + [184]


def test_copy_and_pickle_like_the_reference_module():
    """copy.deepcopy / torch.save(model) / pickle of the whole module work as for the reference nn.Module (EMA copies,
    whole-module checkpoints): the process-local plan handle and workspace are dropped and rebuilt lazily."""
    import copy
    import io
    import pickle
    m = Model(6, "hubert").eval()
    m._plan = (ctypes.c_void_p(0), torch.zeros(1), torch.device("cpu"))   # what a forward leaves behind (opaque handle)
    try:
        for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
            assert clone._plan is None and clone._workspace is None
            assert list(clone.state_dict().keys()) == list(m.state_dict().keys())
            assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), m.state_dict().values()))
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        again = torch.load(buf, weights_only=False)
        assert again._plan is None and not again.training
    finally:
        m._plan = None
