"""Drop-in ``Model`` for the reference's ``module/unet.py::Model`` / ``image_infer_v1/models/unet.py::Model``.

Same constructor ``Model(n_channels=6, mode='hubert', n_blocks=4)`` (module/unet.py:274), same positional
``forward(x, audio_feat)`` (module/unet.py:314; call sites image_infer_v1/tools/frame_synthesizer/infer_api.py:259-260,
step2_train_unet.py:107) and the same 582-entry ``state_dict`` (names, shapes, dtypes, order), so
``Model(6, "hubert").to(device); load_state_dict(torch.load(ckpt)); eval()`` (infer_api.py:41-43) works unchanged.

The module tree below only HOLDS parameters (built from a table, in the reference's registration order so a seeded
default init is bit-identical); ``forward`` hands raw device pointers and the current CUDA stream to the sm_100a
kernels through the C ABI.  Inference only: there is no autograd, no CPU path and no PyTorch fallback -- anything the
CUDA path cannot do raises ``RuntimeError`` (which the reference's callers catch, infer_api.py:352-357).
"""
from __future__ import annotations

import ctypes
import threading

import torch
import torch.nn as nn

from . import _lib, packer

_CH = (32, 64, 128, 256, 512)


def _holder(**children):
    m = nn.Module()
    for k, v in children.items():
        m.add_module(k, v)
    return m


def _ir(inp, oup, stride):
    hid = 2 * inp
    return _holder(conv=nn.Sequential(
        nn.Conv2d(inp, hid, 1, bias=False), nn.BatchNorm2d(hid), nn.LeakyReLU(),
        nn.Conv2d(hid, hid, 3, stride, 1, groups=hid, bias=False), nn.BatchNorm2d(hid), nn.LeakyReLU(),
        nn.Conv2d(hid, oup, 1, bias=False), nn.BatchNorm2d(oup), nn.LeakyReLU()))


def _double(cin, cout, stride):
    return _holder(double_conv=nn.Sequential(_ir(cin, cout, stride), _ir(cout, cout, 1)))


def _audio_hubert():
    c = _CH
    return _holder(conv1=_ir(32, c[1], 1), conv2=_ir(c[1], c[2], 1),
                   conv3=nn.Conv2d(c[2], c[3], 3, 2, 1), bn3=nn.BatchNorm2d(c[3]), conv4=_ir(c[3], c[3], 1),
                   conv5=nn.Conv2d(c[3], c[4], 3, 2, 3), bn5=nn.BatchNorm2d(c[4]),
                   conv6=_ir(c[4], c[4], 1), conv7=_ir(c[4], c[4], 1), bn7=nn.BatchNorm2d(c[4]))


class _Gamma(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query_conv = nn.Conv2d(c, c // 8, 1)
        self.key_conv = nn.Conv2d(c, c // 8, 1)
        self.value_conv = nn.Conv2d(c, c, 1)
        self.gamma = nn.Parameter(torch.zeros(1))


def _attention(c, c2):
    return _holder(cross_attention=_Gamma(c), attention_adjust_p_1=nn.Conv2d(c2, c, 1),
                   attention_adjust_b_1=nn.Conv2d(c, c2, 1), bn=nn.BatchNorm2d(c2))


class Model(nn.Module):
    """B200-native CASync generator (inference).  See module docstring."""

    def __init__(self, n_channels=6, mode="hubert", n_blocks=4):
        super().__init__()
        if mode != "hubert":
            raise NotImplementedError("calipsync_b200.Model implements mode='hubert' only (no reference caller uses "
                                      "%r; image_infer_v1/tools/frame_synthesizer/infer_api.py:41)" % (mode,))
        if n_channels != 6 or n_blocks != 4:
            raise NotImplementedError("kernels are specialised for n_channels=6, n_blocks=4 (module/unet.py:274-277)")
        self.n_channels = n_channels
        c = _CH
        # registration order of module/unet.py:281-311
        self.audio_model = _audio_hubert()
        self.fuse_conv = nn.Sequential(_double(c[4] * 2, c[4], 1), _double(c[4], c[3], 1))
        self.inc = _holder(inconv=nn.Sequential(_ir(n_channels, c[0], 1)))
        for i in range(4):
            setattr(self, "down%d" % (i + 1), _holder(maxpool_conv=nn.Sequential(_double(c[i], c[i + 1], 2))))
        for i, (cin, cout) in enumerate(((c[4], c[3] // 2), (c[3], c[2] // 2), (c[2], c[1] // 2), (c[1], c[0]))):
            setattr(self, "up%d" % (i + 1), _holder(conv=_double(cin, cout, 1)))
        self.outc = _holder(conv=nn.Conv2d(c[0], 3, 1))
        self.outc_bn = nn.BatchNorm2d(3)
        self.mlp_fusion = _holder(fc1=nn.Linear(c[4] * 2, c[4] * 2), bn1=nn.BatchNorm1d(c[4] * 2),
                                  fc2=nn.Linear(c[4] * 2, c[4] * 2), bn2=nn.BatchNorm1d(c[4] * 2))
        self.attention_blocks = nn.ModuleList([_attention(c[4], c[4] * 2) for _ in range(n_blocks)])
        self.bn_kx = nn.BatchNorm2d(c[4] * 2)
        self.bn_tx = nn.BatchNorm2d(c[4] * 2)
        self._plan = None          # (handle, device blob, device)
        self._workspace = None
        self._lock = threading.RLock()   # one forward at a time per Model: plan, lanes and workspace are shared state
        self._last = None          # (stream id, event) of the last forward: orders forwards issued on different streams
        self._frame_io = {}        # forward_frames: model-owned (x, audio) staging per (device, batch): stable addresses

    # ---- packed-weight lifecycle --------------------------------------------------------------------------------
    def _invalidate(self):
        plan = self.__dict__.get("_plan")
        if plan is not None:
            with torch.cuda.device(plan[2]):   # destroy synchronises the plan's own device, not the current one
                _lib.load().casync_plan_destroy(plan[0])
        self._plan = None
        self._workspace = None
        self._last = None
        self.__dict__["_frame_io"] = {}
        self.__dict__.pop("_clip_out", None)

    # The plan handle, the workspace and the lock are process-local (ctypes pointers cannot be pickled): copies and
    # pickles carry the parameters only and rebuild the packed weights lazily, like the reference nn.Module would.
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_plan"] = None
        d["_workspace"] = None
        d["_last"] = None
        d["_frame_io"] = {}
        d.pop("_clip_out", None)
        d.pop("_lock", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._lock = threading.RLock()

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._lock = threading.RLock()
        return new

    def _apply(self, fn, *a, **k):           # .to() / .cuda() / .half() ...
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def repack(self):
        """Call after modifying parameters in place (packed weights are otherwise cached)."""
        self._invalidate()

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _ensure_plan(self, device):
        if self._plan is not None and self._plan[2] == device:
            return self._plan[0]
        self._invalidate()
        lib = _lib.load()
        cap = torch.cuda.get_device_capability(device)
        if cap[0] != 10:
            raise RuntimeError("calipsync_b200 needs a compute-capability 10.x GPU (B200, sm_100a); %s is %d.%d -- "
                               "no fallback path exists" % (device, cap[0], cap[1]))
        blob, offsets = packer.pack(self.state_dict())
        dev_blob = blob.to(device)
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = lib.casync_plan_create(blob.data_ptr(), dev_blob.data_ptr(), blob.numel(),
                                        offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(offsets),
                                        ctypes.byref(handle))
        _lib.check(rc, "casync_plan_create")
        self._plan = (handle, dev_blob, device)
        return handle

    def _ensure_workspace(self, handle, batch, device):
        need = _lib.load().casync_workspace_bytes(handle, batch)
        ws = self._workspace
        if ws is None or ws.numel() < need or ws.device != device:
            self._workspace = ws = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    # ---- forward -------------------------------------------------------------------------------------------------
    def _check_inputs(self, x, audio_feat):
        if self.training:
            raise RuntimeError("calipsync_b200.Model is inference-only: call .eval() first (training-mode BatchNorm "
                               "and autograd are not implemented; step2_train_unet.py is out of scope)")
        if not (torch.is_tensor(x) and torch.is_tensor(audio_feat)):
            raise RuntimeError("forward(x, audio_feat) expects tensors")
        if x.dim() != 4 or tuple(x.shape[1:]) != (6, 160, 160):
            raise RuntimeError("x must be [B,6,160,160], got %s" % (tuple(x.shape),))
        if audio_feat.dim() != 4 or tuple(audio_feat.shape[1:]) != (32, 32, 32) or audio_feat.shape[0] != x.shape[0]:
            raise RuntimeError("audio_feat must be [B,32,32,32] with B=%d, got %s" % (x.shape[0], tuple(audio_feat.shape)))
        if x.dtype != torch.float32 or audio_feat.dtype != torch.float32:
            raise RuntimeError("inputs must be float32 (got %s, %s)" % (x.dtype, audio_feat.dtype))
        if not x.is_cuda or audio_feat.device != x.device:
            raise RuntimeError("inputs must live on one CUDA device (got %s, %s); there is no CPU path"
                               % (x.device, audio_feat.device))
        p = next(self.parameters())
        if p.device != x.device:
            raise RuntimeError("model is on %s but inputs are on %s" % (p.device, x.device))
        if x.shape[0] == 0:
            raise RuntimeError("empty batch")

    def _run(self, x, audio_feat, out, flags):
        device = x.device
        x, audio_feat = x.contiguous(), audio_feat.contiguous()
        # One forward at a time per Model (the reference nn.Module may be called from several threads / streams): the
        # plan's lanes, events, graph cache and the workspace are shared.  Threads serialise on the lock; a forward
        # issued on another CUDA stream than the previous one first waits for that one's completion event.
        with self._lock, torch.cuda.device(device):
            handle = self._ensure_plan(device)
            cur = torch.cuda.current_stream(device)
            last = self._last
            if last is not None and last[0] != cur.cuda_stream:
                cur.wait_event(last[1])
            ws = self._ensure_workspace(handle, x.shape[0], device)
            ws.record_stream(cur)
            rc = _lib.load().casync_forward(handle, x.data_ptr(), audio_feat.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                            x.shape[0], flags, ctypes.c_void_p(cur.cuda_stream))
            _lib.check(rc, "casync_forward")
            ev = last[1] if last is not None else torch.cuda.Event()
            ev.record(cur)
            self._last = (cur.cuda_stream, ev)
        return out

    @torch.no_grad()
    def forward(self, x, audio_feat, *extra):
        """x: fp32 [B,6,160,160] = cat([face, masked face]); audio_feat: fp32 [B,32,32,32] -> fp32 [B,3,160,160].
        Convenience: forward(face[B,3,..], masked[B,3,..], audio_feat) concatenates the two image halves."""
        if extra:
            if len(extra) != 1:
                raise RuntimeError("forward takes (x, audio_feat) or (face, masked_face, audio_feat)")
            x, audio_feat = torch.cat([x, audio_feat], dim=1), extra[0]
        self._check_inputs(x, audio_feat)
        out = torch.empty(x.shape[0], 3, 160, 160, dtype=torch.float32, device=x.device)
        return self._run(x, audio_feat, out, _lib.F_BF16)

    @torch.no_grad()
    def forward_uint8(self, x, audio_feat):
        """Same forward, but emits what the caller computes next (infer_api.py:265-266): uint8 [B,160,160,3] =
        floor(pred * 255) in HWC order, ready for `crop_img[4:164, 4:164] = pred` -- one batched D2H instead of B."""
        self._check_inputs(x, audio_feat)
        out = torch.empty(x.shape[0], 160, 160, 3, dtype=torch.uint8, device=x.device)
        return self._run(x, audio_feat, out, _lib.F_OUT_U8_HWC)

    # ---- caller-side input assembly on the device ---------------------------------------------
    @torch.no_grad()
    def prepare_inputs(self, crops_u8, hubert_feats, frame_idx, _into=None):
        """What FrameSynthesizer.process_batch builds with numpy per frame (infer_api.py:99-145, 238-245), on the GPU:
        crops_u8 uint8 [B,160,160,3] (= ``crop_img[4:164, 4:164]``), hubert_feats fp32 [T,2,1024] (device resident),
        frame_idx int [B]  ->  (x fp32 [B,6,160,160], audio_feat fp32 [B,32,32,32]), bit-identical to the caller's."""
        if not (crops_u8.is_cuda and hubert_feats.is_cuda and frame_idx.is_cuda):
            raise RuntimeError("prepare_inputs needs CUDA tensors (there is no CPU path)")
        if crops_u8.dtype != torch.uint8 or crops_u8.dim() != 4 or tuple(crops_u8.shape[1:]) != (160, 160, 3):
            raise RuntimeError("crops must be uint8 [B,160,160,3], got %s %s" % (crops_u8.dtype, tuple(crops_u8.shape)))
        if hubert_feats.dtype != torch.float32 or hubert_feats.dim() != 3 or tuple(hubert_feats.shape[1:]) != (2, 1024):
            raise RuntimeError("hubert features must be float32 [T,2,1024], got %s %s"
                               % (hubert_feats.dtype, tuple(hubert_feats.shape)))
        b = crops_u8.shape[0]
        if frame_idx.numel() != b or b == 0:
            raise RuntimeError("need one frame index per crop")
        dev = crops_u8.device
        crops_u8, hubert_feats = crops_u8.contiguous(), hubert_feats.contiguous()
        idx = frame_idx.to(torch.int32).contiguous()
        if _into is not None:
            x, a = _into
        else:
            x = torch.empty(b, 6, 160, 160, dtype=torch.float32, device=dev)
            a = torch.empty(b, 32, 32, 32, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _lib.load().casync_prepare_inputs(crops_u8.data_ptr(), hubert_feats.data_ptr(), hubert_feats.shape[0],
                                                   idx.data_ptr(), x.data_ptr(), a.data_ptr(), b, ctypes.c_void_p(stream))
        _lib.check(rc, "casync_prepare_inputs")
        return x, a

    @torch.no_grad()
    def forward_frames(self, crops_u8, hubert_feats, frame_idx, out=None):
        """crops + HuBERT features + frame indices -> uint8 [B,160,160,3] = floor(pred*255) in the caller's HWC layout:
        prepare_inputs -> forward -> uint8 epilogue without leaving the device."""
        dev, b = crops_u8.device, int(crops_u8.shape[0])
        # The assembled (x, audio) tensors live in model-owned buffers, one pair per batch size: their addresses are part
        # of the CUDA-graph key, and fresh torch.empty() blocks move with the allocator's state (measured: a clip in
        # batches of 64 on 4 GPUs at 25.8 instead of 11.7 ms, every call re-capturing).  Reuse is ordered like the
        # workspace: same stream by stream order, another stream by the completion event of the last forward.
        with self._lock, torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            last = self._last
            if last is not None and last[0] != cur.cuda_stream:
                cur.wait_event(last[1])
            io = self._frame_io.get((dev, b))
            if io is None:
                if len(self._frame_io) >= 4:
                    self._frame_io.pop(next(iter(self._frame_io)))
                io = (torch.empty(b, 6, 160, 160, dtype=torch.float32, device=dev),
                      torch.empty(b, 32, 32, 32, dtype=torch.float32, device=dev))
                self._frame_io[(dev, b)] = io
            for t in io:
                t.record_stream(cur)
            x, a = self.prepare_inputs(crops_u8, hubert_feats, frame_idx, _into=io)
            if out is None:
                out = torch.empty(b, 160, 160, 3, dtype=torch.uint8, device=dev)
            return self._run(x, a, out, _lib.F_OUT_U8_HWC)

    # ---- introspection for tests / profiling --------------------------------------------------------------------
    def stage(self, name, batch):
        """bf16 [rows, cols] view of a stage activation left in the workspace by the last forward (batch <= chunk)."""
        handle, ws = self._plan[0], self._workspace
        off, rows, cols, ld = ctypes.c_size_t(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.load().casync_stage_view(handle, batch, name.encode(), ctypes.byref(off), ctypes.byref(rows),
                                                 ctypes.byref(cols), ctypes.byref(ld)), "casync_stage_view")
        flat = ws[off.value: off.value + rows.value * ld.value * 2].view(torch.bfloat16)
        return torch.as_strided(flat, (rows.value, cols.value), (ld.value, 1))

    @torch.no_grad()
    def profile(self, x, audio_feat):
        """One forward with a CUDA event after every launch -> list of dicts(name, ms, flops, bytes)."""
        self._check_inputs(x, audio_feat)
        out = torch.empty(x.shape[0], 3, 160, 160, dtype=torch.float32, device=x.device)
        handle = self._ensure_plan(x.device)
        ws = self._ensure_workspace(handle, x.shape[0], x.device)
        recs = (_lib.LaunchRecord * 1024)()
        n = ctypes.c_int()
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = _lib.load().casync_forward_profiled(handle, x.contiguous().data_ptr(), audio_feat.contiguous().data_ptr(),
                                                     out.data_ptr(), ws.data_ptr(), x.shape[0], _lib.F_BF16,
                                                     ctypes.c_void_p(stream), recs, 1024, ctypes.byref(n))
        _lib.check(rc, "casync_forward_profiled")
        return [dict(name=recs[i].name.decode(), ms=recs[i].ms, flops=recs[i].flops, bytes=recs[i].bytes)
                for i in range(n.value)]

    def launches_per_forward(self, batch):
        return int(_lib.load().casync_launches_per_forward(self._plan[0], batch)) if self._plan else 0

    def graph_replays(self):
        """Forwards replayed from a cached CUDA graph so far (same tensors / batch seen before; CASYNC_GRAPH=0 disables)."""
        return int(_lib.load().casync_graph_replays(self._plan[0])) if self._plan else 0
