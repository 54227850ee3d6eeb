"""Host-buffer pipeline around ``Model.forward``: what ``FrameSynthesizer.process_batch`` does per batch
(image_infer_v1/tools/frame_synthesizer/infer_api.py:256-266: ``torch.from_numpy(...).to(device)`` x2, the model
call, then ``predictions[i].cpu()`` per frame), restructured so that the three legs of consecutive batches overlap:

    copy-in stream    H2D of batch i+1 (pinned host -> one of `depth` device input slots)
    compute stream    forward of batch i    (the drop-in Model, same kernels, same results)
    copy-out stream   D2H of batch i-1      (device output slot -> caller's pinned host buffer)

Nothing here changes what is computed: every batch goes through ``Model._run`` with fp32 NCHW inputs (or, with
``uint8=True``, through the uint8 HWC epilogue of ``Model.forward_uint8``).  Events order the slots, so a slot is
never overwritten before its consumer is done.

``frames=True`` moves the caller's per-frame numpy work onto the device as well (``Model.forward_frames``): the clip's
HuBERT features are uploaded once (``set_features``), a batch is then just uint8 crops [n,160,160,3] + frame indices
in and uint8 frames out -- 8x less H2D and 4x less D2H than the fp32 tensors of infer_api.py:256-266.
"""
from __future__ import annotations

import torch

from . import _lib


class HostPipeline:
    def __init__(self, model, batch: int, uint8: bool = False, depth: int = 2, device=None, frames: bool = False):
        uint8 = uint8 or frames
        self.model, self.batch, self.uint8, self.depth = model, int(batch), bool(uint8), int(depth)
        self.frames, self.feats, self.ev_feats = bool(frames), None, None
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA device (there is no CPU path)")
        d = self.device
        self.s_in, self.s_out = torch.cuda.Stream(d), torch.cuda.Stream(d)
        if frames:   # slots hold uint8 crops and int32 frame indices
            self.x = [torch.empty(batch, 160, 160, 3, dtype=torch.uint8, device=d) for _ in range(depth)]
            self.a = [torch.empty(batch, dtype=torch.int32, device=d) for _ in range(depth)]
        else:
            self.x = [torch.empty(batch, 6, 160, 160, dtype=torch.float32, device=d) for _ in range(depth)]
            self.a = [torch.empty(batch, 32, 32, 32, dtype=torch.float32, device=d) for _ in range(depth)]
        if uint8:
            self.o = [torch.empty(batch, 160, 160, 3, dtype=torch.uint8, device=d) for _ in range(depth)]
        else:
            self.o = [torch.empty(batch, 3, 160, 160, dtype=torch.float32, device=d) for _ in range(depth)]
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]          # noqa: E731
        self.ev_in, self.ev_consumed, self.ev_out, self.ev_drained = mk(), mk(), mk(), mk()
        self.i = 0

    def set_features(self, host_feats: torch.Tensor):
        """frames mode: upload the clip's HuBERT features [T,2,1024] once; later batches only carry frame indices."""
        with torch.cuda.stream(self.s_in):
            self.feats = host_feats.to(self.device, non_blocking=True)
            self.ev_feats = torch.cuda.Event()
            self.ev_feats.record(self.s_in)
        compute = torch.cuda.current_stream(self.device)
        compute.wait_event(self.ev_feats)
        self.feats.record_stream(compute)

    def submit(self, host_x: torch.Tensor, host_audio: torch.Tensor, host_out: torch.Tensor):
        """Enqueue one batch (pinned host tensors in, pinned host tensor out); returns immediately.
        fp32 mode: (x [n,6,160,160], audio_feat [n,32,32,32]); frames mode: (crops uint8 [n,160,160,3], int32 indices)."""
        n = host_x.shape[0]
        if n > self.batch:
            raise RuntimeError("batch %d exceeds the pipeline's slot size %d" % (n, self.batch))
        k, first = self.i % self.depth, self.i < self.depth
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            if not first:
                self.s_in.wait_event(self.ev_consumed[k])       # forward of batch i-depth has read this input slot
            self.x[k][:n].copy_(host_x, non_blocking=True)
            self.a[k][:n].copy_(host_audio, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        compute.wait_event(self.ev_in[k])
        if not first:
            compute.wait_event(self.ev_drained[k])              # D2H of batch i-depth has left this output slot
        x, a, o = self.x[k][:n], self.a[k][:n], self.o[k][:n]
        if self.frames:
            if self.feats is None:
                raise RuntimeError("frames mode: call set_features(hubert_feats) first")
            self.model.forward_frames(x, self.feats, a, out=o)
        else:
            self.model._check_inputs(x, a)
            self.model._run(x, a, o, _lib.F_OUT_U8_HWC if self.uint8 else _lib.F_BF16)
        self.ev_consumed[k].record(compute)
        self.ev_out[k].record(compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_out[k])
            host_out[:n].copy_(o, non_blocking=True)
            self.ev_drained[k].record(self.s_out)
        self.i += 1

    def flush(self):
        """Block until every submitted batch has landed in its host buffer."""
        self.s_out.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
