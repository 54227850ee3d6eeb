"""Paste-back blend of the caller on the device.

``FrameSynthesizer.process_batch`` (image_infer_v1/tools/frame_synthesizer/infer_api.py:333-346) ends every frame with

    result = (crop_img * mask) + (img[ymin:ymax, xmin:xmax] * (1.0 - mask));  img[ymin:ymax, xmin:xmax] = result

in numpy float64, ``mask`` = the dilated face polygon / 255 (times the optional per-frame mask file) -- per frame, on the
CPU.  ``blend_paste`` does the same arithmetic for a whole batch in one launch (``casync_blend_paste``), bit-identical to
numpy.  The steps before it (``cv2.resize`` of the crop, ``fillPoly`` + ``dilate`` of the polygon) stay with the caller:
the resize's rounding depends on the OpenCV build, and the polygon work is a few microseconds per frame.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


@torch.no_grad()
def blend_paste(frames, crops, face_mask, rects, soft_mask=None):
    """frames uint8 [B,H,W,3] (CUDA, updated IN PLACE and returned); crops uint8 [B,L,L,3]: the re-sized crop with the
    prediction pasted in; face_mask uint8 [B,L,L] (0/255); rects int32 [B,4] = (ymin, ymax, xmin, xmax) per frame;
    soft_mask float32 [B,L,L] or None.  Region b is rows ymin..ymax, cols xmin..xmax of frame b and the top-left
    (ymax-ymin) x (xmax-xmin) corner of crops[b] / face_mask[b]."""
    for t, name in ((frames, "frames"), (crops, "crops"), (face_mask, "face_mask"), (rects, "rects")):
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise RuntimeError("frames must be uint8 [B,H,W,3], got %s %s" % (frames.dtype, tuple(frames.shape)))
    b, h, w, _ = frames.shape
    if crops.dtype != torch.uint8 or crops.dim() != 4 or crops.shape[0] != b or crops.shape[1] != crops.shape[2] or crops.shape[3] != 3:
        raise RuntimeError("crops must be uint8 [B,L,L,3], got %s %s" % (crops.dtype, tuple(crops.shape)))
    ldc = crops.shape[1]
    if face_mask.dtype != torch.uint8 or tuple(face_mask.shape) != (b, ldc, ldc):
        raise RuntimeError("face_mask must be uint8 [B,L,L]")
    if soft_mask is not None and (soft_mask.dtype != torch.float32 or tuple(soft_mask.shape) != (b, ldc, ldc) or not soft_mask.is_cuda):
        raise RuntimeError("soft_mask must be a CUDA float32 [B,L,L] tensor")
    if tuple(rects.shape) != (b, 4):
        raise RuntimeError("rects must be [B,4] = (ymin, ymax, xmin, xmax)")
    if not frames.is_contiguous():
        raise RuntimeError("frames are updated in place and must be contiguous")
    crops, face_mask = crops.contiguous(), face_mask.contiguous()
    rects = rects.to(torch.int32).contiguous()
    soft = soft_mask.contiguous() if soft_mask is not None else None
    with torch.cuda.device(frames.device):
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        rc = _lib.load().casync_blend_paste(frames.data_ptr(), h, w, crops.data_ptr(), ldc, face_mask.data_ptr(),
                                            soft.data_ptr() if soft is not None else None, rects.data_ptr(), b,
                                            ctypes.c_void_p(stream))
    _lib.check(rc, "casync_blend_paste")
    return frames
